"""The reference's Accounter tests (pkg/flow/account_test.go) against the compiled C++ host mirror
(netobserv_ebpf_agent_b200/host/accounter.hpp) running on the GPU engine through the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "netobserv_ebpf_agent_b200", "host")
PKG = os.path.join(ROOT, "netobserv_ebpf_agent_b200")


def build():
    exe = os.path.join(HOST, "test_accounter")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-pthread", "-o", exe, os.path.join(HOST, "test_accounter.cpp"),
                    f"-L{PKG}", "-lflowagg", f"-Wl,-rpath,{PKG}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"],
                   check=True)
    return exe


def test_cpp_host_mirror_compiles():
    build()


@pytest.mark.gpu
def test_cpp_accounter_passes_reference_tests():
    exe = build()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ok TestEvict_MaxEntries" in out.stdout and "ok TestEvict_Period" in out.stdout
