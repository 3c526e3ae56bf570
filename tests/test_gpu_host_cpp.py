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


def build_sharded():
    exe = os.path.join(HOST, "test_sharded")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-o", exe, os.path.join(HOST, "test_sharded.cpp"), "-I/usr/local/cuda/include",
                    f"-L{PKG}", "-lflowagg", f"-Wl,-rpath,{PKG}", "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,/usr/local/cuda/lib64"],
                   check=True)
    return exe


def test_cpp_host_mirror_compiles():
    build()
    build_sharded()


@pytest.mark.gpu
def test_cpp_client_drives_every_gpu_through_fa_sharded():
    """No Python, no torch, no NCCL between the client and the GPUs: fa_sharded_* (csrc/sharded.cu) on all the GPUs of the
    box (one GPU: a 1-shard box), result bit-identical to a single engine."""
    exe = build_sharded()
    out = subprocess.run([exe, "0", "2000000", "150000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("bit-identical to one GPU") == 2, out.stdout


@pytest.mark.gpu
def test_cpp_accounter_passes_reference_tests():
    exe = build()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ok TestEvict_MaxEntries" in out.stdout and "ok TestEvict_Period" in out.stdout
