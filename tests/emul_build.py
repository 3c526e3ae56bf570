"""Builds the host-emulation libraries under tests/emul/ (g++, no CUDA device needed; the CUDA toolkit's headers are
only used for the vector types).  Test infrastructure: nothing here is part of, or linked into, libflowagg.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL = os.path.join(HERE, "emul")
BUILD = os.path.join(EMUL, "_build")
CSRC = os.path.join(HERE, "..", "netobserv_ebpf_agent_b200", "csrc")


def build(name, deps, flags=()):
    """Compile tests/emul/<name>.cpp into tests/emul/_build/lib<name>.so if any of `deps` is newer."""
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(EMUL, name + ".cpp")
    extra = os.environ.get("FA_EMUL_CXXFLAGS", "").split()          # e.g. -DFA_K1_EXP=3: a kernel build variant under emulation
    flags = tuple(flags) + tuple(extra)
    so = os.path.join(BUILD, "lib" + name + ("_" + "".join(c for c in "".join(extra) if c.isalnum()) if extra else "") + ".so")
    newest = max(os.path.getmtime(p) for p in [src] + list(deps))
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        tmp = so + f".tmp{os.getpid()}"
        # -Bsymbolic: a library that defines test doubles of CUDA runtime entry points must bind its own calls to them
        # even when the real libcudart is already loaded in the process
        # -fsanitize=alignment: the device faults on a misaligned vector access that x86 silently performs (an 8-byte
        # shared-memory store to a 4-byte-aligned address got through the emulation once and cost a GPU call); UBSan's
        # alignment check sees the same thing through the CUDA vector types' alignment attributes and aborts the test
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas",
                        "-fsanitize=alignment,bounds", "-fno-sanitize-recover=alignment,bounds",   # bounds: fixed-size shared arrays
                        "-Wl,-Bsymbolic", "-I" + cuda_inc, *flags, "-o", tmp, src], check=True)
        os.replace(tmp, so)
    return so


def csrc(*names):
    return [os.path.join(CSRC, n) for n in names]
