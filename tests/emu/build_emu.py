"""Build the host-emulation test library of the CUDA kernels (see pns_emu.h)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT_DIR = os.path.join(ROOT, "tests", "_emu")
OUT = os.path.join(OUT_DIR, "libpns_emu.so")
SRC = os.path.join(ROOT, "pednstream_b200", "csrc", "pns_kernels.cu")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "pednstream_b200", "csrc", "pns_rng.cuh"),
            os.path.join(ROOT, "pednstream_b200", "csrc", "pns_lp.cuh"),
            os.path.join(ROOT, "include", "pns_b200.h"), os.path.join(ROOT, "tests", "emu", "pns_emu.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) > os.path.getmtime(d) for d in deps):
        return OUT
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           "-DPNS_HOST_EMULATION", "-I", os.path.join(ROOT, "tests", "emu"), "-o", OUT, SRC]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
