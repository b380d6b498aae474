"""Builds tests/_build/libpgmm_hostlogic.so: the product's HOST sources + the reference-backed test seam (CPU only)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "_build", "libpgmm_hostlogic.so")
SRCS = [os.path.join(ROOT, "pangraph_b200", "csrc", f) for f in ("mapper.cpp", "chain.cpp", "options.cpp")] + \
       [os.path.join(ROOT, "tests", "hostlogic_backend.cpp")]


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = SRCS + [os.path.join(ROOT, "pangraph_b200", "csrc", f) for f in ("mapper.h", "chain.h", "flag_sort.h", "ksw_extd2.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-ffp-contract=off", "-Wall",
           "-I/usr/local/cuda/include", "-o", OUT] + SRCS + ["-ldl", "-lpthread"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build())
