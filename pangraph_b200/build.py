"""In-tree build of libpgmm_b200.so: nvcc cross-compiles every source under csrc/ for sm_100a (no GPU needed)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpgmm_b200.so")
OBJ = os.path.join(HERE, "csrc", "_obj")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CXX = os.environ.get("CXX", "g++")
HOST_FLAGS = ["-O3", "-std=c++17", "-g", "-fPIC", "-fvisibility=hidden", "-ffp-contract=off", "-Wall", "-Wno-unused-function",
              "-I/usr/local/cuda/include"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
         "-Xptxas", "-v", "--fmad=false"]


def _stale(src, obj, deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src] + deps)


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))
    hdrs = glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    objs, jobs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s) + ".o")
        objs.append(o)
        if force or _stale(s, o, hdrs):
            if s.endswith(".cu"):
                cmd = [NVCC] + ARCH + FLAGS + ["-c", s, "-o", o]
            else:  # host-only translation units go straight to the host compiler (no fused multiply-add contraction)
                cmd = [CXX] + HOST_FLAGS + ["-c", s, "-o", o]
            jobs.append((s, cmd))
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max(1, min(len(jobs), os.cpu_count() or 1))) as ex:  # translation units compile in parallel
        for (s, cmd), r in zip(jobs, ex.map(lambda j: subprocess.run(j[1], capture_output=True, text=True), jobs)):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"compiler failed on {s}")
    if force or not os.path.exists(OUT) or any(os.path.getmtime(o) > os.path.getmtime(OUT) for o in objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", OUT] + objs + ["-Xlinker", "-Bsymbolic", "-lcudart_static", "-lpthread", "-ldl", "-lrt", "-lz"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
