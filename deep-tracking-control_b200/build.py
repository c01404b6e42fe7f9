"""Builds libdtc_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libdtc_b200.so")
SOURCES = ["dtc_env.cu", "dtc_foothold.cu", "dtc_gemm.cu", "dtc_gemm_tc.cu", "dtc_learner.cu", "dtc_gru.cu", "dtc_terrain.cu", "dtc_dp.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(PKG), "include", "dtc_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    procs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            raise FileNotFoundError(f"{sp}: every source in SOURCES must exist (a partial libdtc_b200.so is never linked)")
        obj = os.path.join(PKG, "build", src.replace(".cu", ".o"))
        objs.append(obj)
        procs.append((src, subprocess.Popen([nvcc, *FLAGS, "-c", sp, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    log = []
    for src, p in procs:
        out = p.communicate()[0].decode()
        log.append(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    r = subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-lcudart", "-ldl"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout.decode())
    with open(os.path.join(PKG, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
