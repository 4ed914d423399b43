"""Builds libns_b200.so (sm_100a only) in-tree with nvcc.  No JIT, no fallback."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libns_b200.so")
SOURCES = [os.path.join(CSRC, f) for f in ("ns_api.cu", "ns_multi.cu", "ns_k_gate.cu", "ns_k_fit.cu", "ns_k_post.cu", "ns_table.cpp", "ns_vcf.cpp")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("ns_math.cuh", "ns_quad.cuh", "ns_core.cuh", "ns_fit.cuh", "ns_kernels.cuh", "ns_coop.cuh", "ns_logtab.h", "ns_hostmerge.hpp", "ns_api_internal.hpp")] + [
    os.path.join(ROOT, "include", "ns_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
OBJDIR = os.path.join(HERE, "build")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    """one nvcc -c per translation unit, in parallel, then one link; objects are cached by mtime"""
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [d for d in DEPS if d not in SOURCES]
    hmt = max(os.path.getmtime(h) for h in headers)
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(OBJDIR, os.path.splitext(os.path.basename(src))[0] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(hmt, os.path.getmtime(src)):
            continue
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("NS_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(out)
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc failed building libns_b200.so")
    tmp = LIB + ".tmp"
    # CUDA runtime: the shared libcudart.so.12 (found through the rpath below, or the copy a loaded torch brings along:
    # checked on the GPU box, gpurun_out/r2c_cudart_shared.log), so that the shipped library carries only its own code;
    # NS_CUDART=static restores nvcc's default of linking the runtime in.
    cudart = [] if os.environ.get("NS_CUDART") == "static" else ["-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + cudart + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link of libns_b200.so failed")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
