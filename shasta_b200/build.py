"""Builds shasta_b200/libshasta_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the authoring container; the built .so is git-ignored but
travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libshasta_b200.so")
SOURCES = ["api.cu", "pack.cu", "gather.cu", "anchors.cu", "anchors_tc.cu", "anchors_tc2.cu", "anchors_bf16.cu", "project.cu", "project_tc.cu", "pairwise.cu", "pairwise_tc.cu", "pairwise_tc3.cu",
           "aff_softmax.cu", "aff_tc.cu", "backward.cu", "backward_pair.cu", "backward_anchor.cu", "backward_box.cu", "backward_maps.cu", "optimizer.cu", "decode.cu", "greedy.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _stale(obj, src):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src, os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tc_common.cuh"), os.path.join(CSRC, "dense_tile.cuh"), os.path.join(CSRC, "decode_fused.cuh"), os.path.join(HERE, "..", "include", "shasta_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compiles every .cu to an object (per-file, so one edit recompiles one file) and links the shared library."""
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    relink = force or not os.path.exists(LIB)
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, path):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            if src == "gather.cu":
                cmd.insert(1, "-fmad=false")  # bit-exact bilinear: no FMA contraction anywhere in that file
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed for %s" % src)
            relink = True
    if relink:
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
