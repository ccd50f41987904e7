"""Build libpvr_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libpvr_b200.so")
SOURCES = ["api.cu", "conv_gemm.cu", "preprocess.cu", "preprocess_aa.cu", "pool_head.cu", "policy.cu", "vit.cu", "conv3x3_patch.cu", "conv_b2b.cu", "conv_f32.cu", "policy_conv.cu"]


def _stale(srcs):
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = list(srcs) + [os.path.join(ROOT, "include", "pvr_b200.h"), os.path.abspath(__file__)]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not force and not _stale(srcs):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-O3", "-lineinfo",
           "-gencode", "arch=compute_100a,code=sm_100a",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           "-o", LIB] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
