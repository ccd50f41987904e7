"""Build libpvr_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Every .cu is a separate translation unit (host-side linkage only, no relocatable device code): objects are compiled in
parallel into pvr_habitat_b200/lib/obj/ and only the ones whose source (or any header) changed are recompiled."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libpvr_b200.so")
OBJ = os.path.join(HERE, "lib", "obj")
SOURCES = ["api.cu", "conv_gemm.cu", "preprocess.cu", "preprocess_aa.cu", "pool_head.cu", "policy.cu", "vit.cu",
           "conv3x3_patch.cu", "conv_b2b.cu", "conv_f32.cu", "policy_conv.cu", "lstm_persist.cu", "comm.cu",
           "vit_f32.cu", "small_conv.cu", "attention_mma.cu", "vit_patchify.cu", "clip_rn.cu"]


def _headers():
    hs = [os.path.join(ROOT, "include", "pvr_b200.h"), os.path.abspath(__file__)]
    hs += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return hs


def _compile(nvcc, src, obj, verbose):
    cmd = [nvcc, "-c", "-Xcompiler", "-fPIC", "-std=c++17", "-O3", "-lineinfo",
           "-gencode", "arch=compute_100a,code=sm_100a",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", obj, src]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n" + r.stdout + r.stderr)
    return r.stderr


def build(force=False, verbose=False, only=None):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    os.makedirs(OBJ, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    todo = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_t):
            if only is None or os.path.basename(s) in only or not os.path.exists(o):
                todo.append((s, o))
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in srcs]
    if not todo and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(o) for o in objs):
        return LIB
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        logs = list(ex.map(lambda so: _compile(nvcc, so[0], so[1], verbose), todo))
    if verbose:
        sys.stderr.write("".join(logs))
    r = subprocess.run([nvcc, "-shared", "-o", LIB] + objs + ["-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
