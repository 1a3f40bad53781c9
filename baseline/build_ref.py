"""Build baseline/_ref/kernels3_sm100a.cubin: /root/reference/kernels3.cu compiled by nvcc for sm_100a.

The file needs two shims on CUDA 12.9: the legacy texture reference `texture<unsigned char, 2> tex`
(used only by the OpenGL helper `reorder_tex`) is gone from the toolkit, and so is the intrinsic `int2float`.  The
translation unit is composed IN MEMORY -- `#define int2float(x) __int2float_rn(x)`, then the reference text without the
texture line and without `reorder_tex` -- and piped to nvcc through /dev/stdin: no copy of the reference source is written
anywhere; only the cubin lands in baseline/_ref/ (git-ignored, shipped to the GPU box with the snapshot).
Possible only where /root/reference exists (this container)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = "/root/reference"
CUBIN = os.path.join(OUT, "kernels3_sm100a.cubin")


def available():
    return os.path.exists(os.path.join(REF, "kernels3.cu")) and os.environ.get("GRAAL_RUN_REFERENCE", "1") != "0"


def build(force=False):
    src = os.path.join(REF, "kernels3.cu")
    if not os.path.exists(src):
        raise RuntimeError("reference sources not found under %s" % REF)
    if not force and os.path.exists(CUBIN) and os.path.getmtime(CUBIN) >= max(os.path.getmtime(src), os.path.getmtime(__file__)):
        return CUBIN
    os.makedirs(OUT, exist_ok=True)
    text = open(src).read()
    tex = "texture<unsigned char, 2> tex;"
    k0 = "    __global__ void reorder_tex(unsigned char* data, int* index_new, int n_frags)"
    k1 = "//    __global__ void reorder_tex(unsigned char* data, frag* fragArray"
    if text.count(tex) != 1 or text.count(k0) != 1 or text.count(k1) != 1:
        raise RuntimeError("unexpected reference source: shim anchors not found exactly once")
    a, b = text.index(k0), text.index(k1)
    unit = "#define int2float(x) __int2float_rn(x)\n" + text[:a].replace(tex, "") + text[b:]
    cmd = [os.environ.get("NVCC", "nvcc"), "-arch=sm_100a", "-cubin", "-O3", "-lineinfo", "-x", "cu", "/dev/stdin", "-o", CUBIN]
    r = subprocess.run(cmd, input=unit, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
    return CUBIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
