"""Build oracle/_ref/libgraal_ref_emu.so: the reference's kernels3.cu compiled for the HOST with g++ through
oracle/ref_emu/cuda_shim.h (TEST INFRASTRUCTURE -- it pins the NumPy oracle against the reference's own code).

Only possible where /root/reference exists (this container); the GPU box uses the golden fixtures generated
from it (tests/golden/ref_*.npz, tests/golden/make_ref_golden.py).  Nothing of the reference is committed or copied: the
translation unit (shim + the kernel file with one line patched + the launch wrappers) is composed in memory and
piped to g++; only the library is written, to oracle/_ref/ (git-ignored)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
REF = "/root/reference"        # fixed: the reference tree of this container, never redirected by the environment
LIB = os.path.join(OUT, "libgraal_ref_emu.so")


def available():
    """The reference tree is present and its compilation / execution has not been switched off (GRAAL_RUN_REFERENCE=0:
    the tests then rely on the frozen golden vectors tests/golden/ref_kernels.npz only)."""
    return os.path.exists(os.path.join(REF, "kernels3.cu")) and os.environ.get("GRAAL_RUN_REFERENCE", "1") != "0"


def build(force=False):
    src = os.path.join(REF, "kernels3.cu")
    if not os.path.exists(src):
        raise RuntimeError("reference sources not found under %s" % REF)
    deps = [src] + [os.path.join(HERE, f) for f in ("cuda_shim.h", "curand_kernel.h", "emu_main.cpp", "build.py")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    text = open(src).read()
    needle = "extern __shared__ double res[];"
    if text.count(needle) != 1:
        raise RuntimeError("unexpected reference source: %r found %d times" % (needle, text.count(needle)))
    # dynamic shared memory of sub_compute_likelihood (one double per thread of the block) -> a static array.
    # The translation unit is composed in memory and piped to g++: no copy of the reference source is written.
    main = open(os.path.join(HERE, "emu_main.cpp")).read()
    marker = "#include REF_KERNELS"
    assert main.count(marker) == 1
    unit = main.replace(marker, '#line 1 "%s"\n' % src + text.replace(needle, "static double res[4096];") + '\n#line 1 "emu_main.cpp (after the kernels)"\n')
    cmd = ["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-w", "-x", "c++", "-I", HERE, "-", "-o", LIB]
    r = subprocess.run(cmd, input=unit, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-6000:]))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
