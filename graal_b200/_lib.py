"""ctypes binding of libgraal_b200.so (the C-ABI declared in include/graal_b200.h).

There is NO CPU fallback: if the shared library is missing or no CUDA device is present,
every entry point raises.  ``build()`` compiles the library in-tree with nvcc for sm_100a.
"""
import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libgraal_b200.so")
SRC = os.path.join(_HERE, "csrc", "graal_b200.cu")
HEADER = os.path.join(ROOT, "include", "graal_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false",
              "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]


class GraalError(RuntimeError):
    pass


def build(force=False, verbose=False):
    """Compile csrc/graal_b200.cu -> graal_b200/libgraal_b200.so (sm_100a, -lineinfo)."""
    deps = [SRC, os.path.join(_HERE, "csrc", "moves.cuh"), HEADER]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise GraalError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
    if verbose:
        print(r.stderr)
    return LIB_PATH


def declared_symbols():
    """Every function name declared in include/graal_b200.h."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(graal_[a-z_0-9]+)\s*\(", txt)))


_lib = None

_I, _P, _F, _D, _LL, _U = C.c_int, C.c_void_p, C.c_float, C.c_double, C.c_longlong, C.c_uint
_SIGS = {
    "graal_ctx_create": (_I, [_I, C.POINTER(_P)]),
    "graal_ctx_destroy": (None, [_P]),
    "graal_last_error": (C.c_char_p, []),
    "graal_set_stream": (_I, [_P, _P]),
    "graal_sync": (_I, [_P]),
    "graal_join": (_I, [_P]),
    "graal_version": (C.c_char_p, []),
    "graal_level_bind": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _LL, _F]),
    "graal_coo_to_lists": (_I, [_P, _P, _P, _P, _LL, _I, _P, _P, C.POINTER(_LL)]),
    "graal_set_params": (_I, [_P, C.POINTER(_F)]),
    "graal_set_math_mode": (_I, [_P, _I]),
    "graal_state_bind": (_I, [_P, _P, _I, _I]),
    "graal_relabel_contigs": (_I, [_P, _I, _P]),
    "graal_stats_relabel": (_I, [_P, _I, _P, _P]),
    "graal_apply_move": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "graal_build_candidates": (_I, [_P, _I, _I, _I, _I, _I, _U]),
    "graal_commit": (_I, [_P, _I, _I]),
    "graal_full_loglik": (_I, [_P, _I, C.POINTER(_F), _P]),
    "graal_delta_loglik": (_I, [_P, _I, _I, _I, _I, _I, _I, _P]),
    "graal_score_proposal": (_I, [_P, _I, _I, _I, _I, _I, _I, _P]),
    "graal_commit_scored": (_I, [_P, _I, _I, _I, _I, _I, _I, _I]),
    "graal_state_stats": (_I, [_P, _I, _P]),
    "graal_dist_genome": (_I, [_P, _I, _P, _P, _P, _P, _P]),
    "graal_dist_candidates": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "graal_score_step": (_I, [_P, _I, _I, _I, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "graal_fetch": (_I, [_P, _P, _P, C.c_size_t]),
    "graal_dist_histogram": (_I, [_P, _P, _P, _P, _P, _D, _D, _I, _P, _P]),
    "graal_candidate_weights": (_I, [_P, _I, _I, _P, _P, _P, _P]),
    "graal_draw_candidates": (_I, [_P, _P, _P, _D, _I, _I, _P, _D, _P, _P, _P]),
    "graal_draw_commit": (_I, [_P, _I, _I, _I, _P, _I, _I, _P, _P, _D, _I, _D, _P, _P, _P, _P, _P, C.c_size_t]),
    "graal_fetch_wait": (_I, [_P]),
    "graal_launch_count": (_LL, [_P]),
    "graal_profile_enable": (_I, [_P, _I]),
    "graal_profile_read": (_I, [_P, _I, C.POINTER(_D), C.POINTER(_LL), _I]),
    "graal_profile_counters": (_I, [_P, C.POINTER(_LL), _I]),
}


def load():
    """dlopen the library (no CUDA call is made until a context is created)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GraalError("%s not found: run graal_b200._lib.build() (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise GraalError("graal_b200 error %d: %s" % (rc, load().graal_last_error().decode()))


KERNELS = dict(FULL_CONTACTS=0, FULL_BAND=1, DELTA_CONTACTS=2, DELTA_BAND=3, BUILD=4, RELABEL=5, FULL_WINDOWS=6)
OPS = dict(COPY=0, FLIP=1, SWAP_ACTIV=2, POP_OUT=3, POP_IN_1=4, POP_IN_2=5, POP_IN_3=6, POP_IN_4=7, SPLIT=8, PASTE=9)
