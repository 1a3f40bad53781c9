"""ctypes front end of the host-compiled reference kernels (TEST INFRASTRUCTURE; see build.py, cuda_shim.h).

Slots are the oracle's dicts of int32 arrays (oracle.mutations.FIELDS order == the reference struct,
kernels3.cu:9-24).  Every function launches the reference kernel with the reference's own grid / block shape
(cuda_lib_gl.py) on the CPU and returns / fills NumPy arrays."""
import ctypes as C

import numpy as np

from . import build as _build
from .. import mutations as M

I32, F32 = np.int32, np.float32
OPS = dict(flip=0, swap_activity=1, pop_out=2, pop_in_1=3, pop_in_2=4, pop_in_3=5, pop_in_4=6, split=7, paste=8,
           simple_copy=9, copy_struct=10)
_lib = None


def available():
    return _build.available()


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.emu_rippe_contacts.restype = C.c_float
        _lib.emu_rippe_contacts.argtypes = [C.c_float, C.c_void_p]
        _lib.emu_rippe_contacts_circ.restype = C.c_float
        _lib.emu_rippe_contacts_circ.argtypes = [C.c_float, C.c_float, C.c_void_p]
        _lib.emu_evaluate_likelihood_double.restype = C.c_double
        _lib.emu_evaluate_likelihood_double.argtypes = [C.c_double, C.c_double]
        _lib.emu_factorial.restype = C.c_float
        _lib.emu_factorial.argtypes = [C.c_float]
    return _lib


def pack(slot):
    return np.ascontiguousarray(np.stack([np.asarray(slot[k], dtype=I32) for k in M.FIELDS]))


def unpack(arr):
    return {k: arr[i].copy() for i, k in enumerate(M.FIELDS)}


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def move(op, dst, src, id_a=0, id_b=0, aux=0, max_id=0, ids=None, block=256):
    """One mutation kernel: `dst` (persistent destination slot) is updated in place from `src`; `ids` is the
    kernel's contig-id side array (pop_id_contigs / split_id_contigs / id_contigs), updated in place."""
    n = len(src["pos"])
    d, s = pack(dst), pack(src)
    side = np.zeros(n, dtype=I32) if ids is None else np.ascontiguousarray(ids, dtype=I32)
    rc = lib().emu_move(OPS[op], _p(d), _p(s), _p(side), int(id_a), int(id_b), int(aux), int(max_id), n, int(block))
    assert rc == 0
    for i, k in enumerate(M.FIELDS):
        dst[k][:] = d[i]
    if ids is not None:
        ids[:] = side
    return dst


def fill_sub_index(slot, contig_a, contig_b, l_cont_a, block=512):
    n = len(slot["pos"])
    out = np.zeros(n, dtype=I32)
    s = pack(slot)
    lib().emu_fill_sub_index(_p(s), _p(out), int(contig_a), int(contig_b), int(l_cont_a), n, int(block))
    return out


def params8(p):
    return np.array([p[k] for k in ("kuhn", "lm", "c1", "slope", "d", "d_max", "fact", "v_inter")], dtype=F32)


def rippe_contacts(s, p):
    q = params8(p)
    return np.array([lib().emu_rippe_contacts(float(x), _p(q)) for x in np.atleast_1d(s)], dtype=F32)


def rippe_contacts_circ(s, s_tot, p):
    q = params8(p)
    return np.array([lib().emu_rippe_contacts_circ(float(x), float(t), _p(q)) for x, t in zip(np.atleast_1d(s), np.atleast_1d(s_tot))], dtype=F32)


def evaluate_likelihood_double(ex, ob):
    return np.array([lib().emu_evaluate_likelihood_double(float(a), float(b)) for a, b in zip(np.atleast_1d(ex), np.atleast_1d(ob))])


def factorial(n):
    return np.array([lib().emu_factorial(float(x)) for x in np.atleast_1d(n)], dtype=F32)


def _level_args(lv):
    W = int(lv.obs.shape[0])
    obs = np.ascontiguousarray(lv.obs, dtype=F32)
    collector = np.ascontiguousarray(lv.collector, dtype=I32)
    dispatcher = np.ascontiguousarray(lv.dispatcher, dtype=I32)
    sub_id = np.ascontiguousarray(lv.sub_id, dtype=I32)
    sub_len = np.ascontiguousarray(lv.sub_len, dtype=F32)
    sub_accu = np.ascontiguousarray(lv.sub_accu, dtype=I32)
    return W, obs, collector, dispatcher, sub_id, sub_len, sub_accu


def evaluate_likelihood(slot, lv, p, block=512, stride=1):
    """evaluate_likelihood as launched by cuda_lib_gl.py:545-569: the per-pixel log-likelihood vector
    (N (N - 1) / 2 upper pixels then N diagonal pixels)."""
    N = int(lv.n_frags)
    W, obs, collector, dispatcher, sub_id, sub_len, sub_accu = _level_args(lv)
    triu = N * (N - 1) // 2
    total = triu + N
    out = np.zeros(total, dtype=np.float64)
    s = pack(slot)
    n = s.shape[1]
    rep_sub = np.zeros_like(sub_id)                 # rep_id_sub_frags: only read by commented-out code
    q = params8(p)
    grid = max(1, int((total // block + 1) / stride))
    lib().emu_evaluate_likelihood(_p(obs), _p(s), n, _p(collector), _p(dispatcher), _p(sub_id), _p(rep_sub), _p(sub_len), _p(sub_accu),
                                  _p(out), _p(q), triu, total, N, W, C.c_float(float(lv.nfpb)), grid, int(block))
    return out


def sub_compute_likelihood(slot, lv, p, curr_likelihood, sub_index_no_repeats, list_rep, uniq_frags, block=512, stride=1):
    """sub_compute_likelihood as launched by stream_likelihood (cuda_lib_gl.py:2457-2530)."""
    N = int(lv.n_frags)
    W, obs, collector, dispatcher, sub_id, sub_len, sub_accu = _level_args(lv)
    u = np.asarray(sub_index_no_repeats, dtype=I32)
    r = np.asarray(list_rep, dtype=I32)
    n_u, n_rep = int(u.shape[0]), int(r.shape[0])
    if n_u == 0:
        u = np.array([-1], dtype=I32)
    if n_rep == 0:
        r = np.array([-1], dtype=I32)
    uniq = np.ascontiguousarray(uniq_frags, dtype=I32)
    n_uniq = int(uniq.shape[0])
    n_no_rep = n_u * (n_u - 1) // 2
    lim_rep_uniq = n_no_rep + n_rep * n_uniq
    lim_intra = lim_rep_uniq + n_rep * (n_rep - 1) // 2
    n_values = lim_intra + n_rep
    out = np.zeros(1, dtype=np.float64)
    cur = np.ascontiguousarray(curr_likelihood, dtype=np.float64)
    s = pack(slot)
    q = params8(p)
    grid = (n_values // block + 1) // stride + 1
    lib().emu_sub_compute_likelihood(_p(obs), _p(s), s.shape[1], _p(np.ascontiguousarray(u)), _p(np.ascontiguousarray(r)), _p(uniq),
                                     _p(collector), _p(dispatcher), _p(sub_id), _p(sub_len), _p(sub_accu), _p(out), _p(cur), _p(q),
                                     n_no_rep, lim_rep_uniq, lim_intra, n_values, n_uniq, n_rep, W, N, C.c_float(float(lv.nfpb)),
                                     int(grid), int(block))
    return float(out[0])
