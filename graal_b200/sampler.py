"""Headless ``sampler``: the Python surface of /root/reference/cuda_lib_gl.py (class sampler) over the
sm_100a C-ABI.  Same method names, argument meaning and return tuples as the reference, minus the
PyCUDA / OpenGL objects (gl_window, pos_vbo, col_vbo, vel, pos, raw_im_init, pbo_im_buffer).

Host code is NumPy; torch only owns the device buffers (slots, level tables, outputs) whose raw
pointers go through ctypes.  There is no CPU fallback: constructing a sampler without the CUDA
library or without a GPU raises ``GraalError``.

Reference map (cuda_lib_gl.py): __init__ :33-446, init_likelihood :448, dist_inter_genome :475-541,
eval_likelihood :543-631, setup_texture :637-665, test_copy_struct :1156-1183,
setup_rippe_parameters :1203-1214, estimate_parameters :1229-1294, explode_genome :1539-1557,
apply_replay_simu :1559-1578, modify_gl_cuda_buffer :1695-1788, step_max_likelihood :1793-1980,
step_nuisance_parameters :2022-2107, return_neighbours :2295-2331, setup_distri_frags :2363-2390,
stream_likelihood :2392-2546, temperature :2590-2603, free_gpu :2605-2613.  The MH / MTM steps are in mh.py, the remaining
entry points of the class (per-mode builders, validation step, older proposal rule, scramblers, local_flip) in variants.py.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import GraalError, check
from . import rippe as opti
from .level import FRAG_FIELDS
from .mh import MetropolisMixin
from .variants import VariantsMixin

F32, I32 = np.float32, np.int32
N_TMP_STRUCT = 13
CUR = 0            # slot of the current genome (reference: gpu_vect_frags)
CAND0 = 1          # first of the 13 collector slots (reference: collector_gpu_vect_frags)
OFF_DIST = 16 + 16 * 13   # offset of the candidates' genome distances in the output block
OFF_SUB = 16 + 2 * 16 * 13   # device draw: the normalised weights of the candidates left (sub_score)
OFF_DRAW = OFF_SUB + 16 * 13   # device draw: {score of the drawn candidate, sample_out, n_ok, status}
N_LANES = 3        # proposals of one step scored concurrently: lane q owns candidate slots CAND0 + 13*q .. +13
PARAM_FIELDS = ("kuhn", "lm", "c1", "slope", "d", "d_max", "fact", "v_inter")
PARAM_DTYPE = np.dtype([(k if k != "d_max" else "l_max", F32) for k in PARAM_FIELDS], align=True)

MODIFICATION_STR = ['eject frag', 'flip frag',
                    'pop out split insert @ left or 1', 'pop out split insert @ left or -1',
                    'pop out split insert @ right or 1', 'pop out split insert @ right or -1',
                    'pop out insert @ right or 1', 'pop out insert @ right or -1',
                    'swap activity', 'transloc_1', 'transloc_2', 'transloc_3', 'transloc_4',
                    'local_scramble d1', 'local_scramble d2', 'local_scramble d3', 'local_scramble d4']


def rippe_c1(kuhn, lm, slope):
    """c1 = float32(0.53 * (lm/kuhn)**slope * kuhn**-3) (cuda_lib_gl.py:1208) under the NumPy 1.x
    promotion rules the reference ran with: float32 ** float32 is float32 (nuisance step, slope read
    back from the float32 record), float32 ** float64 / python float is float64 (fit output)."""
    kuhn, lm = F32(kuhn), F32(lm)
    ratio = lm / kuhn
    pw = np.power(ratio, slope) if isinstance(slope, np.float32) else np.power(np.float64(ratio), np.float64(slope))
    return F32((0.53 * np.float64(pw)) * np.float64(np.power(kuhn, F32(-3))))


def _torch():
    import torch
    return torch


class _VectFrags:
    """Stand-in for the reference GPUStruct ``gpu_vect_frags``: ``copy_from_gpu()`` then ``.pos``,
    ``.id_c`` ... as NumPy arrays (gpustruct.py:173-211)."""

    def __init__(self, owner, slot):
        self._o, self._slot = owner, slot
        for k in FRAG_FIELDS:
            setattr(self, k, None)

    def copy_from_gpu(self):
        h = self._o.slot_to_host(self._slot)
        for k in FRAG_FIELDS:
            setattr(self, k, h[k])

    def as_dict(self):
        return {k: getattr(self, k) for k in FRAG_FIELDS}


def build_contact_lists(sub_coo, W, blacklisted_subs=(), mean_value_trans=0.0):
    """Upper triangle (row < col) of the prepared sub-level matrix as row-segmented contact lists.
    Restates sampler.__init__ (cuda_lib_gl.py:153-172): csr + csr.T, diagonal zeroed, the rows and
    columns of blacklisted sub-frags overwritten with mean_value_trans."""
    r, c, v = (np.asarray(a) for a in sub_coo)
    keep = r != c
    r, c, v = r[keep].astype(np.int64), c[keep].astype(np.int64), v[keep].astype(F32)
    lo, hi = np.minimum(r, c), np.maximum(r, c)
    bl = np.unique(np.asarray(list(blacklisted_subs), dtype=np.int64))
    if bl.size:
        isb = np.zeros(W, dtype=bool)
        isb[bl] = True
        keep = ~(isb[lo] | isb[hi])
        lo, hi, v = lo[keep], hi[keep], v[keep]
        al, ah = [], []
        for s in bl:
            j = np.arange(W, dtype=np.int64)
            j = j[(j != s) & ~(isb[j] & (j < s))]          # pairs among blacklisted ones counted once
            al.append(np.minimum(s, j)); ah.append(np.maximum(s, j))
        al, ah = np.concatenate(al), np.concatenate(ah)
        lo = np.concatenate([lo, al]); hi = np.concatenate([hi, ah])
        v = np.concatenate([v, np.full(al.shape[0], F32(mean_value_trans), dtype=F32)])
    order = np.lexsort((hi, lo))
    lo, hi, v = lo[order], hi[order], v[order]
    nz = v != 0
    lo, hi, v = lo[nz], hi[nz], v[nz]
    rowptr = np.zeros(W + 1, dtype=np.int64)
    np.add.at(rowptr, lo + 1, 1)
    rowptr = np.cumsum(rowptr)
    contacts = np.empty((lo.shape[0], 2), dtype=I32)
    contacts[:, 0] = hi.astype(I32)
    contacts[:, 1] = v.view(I32)
    return rowptr, contacts


def neighbour_tables(level_coo, N, blacklisted_bins=(), n_neighbors=10):
    """setup_distri_frags (cuda_lib_gl.py:2363-2390) from the sparse LEVEL matrix: for every bin the
    10 columns of its row with the largest contact counts (reversed argsort; ties are DEFINED by a
    stable sort, i.e. equal values come in decreasing column order) and p ~ v**3 (uniform when the
    row is empty)."""
    r, c, v = (np.asarray(a) for a in level_coo)
    keep = r != c
    r, c, v = r[keep].astype(np.int64), c[keep].astype(np.int64), v[keep].astype(F32)
    rows = np.concatenate([r, c]); cols = np.concatenate([c, r]); vals = np.concatenate([v, v])
    if len(blacklisted_bins):
        isb = np.zeros(N, dtype=bool)
        isb[np.asarray(list(blacklisted_bins), dtype=np.int64)] = True
        keep = ~(isb[rows] | isb[cols])
        rows, cols, vals = rows[keep], cols[keep], vals[keep]
    keep = vals != 0
    rows, cols, vals = rows[keep], cols[keep], vals[keep]
    order = np.lexsort((-cols, -vals.astype(np.float64), rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    start = np.searchsorted(rows, np.arange(N), side="left")
    end = np.searchsorted(rows, np.arange(N), side="right")
    xk = np.zeros((N, n_neighbors), dtype=I32)
    pk = np.zeros((N, n_neighbors), dtype=F32)
    nn = min(n_neighbors, N)
    for i in range(N):
        k = min(end[i] - start[i], nn)
        x = cols[start[i]:start[i] + k]
        val = vals[start[i]:start[i] + k]
        if k < nn:          # pad with zero-valued columns: largest indices first (reversed stable argsort)
            have = set(int(a) for a in cols[start[i]:end[i]])
            pad, j = [], N - 1
            while len(pad) < nn - k and j >= 0:
                if j not in have:
                    pad.append(j)
                j -= 1
            x = np.concatenate([x, np.array(pad, dtype=np.int64)])
            val = np.concatenate([val, np.zeros(len(pad), dtype=F32)])
        dat = val.astype(F32) ** 3
        if dat.sum() > 0:
            p = dat / dat.sum()
        else:
            tmp = np.ones_like(dat, dtype=F32)
            p = tmp / tmp.sum()
        xk[i, :nn] = x
        pk[i, :nn] = p
    return xk[:, :nn], pk[:, :nn]


def dist_inter_genome(prev, next_, ori, id_d, init_prev, init_next, init_ori, init_orientable,
                      blacklisted, is_repeat, n_new_frags, n_frags_4_dist):
    """Normalised neighbour / orientation distance to the initial genome (cuda_lib_gl.py:475-541),
    vectorised over bins (the reference loops in Python)."""
    n = int(n_new_frags)
    norm_distance = 3.0 * (n - n_frags_4_dist)
    keep = ~np.asarray(is_repeat, dtype=bool)
    if len(blacklisted):
        keep = keep.copy()
        keep[np.asarray(list(blacklisted), dtype=np.int64)] = False
    f = np.nonzero(keep)[0]
    p0, n0 = init_prev[f], init_next[f]
    tp, tn = prev[f], next_[f]
    p1 = np.where(tp != -1, id_d[np.maximum(tp, 0)], tp)
    n1 = np.where(tn != -1, id_d[np.maximum(tn, 0)], tn)
    d = np.full(f.shape[0], 3.0)
    d -= ((p1 == p0) & (n1 == n0)) | ((p1 == n0) & (n1 == p0))
    orientable = init_orientable[f] == 1
    flipped = orientable & (init_ori[f] != ori[f])
    swap = np.where(flipped, -1, 1)
    p1s, n1s = np.where(flipped, n1, p1), np.where(flipped, p1, n1)

    def side(t0, t1):
        same = t0 == t1
        end = same & (t0 == -1)
        t1c = np.maximum(t1, 0)
        rigid = same & ~end & (init_orientable[t1c] == 0)
        soft = same & ~end & ~rigid
        agree = soft & (init_ori[np.maximum(t0, 0)] == swap * ori[t1c])
        return 1.0 * end + 1.0 * rigid + 0.5 * soft + 0.5 * agree
    d -= np.where(orientable, side(p0, p1s) + side(n0, n1s),
                  1.0 * ((p1 == p0) | (p1 == n0)) + 1.0 * ((n1 == n0) | (n1 == p0)))
    return float(d.sum()) / norm_distance if norm_distance != 0 else 0.0


class sampler(MetropolisMixin, VariantsMixin):
    def __init__(self, use_rippe, S_o_A_frags, collector_id_repeats, frag_dispatcher,
                 id_frag_duplicated, id_frags_blacklisted,
                 n_frags, n_new_frags, init_n_sub_frags, n_new_sub_frags, np_rep_sub_frags_id,
                 hic_matrix_sub_sampled,
                 np_sub_frags_len_bp, np_sub_frags_id, np_sub_frags_accu,
                 mean_squared_frags_per_bin, norm_vect_accu,
                 S_o_A_sub_frags,
                 hic_matrix, mean_value_trans, n_iterations=0, is_simu=False,
                 device=0, rng=None, sub_sample_factor=0, device_contact_lists=None, proposal_tables=None, share_level_with=None):
        """Arguments as in cuda_lib_gl.py:33-42 without the GL objects.  ``hic_matrix_sub_sampled``
        (current level) and ``hic_matrix`` (sub level) are upper-triangle COO triples
        (rows, cols, counts) instead of dense arrays.  ``rng``: np.random.RandomState (default: the
        global np.random module, as the reference).  ``device_contact_lists`` = (rowptr int64[W+1],
        contacts int32[E, 2]) torch tensors already on the device and ``proposal_tables`` = (xk, pk) replace
        the host-side preparation of the two matrices (levels generated on the GPU).  ``share_level_with``: another
        sampler of the SAME level on the same device -- this one reuses its read-only device buffers (contact lists,
        sub-frag tables) and proposal tables: several chains per GPU hold one copy of the level."""
        if not use_rippe:
            raise GraalError("only the Rippe model exists (kernels4.cu is not part of the reference tree)")
        torch = _torch()
        if not torch.cuda.is_available():
            raise GraalError("no CUDA device: graal_b200 has no CPU path")
        self.lib = _lib.load()
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.rng = np.random if rng is None else rng
        self.o = 0
        self.use_rippe = use_rippe
        self.n_tmp_struct = self.n_modif_metropolis = N_TMP_STRUCT
        self.modification_str = MODIFICATION_STR
        self.n_iterations = n_iterations
        self.is_simu = is_simu
        self.id_frags_blacklisted = list(id_frags_blacklisted)
        self._black_set = set(int(f) for f in self.id_frags_blacklisted)          # O(1) membership on the step path
        self.id_frag_duplicated = id_frag_duplicated
        self.np_id_frag_duplicated = np.asarray(id_frag_duplicated, dtype=I32)
        self._dup_set = set(int(f) for f in self.np_id_frag_duplicated)
        self._n_cand_cache = {}
        self._cdf_cache = {}
        disp = np.asarray(frag_dispatcher).reshape(-1, 2)
        self._identity_dispatch = bool(len(self._dup_set) == 0 and np.all(disp[:, 1] - disp[:, 0] == 1) and
                                       np.array_equal(np.asarray(collector_id_repeats)[disp[:, 0]], np.arange(disp.shape[0])))
        self.n_frags, self.n_new_frags = I32(n_frags), I32(n_new_frags)
        self.init_n_sub_frags, self.n_new_sub_frags = I32(init_n_sub_frags), I32(n_new_sub_frags)
        self.uniq_frags = np.setdiff1d(np.arange(n_frags, dtype=I32), self.np_id_frag_duplicated).astype(I32)
        self.n_frags_uniq = I32(len(self.uniq_frags))
        self.collector_id_repeats = np.asarray(collector_id_repeats, dtype=I32)
        self.frag_dispatcher = np.asarray(frag_dispatcher, dtype=I32).reshape(-1, 2)
        self.S_o_A_frags, self.S_o_A_sub_frags = S_o_A_frags, S_o_A_sub_frags
        self.mean_value_trans = mean_value_trans
        self.np_sub_frags_len_bp = np.ascontiguousarray(np_sub_frags_len_bp, dtype=F32).reshape(-1, 3)
        self.np_sub_frags_id = np.ascontiguousarray(np_sub_frags_id, dtype=I32).reshape(-1, 4)
        self.np_sub_frags_accu = np.ascontiguousarray(np_sub_frags_accu, dtype=I32).reshape(-1, 3)
        self.mean_squared_frags_per_bin = F32(mean_squared_frags_per_bin)
        self.norm_vect_accu = norm_vect_accu
        self._level_coo = hic_matrix_sub_sampled if share_level_with is None else share_level_with._level_coo
        self.n_modif_metropolis = N_TMP_STRUCT
        n = int(n_new_frags)
        # ---- sub-level matrix -> contact lists (cuda_lib_gl.py:153-172, 194)
        black_subs = []
        for f in self.id_frags_blacklisted:
            da = self.np_sub_frags_id[S_o_A_frags["id_d"][f]]
            black_subs.extend(int(da[k]) for k in range(da[3]))
        # ---- context (needed first: the contact lists are built on the device)
        ctx = C.c_void_p()
        check(self.lib.graal_ctx_create(device, C.byref(ctx)))
        self.ctx = ctx
        self.stream = torch.cuda.Stream(device=self.device)
        check(self.lib.graal_set_stream(self.ctx, C.c_void_p(self.stream.cuda_stream)))
        if share_level_with is not None:
            rowptr, contacts = share_level_with.d_rowptr, share_level_with.d_contacts
            proposal_tables = (share_level_with.distri_xk, share_level_with.distri_pk)
        elif device_contact_lists is not None:
            rowptr, contacts = device_contact_lists
        elif black_subs or os.environ.get("GRAAL_HOST_LISTS", "0") == "1":
            rowptr, contacts = build_contact_lists(hic_matrix, int(init_n_sub_frags), black_subs, mean_value_trans)
        else:
            rowptr, contacts = self.device_contact_lists(hic_matrix, int(init_n_sub_frags))
        self.n_contacts = int(contacts.shape[0])
        # ---- proposal tables (cuda_lib_gl.py:444-445)
        self.n_neighbors = 10
        black_bins = [int(S_o_A_frags["id_d"][f]) for f in self.id_frags_blacklisted]
        if proposal_tables is None:
            self.distri_xk, self.distri_pk = neighbour_tables(hic_matrix_sub_sampled, int(n_frags), black_bins, self.n_neighbors)
        else:
            self.distri_xk, self.distri_pk = proposal_tables
        # ---- device buffers (torch owns them)
        dev = self.device
        t = lambda a: a.to(dev) if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        if share_level_with is not None:
            o = share_level_with
            self.d_sub_id, self.d_sub_len, self.d_sub_accu = o.d_sub_id, o.d_sub_len, o.d_sub_accu
            self.d_collector, self.d_dispatcher = o.d_collector, o.d_dispatcher
        else:
            self.d_sub_id, self.d_sub_len, self.d_sub_accu = t(self.np_sub_frags_id), t(self.np_sub_frags_len_bp), t(self.np_sub_frags_accu)
            self.d_collector, self.d_dispatcher = t(self.collector_id_repeats), t(self.frag_dispatcher)
        self.d_rowptr, self.d_contacts = t(rowptr), t(contacts)
        self.ld = (n + 31) // 32 * 32
        self.n_slots = 1 + N_TMP_STRUCT * N_LANES + 4          # + pop, trans1, trans2, forward (the MH / MTM variants, mh.py)
        host = np.zeros((self.n_slots, len(FRAG_FIELDS), self.ld), dtype=I32)
        for fi, k in enumerate(FRAG_FIELDS):
            host[CUR, fi, :n] = np.ones(n, dtype=I32) if k == "ori" else np.asarray(S_o_A_frags[k], dtype=I32)   # Q5
        host[CAND0:, FRAG_FIELDS.index("ori"), :] = 1           # collector initial content (cuda_lib_gl.py:269-287)
        host[CAND0:, FRAG_FIELDS.index("activ"), :] = 1
        self.d_slots = t(host)
        # [0] full, [1] test, [4:8] stats, [8] dist, [16:224] deltas of <= 16 proposals, [224:432] their genome distances,
        # [432:640] sub_score and [640:644] {score, sample_out, n_ok, status} of the draw made on the device
        self.d_out = torch.zeros(OFF_DRAW + 16, dtype=torch.float64, device=dev)
        self.d_max_id = torch.zeros(4, dtype=torch.int32, device=dev)
        self.d_sel = torch.zeros(8, dtype=torch.int32, device=dev)
        # the candidate draw and the commit on the device, right behind the scores (GRAAL_DEVICE_DRAW=0: on the host after the fetch)
        self.device_draw = os.environ.get("GRAAL_DEVICE_DRAW", "1") != "0"
        self.h_out = torch.zeros_like(self.d_out, device="cpu").pin_memory()
        torch.cuda.synchronize(dev)
        check(self.lib.graal_level_bind(self.ctx, int(n_frags), n, int(init_n_sub_frags),
                                        self.d_sub_id.data_ptr(), self.d_sub_len.data_ptr(), self.d_sub_accu.data_ptr(),
                                        self.d_collector.data_ptr(), self.d_dispatcher.data_ptr(),
                                        self.d_rowptr.data_ptr(), self.d_contacts.data_ptr(), self.n_contacts,
                                        float(self.mean_squared_frags_per_bin)))
        check(self.lib.graal_state_bind(self.ctx, self.d_slots.data_ptr(), self.ld, self.n_slots))
        # ---- bookkeeping for dist_inter_genome (cuda_lib_gl.py:226-233, 452-469)
        self.np_init_prev = np.copy(S_o_A_frags['prev'])
        self.np_init_next = np.copy(S_o_A_frags['next'])
        self.np_init_ori = np.ones(n, dtype=I32)
        self.np_init_orientable = (self.np_sub_frags_id[S_o_A_frags['id_d'], 3] > 1).astype(I32)
        self.h_id_d = np.asarray(S_o_A_frags["id_d"], dtype=I32).copy()      # never modified by any kernel
        self.d_init_prev, self.d_init_next = t(self.np_init_prev.astype(I32)), t(self.np_init_next.astype(I32))
        self.d_init_orientable = t(self.np_init_orientable)
        self.gpu_vect_frags = _VectFrags(self, CUR)
        self.collector_gpu_vect_frags = [_VectFrags(self, CAND0 + k) for k in range(N_TMP_STRUCT)]
        self.define_repeats()
        skip = self.is_repeat.astype(np.uint8)
        if self.id_frags_blacklisted:
            skip[np.asarray(self.id_frags_blacklisted, dtype=np.int64)] = 1
        self.d_dist_skip = t(skip)
        # raw addresses of the buffers the step path passes on every call
        self._p_out, self._p_hout = self.d_out.data_ptr(), self.h_out.data_ptr()
        self._n_out_bytes = self.d_out.numel() * self.d_out.element_size()
        self._h_out_np = self.h_out.numpy()
        self._p_dist_args = (self.d_init_prev.data_ptr(), self.d_init_next.data_ptr(), self.d_init_orientable.data_ptr(),
                             self.d_dist_skip.data_ptr())
        self.param_simu = None
        self.likelihood_t = None
        self.incremental_likelihood = False     # see step_max_likelihood
        self.incremental_resync = 256
        self._inc_valid, self._inc_age = False, 0
        self._remove_cache = {}
        self.score = np.zeros(0)
        self.delta_scores = np.zeros(0)
        self.gpu_launches_at_start = self.lib.graal_launch_count(self.ctx)

    def device_contact_lists(self, sub_coo, W):
        """The sub-level matrix as contact lists, built on the device (graal_coo_to_lists): the COO triple is uploaded as it
        is; keys (min, max), radix sort, duplicate sums, zero / diagonal removal and the row prefix sum run on the GPU."""
        torch = self.torch
        r, c, v = (np.asarray(a) for a in sub_coo)
        n = int(r.shape[0])
        if n == 0:
            return np.zeros(W + 1, dtype=np.int64), np.zeros((0, 2), dtype=I32)
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.device)
        d_r, d_c, d_v = t(r, I32), t(c, I32), t(v, F32)
        rowptr = torch.zeros(W + 1, dtype=torch.int64, device=self.device)
        contacts = torch.zeros((n, 2), dtype=torch.int32, device=self.device)
        n_out = C.c_longlong(0)
        torch.cuda.synchronize(self.device)
        check(self.lib.graal_coo_to_lists(self.ctx, d_r.data_ptr(), d_c.data_ptr(), d_v.data_ptr(), n, int(W),
                                          rowptr.data_ptr(), contacts.data_ptr(), C.byref(n_out)))
        return rowptr, contacts[:int(n_out.value)].contiguous()

    @classmethod
    def from_inputs(cls, inp, device=0, rng=None, device_contact_lists=None, proposal_tables=None, share_level_with=None):
        """Build from ``graal_b200.level.SamplerInputs`` (what simulation.__init__ assembles)."""
        return cls(True, inp.S_o_A_frags, inp.collector_id_repeats, inp.frag_dispatcher, inp.id_frag_duplicated,
                   inp.id_frags_blacklisted, inp.n_frags, inp.n_new_frags, inp.init_n_sub_frags, inp.n_new_sub_frags,
                   inp.np_rep_sub_frags_id, inp.level_coo, inp.np_sub_frags_len_bp, inp.np_sub_frags_id,
                   inp.np_sub_frags_accu, inp.mean_squared_frags_per_bin, inp.norm_vect_accu, inp.S_o_A_sub_frags,
                   inp.sub_coo, inp.mean_value_trans, device=device, rng=rng,
                   device_contact_lists=device_contact_lists, proposal_tables=proposal_tables, share_level_with=share_level_with)

    # ------------------------------------------------------------------ plumbing
    def sync(self):
        check(self.lib.graal_sync(self.ctx))

    def _ptr(self, tensor, offset=0):
        return C.c_void_p(tensor.data_ptr() + offset * tensor.element_size())

    def _fetch(self):
        """One D2H of the whole output block (pinned), after the stream has drained."""
        # (lanes joined, copy on the context stream, wait: one library call instead of torch's stream machinery)
        check(self.lib.graal_fetch(self.ctx, self._p_out, self._p_hout, self._n_out_bytes))
        return self._h_out_np

    def slot_to_host(self, slot):
        self.sync()
        n = int(self.n_new_frags)
        a = self.d_slots[slot, :, :n].cpu().numpy()
        return {k: a[i].copy() for i, k in enumerate(FRAG_FIELDS)}

    def slot_from_host(self, slot, arrays):
        self._inc_valid = False          # the cached likelihood no longer describes the state / parameters
        n = int(self.n_new_frags)
        self.sync()
        for i, k in enumerate(FRAG_FIELDS):
            self.d_slots[slot, i, :n] = self.torch.from_numpy(np.ascontiguousarray(arrays[k], dtype=I32)).to(self.device)
        self.torch.cuda.synchronize(self.device)
        check(self.lib.graal_state_bind(self.ctx, self.d_slots.data_ptr(), self.ld, self.n_slots))   # drops cached geometry

    @property
    def gpu_launches(self):
        return int(self.lib.graal_launch_count(self.ctx))

    # ------------------------------------------------------------------ cuda_lib_gl.py:452-469
    def define_repeats(self):
        s = self.S_o_A_frags
        rep_ids = np.unique(np.asarray(s["id_d"])[np.asarray(s["id"]) != np.asarray(s["id_d"])])
        self.is_repeat = np.isin(s["id_d"], rep_ids)
        self.n_frags_duplicated = int(self.is_repeat.sum())
        self.n_frags_4_dist = len(np.unique(list(self.id_frags_blacklisted) + list(np.nonzero(self.is_repeat)[0])))

    def setup_texture(self):
        """cuda_lib_gl.py:637-665: selects the data matrix; the contact lists are already resident."""
        self.data = self.d_contacts

    # ------------------------------------------------------------------ parameters
    def setup_rippe_parameters(self, param, d_max):
        """cuda_lib_gl.py:1203-1214."""
        kuhn, lm, slope, d, fact = param
        kuhn, lm = F32(kuhn), F32(lm)
        c1 = rippe_c1(kuhn, lm, slope)
        return np.array([(kuhn, lm, c1, F32(slope), F32(d), F32(d_max), F32(fact), self.mean_value_trans)], dtype=PARAM_DTYPE)

    def _set_device_params(self, p):
        arr = (C.c_float * 8)(*[float(x) for x in p[0]])
        check(self.lib.graal_set_params(self.ctx, arr))

    def set_math_mode(self, mode):
        """0: the reference's float32 chain op for op; 1: log-space float64 evaluation of in-band pixels;
        2 (default): tabulated law (see include/graal_b200.h)."""
        self._inc_valid = False          # the cached likelihood no longer describes the state / parameters
        check(self.lib.graal_set_math_mode(self.ctx, int(mode)))
        if self.param_simu is not None:
            self._set_device_params(self.param_simu)

    def set_parameters(self, param, d_max):
        self._inc_valid = False          # the cached likelihood no longer describes the state / parameters
        self.param_simu = self.setup_rippe_parameters(param, d_max)
        self._set_device_params(self.param_simu)

    def distance_histogram(self, max_dist_kb, size_bin_kb):
        """Device reduction of the O(W^2) loop of estimate_parameters (cuda_lib_gl.py:1236-1270)."""
        torch = self.torch
        bins = np.arange(size_bin_kb, max_dist_kb + size_bin_kb, size_bin_kb)
        nb = len(bins)
        sub = self.S_o_A_sub_frags
        t = lambda k: torch.from_numpy(np.ascontiguousarray(sub[k], dtype=I32)).to(self.device)
        d_idc, d_st, d_ln, d_pos = t("id_c"), t("start_bp"), t("len_bp"), t("pos")
        d_sum = torch.zeros(nb, dtype=torch.float64, device=self.device)
        d_cnt = torch.zeros(nb, dtype=torch.int64, device=self.device)
        torch.cuda.synchronize(self.device)
        check(self.lib.graal_dist_histogram(self.ctx, self._ptr(d_idc), self._ptr(d_st), self._ptr(d_ln), self._ptr(d_pos),
                                            float(max_dist_kb), float(size_bin_kb), nb, self._ptr(d_sum), self._ptr(d_cnt)))
        self.sync()
        sums, cnts = d_sum.cpu().numpy(), d_cnt.cpu().numpy()
        mean = np.full(nb, 1e-10, dtype=F32)
        ok = (cnts > 0) & (sums > 0)
        mean[ok] = (sums[ok] / cnts[ok]).astype(F32)
        return bins, mean, sums, cnts

    def estimate_parameters(self, max_dist_kb, size_bin_kb):
        """cuda_lib_gl.py:1229-1294."""
        self.bins, self.mean_contacts, _, _ = self.distance_histogram(max_dist_kb, size_bin_kb)
        p, self.y_estim = opti.estimate_param_rippe(self.mean_contacts, self.bins)
        estim_max_dist = opti.estimate_max_dist_intra(p, self.mean_value_trans)
        self.set_parameters(p, estim_max_dist)

    # ------------------------------------------------------------------ likelihood
    def eval_likelihood(self, test_params=None):
        """cuda_lib_gl.py:543-631 (and :1986-2019 with test parameters): full log-likelihood."""
        arr = None
        if test_params is not None:
            arr = (C.c_float * 8)(*[float(x) for x in test_params[0]])
        check(self.lib.graal_full_loglik(self.ctx, CUR, arr, self._ptr(self.d_out, 0)))
        return np.float64(self._fetch()[0])

    def init_likelihood(self):
        self.likelihood_t = self.eval_likelihood()

    def modify_gl_cuda_buffer(self, id_fi=0, dt=0):
        """cuda_lib_gl.py:1695-1788, structure part only: relabel contig ids, return max_id."""
        check(self.lib.graal_relabel_contigs(self.ctx, CUR, self._ptr(self.d_max_id)))
        self.sync()
        return I32(int(self.d_max_id[0].item()))

    relabel_contigs = modify_gl_cuda_buffer

    def perform_modifications(self, id_fA, id_fB, max_id=-1, mask=0x1FFF):
        """new_perform_modificationS (cuda_lib_gl.py:1045-1048): the 13 candidates into the collector slots."""
        check(self.lib.graal_build_candidates(self.ctx, CUR, CAND0, int(id_fA), int(id_fB), int(max_id), mask))

    def apply_move(self, src_slot, dst_slot, op, id_fA, id_fB=0, aux=0, max_id=0):
        """One mutation kernel of the reference (``op`` = name in _lib.OPS) between two slots; returns
        max(id_c) of the destination (the ga.max that follows pop_out / split, cuda_lib_gl.py:857,934)."""
        self._inc_valid = False          # the cached likelihood no longer describes the state / parameters
        check(self.lib.graal_apply_move(self.ctx, int(src_slot), int(dst_slot), _lib.OPS[op], int(id_fA), int(id_fB),
                                        int(aux), int(max_id), self._ptr(self.d_max_id, 1)))
        self.sync()
        return int(self.d_max_id[1].item())

    def test_copy_struct(self, id_fA, id_f_sampled, mode, max_id):
        """cuda_lib_gl.py:1156-1183: rebuild the sampled candidate, commit it to the current slot."""
        self._inc_valid = False          # the cached likelihood no longer describes the state / parameters
        mode = int(mode)
        mask = (1 << mode) if mode < 9 else 0x1E00
        self.perform_modifications(id_fA, id_f_sampled, max_id, mask)
        check(self.lib.graal_commit(self.ctx, CUR, CAND0 + mode))

    def explode_genome(self, dt=0):
        """cuda_lib_gl.py:1539-1557."""
        self._inc_valid = False          # the cached likelihood no longer describes the state / parameters
        for i in range(int(self.n_new_frags)):
            check(self.lib.graal_relabel_contigs(self.ctx, CUR, self._ptr(self.d_max_id)))
            self.test_copy_struct(i, 0, 0, -1)
        self.sync()

    def apply_replay_simu(self, id_fA, id_fB, op_sampled, dt=0):
        """cuda_lib_gl.py:1559-1578."""
        self._inc_valid = False          # the cached likelihood no longer describes the state / parameters
        check(self.lib.graal_relabel_contigs(self.ctx, CUR, self._ptr(self.d_max_id)))
        self.test_copy_struct(id_fA, id_fB, op_sampled, -1)

    # ------------------------------------------------------------------ proposals
    def return_neighbours(self, id_fA, delta0):
        """cuda_lib_gl.py:2295-2331."""
        ori_id = int(self.h_id_d[id_fA])
        delta = min(self.n_neighbors, delta0)
        distri = self.distri_pk[ori_id]
        n_nonzero = self._n_cand_cache.get(ori_id)
        if n_nonzero is None:
            n_nonzero = self._n_cand_cache[ori_id] = int(np.count_nonzero(distri))
        n_max_candidates = min(delta, n_nonzero)
        init_id = self._choice_no_replace(ori_id, n_max_candidates)
        if self._identity_dispatch:             # no duplicated bins: every dispatcher range is the bin itself
            return [e for e in init_id.tolist() if e not in self._black_set]
        out = []
        if ori_id in self._dup_set:
            d = self.frag_dispatcher[ori_id]
            out.extend(np.setdiff1d(self.collector_id_repeats[d[0]:d[1]], id_fA))
        for id_fB in init_id:
            d = self.frag_dispatcher[id_fB]
            out.extend(self.collector_id_repeats[d[0]:d[1]])
        return [int(e) for e in out if int(e) not in self._black_set]

    def _choice_no_replace(self, ori_id, size):
        """``self.rng.choice(self.distri_xk[ori_id], size, p=self.distri_pk[ori_id], replace=False)`` of a legacy
        RandomState without its argument validation: the same draws in the same order (numpy/random/mtrand.pyx, the
        replace=False branch: `size - n_found` uniforms per round, searched in the cumulative weights with the bins already
        found zeroed, first occurrences kept in draw order).  The cumulative weights of the FIRST round only depend on the
        bin: cached.  Pinned by tests/test_reference_host_logic.py against the reference's own lines."""
        rs = getattr(self.rng, "random_sample", None)
        xk = self.distri_xk[ori_id]
        if rs is None or size <= 0:
            return self.rng.choice(xk, size, p=self.distri_pk[ori_id], replace=False)
        ent = self._cdf_cache.get(ori_id)
        if ent is None:
            p = np.array(self.distri_pk[ori_id], dtype=np.float64)
            cdf = np.cumsum(p)
            cdf /= cdf[-1]
            ent = self._cdf_cache[ori_id] = (p, cdf)
        p0, cdf = ent
        found = []
        p = None
        while len(found) < size:
            x = rs(size - len(found))
            if found:
                if p is None:
                    p = p0.copy()
                p[found] = 0
                cdf = np.cumsum(p)
                cdf /= cdf[-1]
            for j in cdf.searchsorted(x, side="right").tolist():
                if j not in found:            # first occurrence of every new index, in draw order (np.unique + sorted first indices);
                    found.append(j)           # an index found in an earlier round has zero weight now and cannot come back
        return xk[found]

    def temperature(self, t, n_step):
        return 1.0

    def score_neighbours(self, id_fA, id_neighbours, with_dist=False):
        """stream_likelihood (cuda_lib_gl.py:2392-2546) for every neighbour: candidates + deltas, queued
        on the stream; results land in d_out[16 + 13*x + j]."""
        n = len(id_neighbours)
        if n == 0:
            return
        fbs = (C.c_int32 * n)(*[int(f) for f in id_neighbours])
        # one call for the whole neighbour loop; with_dist: the genome distance of every candidate comes back with
        # the scores (no second round trip).  Proposal x is built into the candidate slots of lane x % N_LANES.
        check(self.lib.graal_score_step(self.ctx, CUR, CAND0, int(id_fA), fbs, n, -1, self._p_out + 8 * 16,
                                        self._p_dist_args[0], self._p_dist_args[1], self._p_dist_args[2], self._p_dist_args[3],
                                        (self._p_out + 8 * OFF_DIST) if with_dist else None))

    def step_max_likelihood(self, id_fA, delta, size_block=512, dt=0, t=0, n_step=1):
        """cuda_lib_gl.py:1793-1980.  Returns (o, n_contigs, min_len, mean_len_bp, max_len, op_sampled,
        id_f_sampled, dist, F_t)."""
        self.step_begin(id_fA, delta, t, n_step)
        return self.step_end(t, n_step)

    MAX_PROPOSALS = 16      # proposals scored per device round trip (output block / band-delta history of the library)

    def step_begin(self, id_fA, delta, t=0, n_step=1, device_draw=True):
        """First half of step_max_likelihood: everything up to the device round trip is ENQUEUED (statistics, relabel, the
        proposals on their lanes, the full likelihood beside them); nothing is waited for.  Several chains on one GPU call
        step_begin on each chain, then step_end on each (graal_b200.replica.step_chains): their kernels overlap."""
        lib = self.lib
        self._step_fA = id_fA
        check(lib.graal_stats_relabel(self.ctx, CUR, self._ptr(self.d_out, 4), self._ptr(self.d_max_id)))
        if id_fA in self._black_set:
            self.id_neighbours = None
            return
        id_neighbours = self.return_neighbours(id_fA, delta)
        id_neighbours.sort()
        self.id_neighbours = id_neighbours
        # incremental mode (off by default; the reference recomputes, cuda_lib_gl.py:1828-1848): the score of
        # the candidate committed by the previous step IS the likelihood of the current state; the full pass
        # only runs to resynchronise
        self._step_incremental = self.incremental_likelihood and self._inc_valid and self._inc_age < self.incremental_resync
        # the proposals go to their lanes first; the full likelihood of the current state does not depend
        # on them and runs on the context stream next to them
        self.score_neighbours(id_fA, id_neighbours[:self.MAX_PROPOSALS], with_dist=True)
        if not self._step_incremental:
            check(lib.graal_full_loglik(self.ctx, CUR, None, self._ptr(self.d_out, 0)))
        # the draw (cuda_lib_gl.py:1899-1947) and the commit (:1952) follow on the device without waiting for the host: the
        # NEXT uniform of the stream goes along, and is only consumed if the draw needed it (known after the fetch)
        self._drawn_on_device = False
        rs = getattr(self.rng, "random_sample", None)
        n_nb = len(id_neighbours)
        if (self.device_draw and device_draw and 0 < n_nb <= self.MAX_PROPOSALS and rs is not None and hasattr(self.rng, "get_state")
                and self.temperature(t, n_step) == 1.0 and getattr(self, "_fast_weights", True)):
            self._rng_before_draw = self.rng.get_state()
            u = float(rs())
            fbs = (C.c_int32 * n_nb)(*id_neighbours)
            check(lib.graal_draw_commit(self.ctx, CUR, CAND0, int(id_fA), fbs, n_nb, -1, self._p_out + 8 * 16, self._p_out,
                                        float(self.likelihood_t) if self._step_incremental else 0.0, 1 if self._step_incremental else 0, u,
                                        self.d_sel.data_ptr(), self._p_out + 8 * OFF_SUB, self._p_out + 8 * OFF_DRAW,
                                        self._p_out, self._p_hout, self._n_out_bytes))
            self._drawn_on_device = True

    def step_end(self, t=0, n_step=1):
        """Second half: the one device round trip, the candidate draw (cuda_lib_gl.py:1899-1947), the commit."""
        lib = self.lib
        id_fA = self._step_fA
        if self.id_neighbours is not None:
            id_neighbours = self.id_neighbours
            n_neighbours = len(id_neighbours)
            if self._drawn_on_device:                 # the block left for the host before the commit kernels: wait for the copy only
                check(lib.graal_fetch_wait(self.ctx))
                out = self._h_out_np
            else:
                out = self._fetch()
            incremental = self._step_incremental
            likelihood_t = np.float64(self.likelihood_t) if incremental else np.float64(out[0])
            self._inc_age = self._inc_age + 1 if incremental else 0
            self._inc_valid = True
            self.likelihood_t = likelihood_t
            n_contigs, min_len, mean_len_bp, max_len = int(out[4]), int(out[5]), out[6], int(out[7])
            n_first = min(n_neighbours, self.MAX_PROPOSALS)
            last_chunk = 0
            if n_neighbours <= self.MAX_PROPOSALS:
                # one round trip (the usual case): nothing but the draw stands between the fetch and the commit -- the pinned
                # block is only overwritten by the next fetch, copies are taken from it once
                self.delta_scores = np.array(out[16:16 + n_first * N_TMP_STRUCT], dtype=np.float64)
                dist_all = out[OFF_DIST:OFF_DIST + n_first * N_TMP_STRUCT]
            else:
                deltas = [np.array(out[16:16 + n_first * N_TMP_STRUCT], dtype=np.float64)]
                dists = [np.array(out[OFF_DIST:OFF_DIST + n_first * N_TMP_STRUCT], dtype=np.float64)]
                # more neighbours than one round trip holds (repeats expand every partner to all its copies,
                # cuda_lib_gl.py:2316-2327): further chunks of <= 16 proposals, one round trip each
                for c0 in range(self.MAX_PROPOSALS, n_neighbours, self.MAX_PROPOSALS):
                    chunk = id_neighbours[c0:c0 + self.MAX_PROPOSALS]
                    self.score_neighbours(id_fA, chunk, with_dist=True)
                    o2 = self._fetch()
                    deltas.append(np.array(o2[16:16 + len(chunk) * N_TMP_STRUCT], dtype=np.float64))
                    dists.append(np.array(o2[OFF_DIST:OFF_DIST + len(chunk) * N_TMP_STRUCT], dtype=np.float64))
                    last_chunk = c0
                self.delta_scores = np.concatenate(deltas)
                dist_all = np.concatenate(dists)
            self.score = self.delta_scores + likelihood_t
            drawn = self._drawn_on_device and out[OFF_DRAW + 3] == 0.0
            if self._drawn_on_device and not drawn:
                self.rng.set_state(self._rng_before_draw)        # NaN scores: nothing was drawn or committed, the host path decides
            if drawn:
                sample_out, n_ok = int(out[OFF_DRAW + 1]), int(out[OFF_DRAW + 2])
                if n_ok <= 1:
                    self.rng.set_state(self._rng_before_draw)    # one candidate left: the reference draws nothing (:1936-1939)
                self.sub_score = np.array(out[OFF_SUB:OFF_SUB + n_ok], dtype=np.float64)
                x = sample_out // N_TMP_STRUCT
                id_f_sampled = id_neighbours[x]
                op_sampled = sample_out % N_TMP_STRUCT
            else:
                sample_out = self._sample(self.score, self.temperature(t, n_step))
                x = sample_out // N_TMP_STRUCT
                id_f_sampled = id_neighbours[x]
                op_sampled = sample_out % N_TMP_STRUCT
                # the band delta of a scored proposal is only remembered for the chunk scored last
                x_lib = x - last_chunk if x >= last_chunk else -1
                check(lib.graal_commit_scored(self.ctx, CUR, CAND0, int(id_fA), int(id_f_sampled), -1, int(op_sampled), int(x_lib)))
            o = self.score[sample_out]
            self.o = o
            # dist_inter_genome of the committed candidate (cuda_lib_gl.py:1962) came back with the scores
            norm_distance = 3.0 * (int(self.n_new_frags) - self.n_frags_4_dist)
            dist = float(dist_all[sample_out]) / norm_distance if norm_distance != 0 else 0.0
        else:
            o = self.o
            out = self._fetch()
            n_contigs, min_len, mean_len_bp, max_len = int(out[4]), int(out[5]), out[6], int(out[7])
            op_sampled, id_f_sampled = -1, id_fA
            dist = self.dist_inter_genome_device()
        F_t = self.temperature(t, n_step)
        self.likelihood_t = o
        return o, n_contigs, min_len, mean_len_bp, max_len, op_sampled, id_f_sampled, dist, F_t

    def dist_inter_genome_device(self):
        """dist_inter_genome of the current genome (cuda_lib_gl.py:475-541) reduced on the device: one
        scalar comes back instead of the 14 state arrays."""
        check(self.lib.graal_dist_genome(self.ctx, CUR, self._ptr(self.d_init_prev), self._ptr(self.d_init_next),
                                         self._ptr(self.d_init_orientable), self._ptr(self.d_dist_skip), self._ptr(self.d_out, 8)))
        norm_distance = 3.0 * (int(self.n_new_frags) - self.n_frags_4_dist)
        return float(self._fetch()[8]) / norm_distance if norm_distance != 0 else 0.0

    def step_device(self, id_fA, id_neighbours, id_f_sampled, op_sampled, full=True):
        """Device-resident replay of one step: the same kernel sequence as step_max_likelihood with the
        proposal and the sampled candidate supplied by the caller -- nothing is copied to the host and
        the stream is not synchronised (bench.py's resident-input timing, replay of a recorded run)."""
        lib = self.lib
        check(lib.graal_stats_relabel(self.ctx, CUR, self._ptr(self.d_out, 4), self._ptr(self.d_max_id)))
        if op_sampled < 0:
            return
        for c0 in range(0, len(id_neighbours), self.MAX_PROPOSALS):
            self.score_neighbours(id_fA, id_neighbours[c0:c0 + self.MAX_PROPOSALS], with_dist=True)
        if full:                                   # False: a step of the incremental mode between two resyncs
            check(lib.graal_full_loglik(self.ctx, CUR, None, self._ptr(self.d_out, 0)))
        x = id_neighbours.index(id_f_sampled) if id_f_sampled in id_neighbours else -1
        last_chunk = (max(len(id_neighbours), 1) - 1) // self.MAX_PROPOSALS * self.MAX_PROPOSALS
        x = x - last_chunk if x >= last_chunk else -1
        check(lib.graal_commit_scored(self.ctx, CUR, CAND0, int(id_fA), int(id_f_sampled), -1, int(op_sampled), x))

    def _sample(self, score, F_t):
        """Candidate filtering and draw (cuda_lib_gl.py:1899-1947)."""
        nt = N_TMP_STRUCT
        rs = getattr(self.rng, "random_sample", None)
        if F_t == 1.0 and rs is not None and getattr(self, "_fast_weights", True):
            # the same float64 arithmetic in C (NumPy's operation order, pairwise sums included; graal_candidate_weights):
            # ~40 microseconds of small-array NumPy calls off the critical path between two steps
            n = len(score)
            buf = self._weights_buf.get(n) if hasattr(self, "_weights_buf") else None
            if buf is None:
                if not hasattr(self, "_weights_buf"):
                    self._weights_buf = {}
                buf = self._weights_buf[n] = (np.empty(n, dtype=np.float64), np.empty(n, dtype=np.int32), np.empty(n, dtype=np.float64),
                                              C.c_int32(0), _lib.load().graal_candidate_weights)
            work, ids, cdf, imax, fn = buf
            sc = score if (score.dtype == np.float64 and score.flags.c_contiguous) else np.ascontiguousarray(score, dtype=np.float64)
            n_ok = fn(sc.ctypes.data, n, nt, work.ctypes.data, ids.ctypes.data, cdf.ctypes.data, C.byref(imax))
            if n_ok >= 0:
                self.sub_score = work[:n_ok].copy()
                if n_ok <= 1:
                    return int(imax.value)
                return int(ids[int(cdf[:n_ok].searchsorted(rs(), side="right"))])
        scores_2_remove = self._remove_cache.get(len(score))
        if scores_2_remove is None:
            scores_2_remove = self._remove_cache[len(score)] = list(range(nt, len(score), nt)) + list(range(nt + 1, len(score), nt))
        id_max = int(score.argmax())
        filtered_score = score - score.min()
        filtered_score[scores_2_remove] = 0
        max_score = filtered_score.max()
        thresh_overflow = 30
        filtered_score = filtered_score - (max_score - thresh_overflow)
        filtered_score[filtered_score < 0] = 0
        id_ok = np.nonzero(filtered_score > 0)[0]
        sub_score = filtered_score[id_ok]
        with np.errstate(all="ignore"):
            sub_score = sub_score / sub_score.sum()
            sub_score[sub_score > 0] = np.power(sub_score[sub_score > 0], 1. / F_t)
            sub_score = sub_score / sub_score.sum()
        self.sub_score = sub_score
        if len(id_ok) <= 1:
            return id_max
        rs = getattr(self.rng, "random_sample", None)
        if rs is None or not np.all(np.isfinite(sub_score)):
            return int(self.rng.choice(id_ok, 1, p=sub_score)[0])      # (also raises like the reference on bad weights)
        # np.random.choice(id_ok, 1, p=sub_score) of a RandomState, without its argument validation: one uniform
        # draw searched in the normalised cumulative weights (tests/test_reference_host_logic.py: same candidate,
        # same stream position as the reference's own lines)
        cdf = sub_score.cumsum()
        cdf /= cdf[-1]
        return int(id_ok[int(cdf.searchsorted(rs(), side="right"))])

    # ------------------------------------------------------------------ cuda_lib_gl.py:2022-2107
    def step_nuisance_parameters(self, dt=0, t=0, n_step=1):
        self._inc_valid = False          # the cached likelihood no longer describes the state / parameters
        curr_param = np.copy(self.param_simu)
        kuhn, lm, c1, slope, d, d_max, fact, d_nuc = curr_param[0]
        self.sigma_fact = 10 ** (np.log10(fact) - 2)
        self.sigma_slope, self.sigma_d_max, self.sigma_d_nuc = 0.05, 100, 0.5
        id_modif = self.rng.choice(4)
        # (the reference ran under NumPy 1.x, where float32 scalar + Python float is float64: cast explicitly, NumPy 2 keeps float32)
        c1f = lambda sl: rippe_c1(kuhn, lm, sl)
        if id_modif == 0:
            new_fact = np.float64(fact) + self.rng.normal(loc=0.0, scale=self.sigma_fact)
            new_d_max = opti.estimate_max_dist_intra([kuhn, lm, slope, d, new_fact], d_nuc)
            out_test_param = [(kuhn, lm, c1f(slope), slope, d, new_d_max, new_fact, d_nuc)]
        elif id_modif == 1:
            new_slope = np.float64(slope) + self.rng.normal(loc=0.0, scale=self.sigma_slope)
            new_d_max = opti.estimate_max_dist_intra([kuhn, lm, new_slope, d, fact], d_nuc)
            out_test_param = [(kuhn, lm, c1f(new_slope), new_slope, d, new_d_max, fact, d_nuc)]
        elif id_modif == 2:
            new_d_max = np.float64(d_max) + self.rng.normal(loc=0.0, scale=self.sigma_d_max)
            new_d_nuc = opti.peval(new_d_max, [kuhn, lm, slope, d, fact])
            out_test_param = [(kuhn, lm, c1f(slope), slope, d, new_d_max, fact, new_d_nuc)]
        else:
            new_d_nuc = np.float64(d_nuc) + self.rng.normal(loc=0.0, scale=self.sigma_d_nuc)
            new_d_max = opti.estimate_max_dist_intra([kuhn, lm, slope, d, fact], new_d_nuc)
            out_test_param = [(kuhn, lm, c1f(slope), slope, d, new_d_max, fact, new_d_nuc)]
        out_test_param = np.array(out_test_param, dtype=PARAM_DTYPE)
        self.param_simu_test = out_test_param                 # (gpu_param_simu_test, cuda_lib_gl.py:2088)
        test_likelihood = self.eval_likelihood(out_test_param)
        F_t = self.temperature(t, n_step)
        with np.errstate(over="ignore"):
            ratio = np.exp((test_likelihood - self.likelihood_t) / F_t)
        u = self.rng.rand()
        success = 0
        if ratio >= u:
            success = 1
            self.param_simu = out_test_param
            self._set_device_params(self.param_simu)
            self.likelihood_t = test_likelihood
        kuhn, lm, c1, slope, d, d_max, fact, d_nuc = self.param_simu[0]
        y_rippe = opti.peval(self.bins, [kuhn, lm, slope, d, fact]) if hasattr(self, "bins") else None
        return fact, d, d_max, d_nuc, slope, self.likelihood_t, success, y_rippe

    # ------------------------------------------------------------------ cuda_lib_gl.py:475-541
    def dist_inter_genome(self, tmp_gpu_vect_frags):
        tmp_gpu_vect_frags.copy_from_gpu()
        g1 = tmp_gpu_vect_frags
        return dist_inter_genome(g1.prev, g1.next, g1.ori, g1.id_d, self.np_init_prev, self.np_init_next,
                                 self.np_init_ori, self.np_init_orientable, self.id_frags_blacklisted,
                                 self.is_repeat, int(self.n_new_frags), self.n_frags_4_dist)

    def genome_content(self):
        """cuda_lib_gl.py:1625-1668: (full_order, dict_contig) -- per contig id (increasing) the data ids, positions,
        start_bp, id_c, prev, next of its bins in position order; contigs with an inactive bin stay empty."""
        self.gpu_vect_frags.copy_from_gpu()
        c = self.gpu_vect_frags
        dict_contig, full_order = dict(), []
        for k in np.unique(c.id_c):
            d = dict_contig[k] = {"id": [], "pos": [], "next": [], "prev": [], "start_bp": [], "id_c": []}
            id_pos = np.nonzero(c.id_c == k)[0]
            if np.all(c.activ[id_pos] == 1):
                o = id_pos[np.argsort(c.pos[id_pos])]
                d["id"].extend(c.id_d[o]); d["pos"].extend(c.pos[o]); d["start_bp"].extend(c.start_bp[o])
                d["id_c"].extend(c.id_c[o]); d["prev"].extend(c.prev[o]); d["next"].extend(c.next[o])
                full_order.extend(c.id_d[o])
        return full_order, dict_contig

    def display_current_matrix(self, file=None):
        """cuda_lib_gl.py:1581-1625, the ordering part: (full_order, dict_contig, full_order_high) -- the data ids of the
        bins contig by contig in position order, the same per contig, and the sub-frag ids in genome order (the sub-frags
        of a flipped bin reversed; like the reference, the orientation is looked up under the DATA id).  The reference
        also writes the reordered dense sub-level matrix to ``file`` as a TIFF (GUI snapshot, out of scope): ``file`` is
        accepted and ignored."""
        self.gpu_vect_frags.copy_from_gpu()
        c = self.gpu_vect_frags
        dict_contig, full_order, full_order_high = dict(), [], []
        for k in np.unique(c.id_c):
            dict_contig[k] = []
            id_pos = np.nonzero(c.id_c == k)[0]
            if np.all(c.activ[id_pos] == 1):
                ordered_frag = c.id_d[id_pos[np.argsort(c.pos[id_pos])]]
                dict_contig[k].extend(ordered_frag)
                full_order.extend(ordered_frag)
                for i in ordered_frag:
                    v = list(self.np_sub_frags_id[i])
                    ids = v[:v[3]]
                    if c.ori[i] == -1:
                        ids.reverse()
                    full_order_high.extend(ids)
        return full_order, dict_contig, full_order_high

    def export_new_fasta(self, level, contig_names, sequences, new_fasta, info_frags):
        """simulation.export_new_fasta (simulation_loader.py:781-783): genome.fasta + info_frags.txt of the
        current genome (graal_b200.export.generate_new_fasta)."""
        from .export import generate_new_fasta
        self.gpu_vect_frags.copy_from_gpu()
        return generate_new_fasta(self.gpu_vect_frags, level, contig_names, sequences, new_fasta, info_frags)

    def free_gpu(self):
        """cuda_lib_gl.py:2605-2613."""
        if getattr(self, "ctx", None) is not None:
            self.lib.graal_ctx_destroy(self.ctx)
            self.ctx = None
        for k in [a for a in vars(self) if a.startswith("d_")]:
            setattr(self, k, None)

    def __del__(self):
        try:
            self.free_gpu()
        except Exception:
            pass
