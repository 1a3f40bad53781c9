"""Pyramid levels and sampler inputs (host side, NumPy).

Restates, for in-memory synthetic data, the parts of the reference that feed the sampler:

* the binning rule of ``subsample_data_set`` (/root/reference/pyramid_sparse.py:406-431 bins,
  :520-557 contact aggregation) -- ``factor`` consecutive fragments of a contig form one bin, the
  last bin of a contig may be shorter, a contig with fewer than ``factor`` fragments is kept 1:1
  (``min_bin_per_contig = 1``, pyramid_sparse.py:28);
* ``level.load_data`` (pyramid_sparse.py:1261-1327 field derivation, :1350-1372 mean_value_trans);
* ``simulation.create_sub_frags`` (simulation_loader.py:673-704), ``select_repeated_frags``
  (:369-394), ``modify_vect_frags`` (:182-299), ``blacklist_contig`` (:129-163),
  ``create_new_sub_frags`` (:706-720).

The reference densifies both matrices (simulation_loader.py:81-82); here they stay sparse
(upper-triangle COO, int32 counts) and are turned into segment-sorted contact lists for the device.
The file-format quirks of the text pipeline do not apply to in-memory levels.
"""
from dataclasses import dataclass, field
import numpy as np

I32 = np.int32
F32 = np.float32
FRAG_FIELDS = ("pos", "id_c", "start_bp", "len_bp", "circ", "id", "prev", "next",
               "l_cont", "l_cont_bp", "ori", "rep", "activ", "id_d")

# S. cerevisiae S288C chromosome lengths (bp), used only as length PROPORTIONS for config C1
YEAST_CHROM_BP = (230218, 813184, 316620, 1531933, 576874, 270161, 1090940, 562643,
                  439888, 745751, 666816, 1078177, 924431, 784333, 1091291, 948066)


@dataclass
class PyramidLevel:
    """One level of the pyramid: fragment table + upper-triangle contact COO."""
    level: int
    contig_id: np.ndarray      # 1-based contig of each bin
    start_pos: np.ndarray      # bp offset of the bin inside its contig
    end_pos: np.ndarray
    n_accu: np.ndarray         # number of level-0 fragments in the bin
    sub_low: np.ndarray        # first / last index (0-based) of the bin's members in the level below
    sub_high: np.ndarray
    rows: np.ndarray           # COO, rows <= cols (diagonal kept, as the reference does)
    cols: np.ndarray
    vals: np.ndarray
    S_o_A_frags: dict = field(default_factory=dict)
    mean_value_trans: float = 0.0

    @property
    def n_frags(self):
        return int(self.contig_id.shape[0])

    @property
    def size(self):
        return (self.end_pos - self.start_pos).astype(I32)


def _derive_frag_arrays(lv):
    """level.load_data (pyramid_sparse.py:1261-1327): SoA fragment state of the initial genome."""
    n = lv.n_frags
    cid = lv.contig_id
    first = np.r_[True, cid[1:] != cid[:-1]]
    last = np.r_[cid[1:] != cid[:-1], True]
    starts = np.nonzero(first)[0]
    seg = np.cumsum(first) - 1
    pos = np.arange(n) - starts[seg]
    size = lv.size.astype(np.int64)
    l_cont = np.bincount(seg)[seg]
    l_cont_bp = np.bincount(seg, weights=size).astype(np.int64)[seg]
    ids = np.arange(n)
    s = {
        "pos": pos, "id_c": cid, "start_bp": lv.start_pos, "len_bp": size, "circ": np.zeros(n),
        "id": ids, "prev": np.where(first, -1, ids - 1), "next": np.where(last, -1, ids + 1),
        "l_cont": l_cont, "l_cont_bp": l_cont_bp, "n_accu": lv.n_accu,
    }
    lv.S_o_A_frags = {k: np.asarray(v, dtype=I32) for k, v in s.items()}


def _mean_value_trans(lv):
    """pyramid_sparse.py:1350-1372: sum of (upper-triangle) contacts leaving each contig's rows over
    the number of row x column cells outside the contig's diagonal block."""
    n = lv.n_frags
    cid = lv.contig_id
    trans = cid[lv.rows] != cid[lv.cols]
    total_trans = float(lv.vals[trans].sum())
    sizes = np.bincount(cid)[1:].astype(np.float64)
    sizes = sizes[sizes > 0]
    n_tot = float((sizes * n - sizes * sizes).sum())
    return total_trans / F32(n_tot) if n_tot > 0 else 0.0


def _coarsen(prev, factor, level):
    """subsample_data_set (pyramid_sparse.py:406-431, 520-557)."""
    cid = prev.contig_id
    n = prev.n_frags
    first = np.r_[True, cid[1:] != cid[:-1]]
    starts = np.nonzero(first)[0]
    seg = np.cumsum(first) - 1
    rel = np.arange(n) - starts[seg]
    n_in_contig = np.bincount(seg)[seg]
    binned = (n_in_contig / F32(factor) >= 1) & (factor != 1)
    new_start = np.where(binned, rel % factor == 0, True)          # id_frag_rel % fact == 1, 1-based
    old2new = np.cumsum(new_start) - 1
    m = int(old2new[-1]) + 1
    lo = np.nonzero(new_start)[0]
    hi = np.r_[lo[1:] - 1, n - 1]
    # contacts: sum counts per (min, max) bin pair, diagonal kept
    fa, fb = old2new[prev.rows], old2new[prev.cols]
    f1, f2 = np.minimum(fa, fb), np.maximum(fa, fb)
    key = f1.astype(np.int64) * m + f2
    uk, inv = np.unique(key, return_inverse=True)
    vals = np.bincount(inv, weights=prev.vals.astype(np.float64)).astype(np.int64)
    lv = PyramidLevel(level=level, contig_id=cid[lo].astype(I32), start_pos=prev.start_pos[lo].astype(I32),
                      end_pos=prev.end_pos[hi].astype(I32),
                      n_accu=(np.add.reduceat(prev.n_accu.astype(np.int64), lo)).astype(I32),
                      sub_low=lo.astype(I32), sub_high=hi.astype(I32),
                      rows=(uk // m).astype(I32), cols=(uk % m).astype(I32), vals=vals.astype(I32))
    _derive_frag_arrays(lv)
    lv.mean_value_trans = _mean_value_trans(lv)
    return lv


def rippe_law(s_kb, kuhn=1.0, lm=9.6, slope=-1.5, d=3.0):
    """Shape of the Rippe contact law (optim_rippe_curve_update.py:22-28 with A = 1), float64."""
    x = lm * np.asarray(s_kb, dtype=np.float64) / kuhn
    return 0.53 * kuhn ** -3.0 * np.power(x, slope) * np.exp((d - 2) / (x * x + d))


@dataclass
class Pyramid:
    levels: list
    factor: int
    spec: dict

    def get_level(self, k):
        return self.levels[k]


def build_synthetic_pyramid(contig_bp, n_frags0, n_levels, factor=3, seed=20141217,
                            cis_rowsum=500.0, v_inter=0.02, min_frag_bp=50, max_band=4000):
    """Synthetic pyramid: ``len(contig_bp)`` contigs cut into ``n_frags0`` level-0
    fragments at uniform random sites (>= min_frag_bp); per level-0 pair a Poisson count with mean
    max(A*rippe(s), v_inter) in cis and v_inter in trans, A set so the mean cis row-sum is
    ``cis_rowsum``; then ``n_levels - 1`` coarsenings by ``factor``."""
    rs = np.random.RandomState(seed)
    contig_bp = np.asarray(contig_bp, dtype=np.int64)
    nc = contig_bp.shape[0]
    share = np.maximum(1, np.round(n_frags0 * contig_bp / contig_bp.sum()).astype(np.int64))
    share[np.argmax(share)] += n_frags0 - share.sum()
    cid, st, en = [], [], []
    for c in range(nc):
        k, L = int(share[c]), int(contig_bp[c])
        k = max(1, min(k, L // min_frag_bp))
        cuts = np.sort(rs.choice(np.arange(1, max(2, L // min_frag_bp)), size=k - 1, replace=False)) * min_frag_bp \
            if k > 1 else np.zeros(0, dtype=np.int64)
        b = np.r_[0, cuts, L]
        cid.append(np.full(k, c + 1)); st.append(b[:-1]); en.append(b[1:])
    cid = np.concatenate(cid).astype(I32); st = np.concatenate(st).astype(I32); en = np.concatenate(en).astype(I32)
    n = cid.shape[0]
    mid_kb = (st + en) / 2000.0
    first = np.r_[True, cid[1:] != cid[:-1]]
    starts = np.nonzero(first)[0]
    ends = np.r_[starts[1:], n]
    # amplitude: mean over fragments of sum_j A*rippe(s_ij) = cis_rowsum (both directions)
    rs2 = np.random.RandomState(seed + 1)
    mean_len_kb = float((en - st).mean()) / 1000.0
    grid = np.arange(1, max_band + 1) * mean_len_kb
    shape = rippe_law(grid)
    A = cis_rowsum / (2.0 * shape.sum())
    above = np.nonzero(A * shape > v_inter)[0]
    band_k = int(min(max_band, (above[-1] + 1) * 2 + 8)) if above.size else 8
    rows, cols, vals = [], [], []
    for a, b in zip(starts, ends):
        m = b - a
        x = mid_kb[a:b]
        for k in range(1, min(band_k, m - 1) + 1):
            lam = A * rippe_law(x[k:] - x[:-k]) - v_inter
            lam = np.maximum(lam, 0.0)
            if lam.max() <= 0:
                continue
            cnt = rs2.poisson(lam)
            nz = np.nonzero(cnt)[0]
            rows.append(a + nz); cols.append(a + nz + k); vals.append(cnt[nz])
    # uniform background at rate v_inter over all unordered pairs (cis and trans alike)
    n_pairs = n * (n - 1) // 2
    nb = rs2.poisson(v_inter * n_pairs)
    i = rs2.randint(0, n, size=nb); j = rs2.randint(0, n - 1, size=nb)
    j = np.where(j >= i, j + 1, j)
    rows.append(np.minimum(i, j)); cols.append(np.maximum(i, j)); vals.append(np.ones(nb, dtype=np.int64))
    rows = np.concatenate(rows).astype(np.int64); cols = np.concatenate(cols).astype(np.int64)
    vals = np.concatenate(vals).astype(np.int64)
    key = rows * n + cols
    uk, inv = np.unique(key, return_inverse=True)
    v = np.bincount(inv, weights=vals.astype(np.float64)).astype(np.int64)
    lv0 = PyramidLevel(level=0, contig_id=cid, start_pos=st, end_pos=en, n_accu=np.ones(n, dtype=I32),
                       sub_low=np.arange(n, dtype=I32), sub_high=np.arange(n, dtype=I32),
                       rows=(uk // n).astype(I32), cols=(uk % n).astype(I32), vals=v.astype(I32))
    _derive_frag_arrays(lv0)
    lv0.mean_value_trans = _mean_value_trans(lv0)
    levels = [lv0]
    for k in range(1, n_levels):
        levels.append(_coarsen(levels[-1], factor, k))
    spec = dict(seed=seed, cis_rowsum=cis_rowsum, v_inter=v_inter, amplitude=A, band_frags=band_k,
                law=dict(kuhn=1.0, lm=9.6, slope=-1.5, d=3.0))
    return Pyramid(levels=levels, factor=factor, spec=spec)


def yeast_shaped_pyramid(n_frags0=5000, total_bp=12_000_000, n_levels=4, **kw):
    """BASELINE config C1."""
    prop = np.array(YEAST_CHROM_BP, dtype=np.float64)
    return build_synthetic_pyramid(np.round(prop / prop.sum() * total_bp), n_frags0, n_levels, **kw)


def treesei_shaped_pyramid(n_frags0=100_000, total_bp=33_000_000, n_contigs=77, n_levels=6,
                           seed=20141217, cis_rowsum=400.0, v_inter=0.004, **kw):
    """BASELINE config C2: 77 contigs with log-normal lengths; ~30 M distinct level-0 contact
    entries (240 MB of contact lists: larger than the 126 MB L2)."""
    rs = np.random.RandomState(seed + 7)
    w = rs.lognormal(mean=0.0, sigma=1.0, size=n_contigs)
    bp = np.maximum(20_000, np.round(w / w.sum() * total_bp))
    return build_synthetic_pyramid(bp, n_frags0, n_levels, seed=seed, cis_rowsum=cis_rowsum,
                                   v_inter=v_inter, **kw)


# ----------------------------------------------------------------------------------------------
# sampler inputs (what simulation.__init__ hands to sampler.__init__, minus the GL objects)
# ----------------------------------------------------------------------------------------------
@dataclass
class SamplerInputs:
    S_o_A_frags: dict                 # 14 int32 arrays (+ n_accu) of length n_new_frags
    collector_id_repeats: np.ndarray  # int32
    frag_dispatcher: np.ndarray       # int32 (n_frags, 2)
    id_frag_duplicated: np.ndarray    # data ids with copies
    id_frags_blacklisted: list
    n_frags: int                      # data bins (reference: init_n_frags)
    n_new_frags: int                  # bins + copies
    init_n_sub_frags: int             # W
    n_new_sub_frags: int
    np_rep_sub_frags_id: np.ndarray   # (n_new_frags, 4)
    np_sub_frags_len_bp: np.ndarray   # (n_frags, 3) float32, kb
    np_sub_frags_id: np.ndarray       # (n_frags, 4) int32: x, y, z, count
    np_sub_frags_accu: np.ndarray     # (n_frags, 3) int32
    mean_squared_frags_per_bin: np.float32
    norm_vect_accu: np.ndarray
    S_o_A_sub_frags: dict
    level_coo: tuple                  # (rows, cols, vals) upper triangle, current level
    sub_coo: tuple                    # (rows, cols, vals) upper triangle, sub level
    mean_value_trans: float

    def dense_sub_matrix(self):
        """The reference's ``hic_matrix`` before the sampler zeroes its diagonal
        (simulation_loader.py:82): csr + csr.T, float32.  Small levels only."""
        return _dense_sym(self.sub_coo, self.init_n_sub_frags)

    def dense_level_matrix(self):
        return _dense_sym(self.level_coo, self.n_frags)


def _dense_sym(coo, n):
    r, c, v = coo
    m = np.zeros((n, n), dtype=F32)
    np.add.at(m, (r, c), v.astype(F32))
    np.add.at(m, (c, r), v.astype(F32))
    return m


def prepare_sampler_inputs(pyr, level, allow_repeats=False, blacklist_contigs=()):
    """simulation.__init__ data prep (simulation_loader.py:64-107) for ``level`` (>= 1)."""
    if level < 1:
        raise ValueError("the likelihood is evaluated on level-1 sub-fragments: level must be >= 1")
    lv, sub = pyr.get_level(level), pyr.get_level(level - 1)
    N = lv.n_frags
    # create_sub_frags (:673-704)
    n_sub = (lv.sub_high - lv.sub_low + 1).astype(I32)
    if n_sub.max() > 3:
        raise ValueError("a bin holds more than 3 sub-fragments: the device structs are 3-wide (factor must be <= 3)")
    sub_id = np.zeros((N, 4), dtype=I32); sub_len = np.zeros((N, 3), dtype=F32); sub_accu = np.zeros((N, 3), dtype=I32)
    sub_id[:, 3] = n_sub
    sub_size = sub.S_o_A_frags["len_bp"]
    for k in range(3):
        ok = k < n_sub
        idx = np.where(ok, lv.sub_low + k, 0)
        sub_id[:, k] = np.where(ok, idx, 0)
        sub_len[:, k] = np.where(ok, sub_size[idx].astype(F32) / F32(1000.0), F32(0))
        sub_accu[:, k] = np.where(ok, sub.n_accu[idx], 0)
    collect_accu = sub.n_accu.astype(F32)          # every sub-fragment belongs to exactly one bin, in index order
    norm_vect = sub_accu.sum(axis=1)
    msfpb = F32(collect_accu.mean() ** 2)
    W = int(n_sub.sum())
    # select_repeated_frags (:369-394)
    cov = np.bincount(lv.rows, weights=lv.vals, minlength=N) + np.bincount(lv.cols, weights=lv.vals, minlength=N)
    thr = cov.mean() + 3 * cov.std()
    dup = np.nonzero(cov > thr)[0] if allow_repeats else np.zeros(0, dtype=np.int64)
    n_dup = [int(max(1, np.round(cov[e] / thr) - 1)) for e in dup]
    # modify_vect_frags (:182-299)
    base = lv.S_o_A_frags
    s = {k: list(base[k]) for k in ("pos", "id_c", "start_bp", "len_bp", "circ", "id", "prev", "next",
                                    "l_cont", "l_cont_bp", "n_accu")}
    s["rep"] = [0] * N; s["activ"] = [1] * N; s["id_d"] = list(base["id"])
    max_f, max_c = N, int(base["id_c"].max()) + 1
    for e, k in zip(dup, n_dup):
        for _ in range(k):
            s["pos"].append(0); s["id_c"].append(max_c); s["start_bp"].append(0)
            s["len_bp"].append(base["len_bp"][e]); s["circ"].append(base["circ"][e]); s["id"].append(max_f)
            s["prev"].append(-1); s["next"].append(-1); s["l_cont"].append(1)
            s["l_cont_bp"].append(base["len_bp"][e]); s["n_accu"].append(base["n_accu"][e])
            s["rep"].append(1); s["activ"].append(1); s["id_d"].append(base["id"][e])
            max_f += 1; max_c += 1
    s = {k: np.array(v, dtype=I32) for k, v in s.items()}
    s["ori"] = np.ones(max_f, dtype=I32)              # Q5: initial orientation is +1 whatever the loader says
    dupset = set(int(e) for e in dup)
    coll, disp, x = [], [], 0
    if dupset:
        for f in range(N):
            if f in dupset:
                ids = np.nonzero(s["id_d"] == f)[0]
                coll.extend(int(i) for i in ids); disp.append((x, x + len(ids))); x += len(ids)
            else:
                coll.append(f); disp.append((x, x + 1)); x += 1
        coll = np.array(coll, dtype=I32); disp = np.array(disp, dtype=I32)
    else:
        coll = np.arange(N, dtype=I32)
        disp = np.stack([np.arange(N), np.arange(N) + 1], axis=1).astype(I32)
    # blacklist_contig (:129-163)
    black = []
    for c in blacklist_contigs:
        for f in np.nonzero(base["id_c"] == c)[0]:
            black.extend(int(i) for i in coll[disp[f, 0]:disp[f, 1]])
    # create_new_sub_frags (:706-720)
    cnt = sub_id[s["id_d"], 3]
    offs = np.r_[0, np.cumsum(cnt)[:-1]]
    rep_sub = np.zeros((max_f, 4), dtype=I32)
    rep_sub[:, 3] = cnt
    for k in range(3):
        rep_sub[:, k] = np.where(k < cnt, offs + k, 0)
    return SamplerInputs(
        S_o_A_frags=s, collector_id_repeats=coll, frag_dispatcher=disp,
        id_frag_duplicated=np.asarray(dup, dtype=I32), id_frags_blacklisted=black,
        n_frags=N, n_new_frags=max_f, init_n_sub_frags=W, n_new_sub_frags=int(cnt.sum()),
        np_rep_sub_frags_id=rep_sub, np_sub_frags_len_bp=sub_len, np_sub_frags_id=sub_id,
        np_sub_frags_accu=sub_accu, mean_squared_frags_per_bin=msfpb, norm_vect_accu=norm_vect,
        S_o_A_sub_frags=sub.S_o_A_frags, level_coo=(lv.rows, lv.cols, lv.vals),
        sub_coo=(sub.rows, sub.cols, sub.vals), mean_value_trans=float(sub.mean_value_trans))


# ----------------------------------------------------------------------------------------------
# BASELINE config C4 / C5: one large level generated on the GPU
# ----------------------------------------------------------------------------------------------
def synthetic_roofline_level(n_bins=200_000, n_contigs=24, n_draws=240_000_000, d_max_kb=1000.0, trans_frac=0.05,
                             seed=20141217, device="cuda"):
    """One level of ``n_bins`` bins (3 sub-frags each, uniform accu) in ``n_contigs`` contigs with roughly
    ``n_draws`` stored contacts: the row is uniform, a cis partner sits at a sub-frag offset drawn from
    p(k) ~ k^-1.5 truncated at the band (d_max), 5 % of the draws are uniform trans pairs, counts are
    1 + Poisson(0.5), duplicate pairs are summed.  Everything is generated with torch on ``device``.
    Returns (SamplerInputs with sub_coo = None, (rowptr, contacts) device tensors, (xk, pk) proposal tables)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rs = np.random.RandomState(seed)
    w = rs.lognormal(0.0, 0.6, size=n_contigs)
    sizes = np.maximum(3, np.round(w / w.sum() * n_bins).astype(np.int64))
    sizes[np.argmax(sizes)] += n_bins - sizes.sum()
    cid = np.repeat(np.arange(1, n_contigs + 1), sizes).astype(I32)
    N, W = n_bins, 3 * n_bins
    sub_len_bp = rs.randint(500, 1500, size=W).astype(np.int64)
    len_bp = sub_len_bp.reshape(N, 3).sum(axis=1)
    first = np.r_[True, cid[1:] != cid[:-1]]
    starts = np.nonzero(first)[0]
    seg = np.cumsum(first) - 1
    csum = np.cumsum(len_bp) - len_bp
    start_bp = csum - csum[starts][seg]
    ids = np.arange(N)
    last = np.r_[cid[1:] != cid[:-1], True]
    soa = {"pos": ids - starts[seg], "id_c": cid, "start_bp": start_bp, "len_bp": len_bp, "circ": np.zeros(N),
           "id": ids, "prev": np.where(first, -1, ids - 1), "next": np.where(last, -1, ids + 1),
           "l_cont": np.bincount(seg)[seg], "l_cont_bp": np.bincount(seg, weights=len_bp).astype(np.int64)[seg],
           "n_accu": np.full(N, 3), "ori": np.ones(N), "rep": np.zeros(N), "activ": np.ones(N), "id_d": ids}
    soa = {k: np.asarray(v, dtype=I32) for k, v in soa.items()}
    sub_id = np.zeros((N, 4), dtype=I32)
    sub_id[:, 0] = 3 * ids; sub_id[:, 1] = 3 * ids + 1; sub_id[:, 2] = 3 * ids + 2; sub_id[:, 3] = 3
    sub_len = (sub_len_bp.astype(F32) / F32(1000.0)).reshape(N, 3)
    sub_accu = np.ones((N, 3), dtype=I32)
    # sub-level SoA (initial layout) for the distance histogram
    sub_cid = np.repeat(cid, 3)
    sfirst = np.r_[True, sub_cid[1:] != sub_cid[:-1]]
    sstarts = np.nonzero(sfirst)[0]
    sseg = np.cumsum(sfirst) - 1
    scs = np.cumsum(sub_len_bp) - sub_len_bp
    sub_soa = {"id_c": sub_cid.astype(I32), "start_bp": (scs - scs[sstarts][sseg]).astype(I32),
               "len_bp": sub_len_bp.astype(I32), "pos": (np.arange(W) - sstarts[sseg]).astype(I32)}
    # ---- contacts on the device: per row, offset k is present with probability min(1, lam * k^-1.5)
    # (a thinned power law: the near diagonal is full, the tail sparse); lam is set so that the expected
    # number of cis entries per row is (1 - trans_frac) * n_draws / W
    t_cid = torch.from_numpy(sub_cid.astype(np.int64)).to(device)
    mean_sub_kb = float(sub_len_bp.mean()) / 1000.0
    kmax = max(2, int(d_max_kb / mean_sub_kb))
    target = (1.0 - trans_frac) * n_draws / W
    kk = np.arange(1, kmax + 1, dtype=np.float64)
    lo_l, hi_l = 1e-3, 1e9
    for _ in range(200):
        lam = np.sqrt(lo_l * hi_l)
        if np.minimum(1.0, lam * kk ** -1.5).sum() < target:
            lo_l = lam
        else:
            hi_l = lam
    pk_dev = torch.from_numpy(np.minimum(1.0, lam * kk ** -1.5).astype(np.float32)).to(device)
    keys = []
    rows_per_chunk = max(1, 40_000_000 // kmax)
    for r0 in range(0, W, rows_per_chunk):
        r1 = min(W, r0 + rows_per_chunk)
        hit = torch.rand((r1 - r0, kmax), device=device, generator=g) < pk_dev[None, :]
        rr, ko = torch.nonzero(hit, as_tuple=True)
        rr = rr + r0
        cc = rr + ko + 1
        ok = cc < W
        rr, cc = rr[ok], cc[ok]
        ok = t_cid[rr] == t_cid[cc]
        keys.append(rr[ok] * W + cc[ok])
        del hit, rr, ko, cc, ok
    n_trans_draws = int(trans_frac * n_draws)
    r = torch.randint(0, W, (n_trans_draws,), device=device, generator=g)
    c = torch.randint(0, W, (n_trans_draws,), device=device, generator=g)
    ok = t_cid[r] != t_cid[c]
    r, c = r[ok], c[ok]
    keys.append(torch.minimum(r, c) * W + torch.maximum(r, c))
    del r, c, ok
    keys = torch.cat(keys)
    keys, _ = torch.sort(keys)
    uk, cnt = torch.unique_consecutive(keys, return_counts=True)
    del keys
    extra = torch.poisson(torch.full((uk.numel(),), 0.5, device=device), generator=g)
    vals = (cnt.to(torch.float32) + extra)
    rows = torch.div(uk, W, rounding_mode="floor")
    cols = (uk - rows * W).to(torch.int32)
    rowptr = torch.zeros(W + 1, dtype=torch.int64, device=device)
    rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=W), 0)
    contacts = torch.stack([cols, vals.view(torch.int32)], dim=1).contiguous()
    n_trans = int((t_cid[rows] != t_cid[cols.to(torch.int64)]).sum().item())
    sizes_f = sizes.astype(np.float64) * 3
    mvt = float(vals[t_cid[rows] != t_cid[cols.to(torch.int64)]].sum().item()) / float((sizes_f * W - sizes_f * sizes_f).sum())
    del rows, cols, vals, uk, cnt, extra
    # proposal tables: the ten nearest bins in the initial layout, uniform weights
    offs = np.array([1, -1, 2, -2, 3, -3, 4, -4, 5, -5])
    xk = np.clip(ids[:, None] + offs[None, :], 0, N - 1).astype(I32)
    pk = np.full((N, 10), F32(0.1), dtype=F32)
    inp = SamplerInputs(
        S_o_A_frags=soa, collector_id_repeats=np.arange(N, dtype=I32),
        frag_dispatcher=np.stack([ids, ids + 1], axis=1).astype(I32), id_frag_duplicated=np.zeros(0, dtype=I32),
        id_frags_blacklisted=[], n_frags=N, n_new_frags=N, init_n_sub_frags=W, n_new_sub_frags=W,
        np_rep_sub_frags_id=sub_id.copy(), np_sub_frags_len_bp=sub_len, np_sub_frags_id=sub_id, np_sub_frags_accu=sub_accu,
        mean_squared_frags_per_bin=F32(1.0), norm_vect_accu=sub_accu.sum(axis=1), S_o_A_sub_frags=sub_soa,
        level_coo=None, sub_coo=None, mean_value_trans=mvt)
    return inp, (rowptr, contacts), (xk, pk), dict(kmax=kmax, n_trans_entries=n_trans, d_max_kb=d_max_kb)
