"""NumPy restatement of GRAAL's DENSE likelihood kernels (TEST INFRASTRUCTURE).

Follows /root/reference/kernels3.cu:
  factorial                   :80-93
  rippe_contacts              :120-133
  rippe_contacts_circ         :135-166
  evaluate_likelihood_double  :191-210
  lin_2_2dpos / conv_plan_pos_2_lin :214-235 (exact integer form; the reference's float32 sqrt
                                     decode is wrong for N >= 4609)
  evaluate_likelihood         :2802-3222   (full, one value per bin-pair pixel)
  sub_compute_likelihood      :3259-3718   (delta vs the cached per-pixel values)
in the reference's arithmetic: float32 expected values (float32 sums over replicate pairs in
the kernel's loop order), float64 Poisson terms and accumulation.

Inputs are the arrays the reference sampler constructor receives (cuda_lib_gl.py:33-42):
the sub-level matrix is DENSE float32 (W x W), symmetric, diagonal zeroed.
"""
import numpy as np

F32 = np.float32
I32 = np.int32
PARAM_FIELDS = ("kuhn", "lm", "c1", "slope", "d", "d_max", "fact", "v_inter")


def make_params(kuhn, lm, slope, d, fact, d_max, v_inter):
    """setup_rippe_parameters (cuda_lib_gl.py:1203-1214): c1 is rounded to float32 once."""
    kuhn, lm = F32(kuhn), F32(lm)
    # NumPy 1.x promotion, as the reference ran: float32 ** float32 stays float32 (nuisance step,
    # slope read back from the float32 record), float32 ** float64/python float is float64 (fit)
    ratio = lm / kuhn
    pw = np.power(ratio, slope) if isinstance(slope, np.float32) else np.power(np.float64(ratio), np.float64(slope))
    c1 = F32((0.53 * np.float64(pw)) * np.float64(np.power(kuhn, F32(-3))))
    return dict(kuhn=kuhn, lm=lm, c1=c1, slope=F32(slope), d=F32(d), d_max=F32(d_max),
                fact=F32(fact), v_inter=F32(v_inter))


def params_to_array(p):
    return np.array([p[k] for k in PARAM_FIELDS], dtype=F32)


def factorial_f32(n):
    """kernels3.cu:80-93 on a float32 array."""
    n = np.floor(np.asarray(n, dtype=F32))
    out = np.ones_like(n, dtype=F32)
    small = n < 10
    for c in range(1, 10):
        out = np.where(small & (c <= n), out * F32(c), out).astype(F32)
    big = ~small
    if np.any(big):
        nb = n[big]
        with np.errstate(over="ignore"):
            v = np.power(nb, nb).astype(F32) * np.exp(-nb).astype(F32) \
                * np.sqrt((2 * np.pi * nb.astype(np.float64)).astype(F32)).astype(F32)
        out[big] = v.astype(F32)
    return out


def rippe_contacts(s, p):
    """kernels3.cu:120-133, float32 in / float32 out."""
    s = np.asarray(s, dtype=F32)
    res = np.zeros_like(s, dtype=F32)
    m = (s > 0) & (s < p["d_max"])
    if np.any(m):
        sm = s[m]
        x = sm * p["lm"] / p["kuhn"]
        e = np.exp((p["d"] - F32(2)) / (np.power(x, F32(2.0)) + p["d"]))
        res[m] = (p["c1"] * np.power(sm, p["slope"]) * e) * p["fact"]
    return np.fmax(res, p["v_inter"]).astype(F32)


def rippe_contacts_circ(s, s_tot, p):
    """kernels3.cu:135-166 (CUDA max(float,float) drops NaN -> np.fmax)."""
    s = np.asarray(s, dtype=F32)
    s_tot = np.broadcast_to(np.asarray(s_tot, dtype=F32), s.shape)
    res = np.zeros_like(s, dtype=F32)
    m = (s > 0) & (s < p["d_max"])
    if np.any(m):
        sm, st = s[m], s_tot[m]
        K = p["lm"] / p["kuhn"]
        nmax = K * F32(1)
        with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
            n = K * sm * (st - sm) / st
            norm_lin = rippe_contacts(sm, p)
            k3 = np.power(p["kuhn"], F32(-3.0))
            dm2 = p["d"] - F32(2.0)
            norm_circ = (k3 * np.power(nmax, p["slope"])
                         * np.exp(dm2 / (np.power(nmax, F32(2.0)) + p["d"]))) * p["fact"]
            val = (k3 * np.power(n, p["slope"]) * np.exp(dm2 / (np.power(n, F32(2.0)) + p["d"]))) * p["fact"]
            res[m] = (val * norm_lin / norm_circ).astype(F32)
    return np.fmax(res, p["v_inter"]).astype(F32)


def log_factorial_term(ob):
    """The observation-only part of evaluate_likelihood_double (kernels3.cu:198-203), float64."""
    ob = np.asarray(ob, dtype=np.float64)
    out = np.zeros_like(ob)
    hi = ob >= 15
    if np.any(hi):
        o = ob[hi]
        out[hi] = o * np.log(o) - o + np.log(np.sqrt(o * 2.0 * np.pi))
    lo = (ob > 0) & (ob < 15)
    if np.any(lo):
        out[lo] = np.log(factorial_f32(ob[lo].astype(F32)).astype(np.float64))
    return out


def evaluate_likelihood_double(ex, ob):
    """kernels3.cu:191-210 on float64 arrays."""
    ex = np.asarray(ex, dtype=np.float64)
    ob = np.asarray(ob, dtype=np.float64)
    res = np.zeros_like(ex)
    nz = ex != 0
    pos = nz & (ob > 0)
    if np.any(pos):
        res[pos] = ob[pos] * np.log(ex[pos]) - ex[pos] - log_factorial_term(ob[pos])
    zero = nz & (ob == 0)
    res[zero] = -ex[zero]
    return res


def pix_index(a, b):
    """conv_plan_pos_2_lin (kernels3.cu:226-235) for a < b: b(b-1)/2 + a, in int64."""
    a = np.asarray(a, dtype=np.int64)
    b = np.asarray(b, dtype=np.int64)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    return hi * (hi - 1) // 2 + lo


def pix_decode(idx):
    """lin_2_2dpos (kernels3.cu:214-224) in exact integer arithmetic: idx -> (a, b), a < b."""
    idx = np.asarray(idx, dtype=np.int64)
    b = ((1 + np.sqrt(1.0 + 8.0 * idx.astype(np.float64))) / 2).astype(np.int64)
    b = np.where(b * (b - 1) // 2 > idx, b - 1, b)
    b = np.where((b + 1) * b // 2 <= idx, b + 1, b)
    a = idx - b * (b - 1) // 2
    return a, b


class DenseLevel:
    """Read-only inputs of the likelihood kernels, as passed to the reference sampler."""

    def __init__(self, n_frags, sub_id, sub_len_kb, sub_accu, collector, dispatcher, obs, nfpb):
        self.n_frags = int(n_frags)
        self.sub_id = np.asarray(sub_id, dtype=I32).reshape(-1, 4)        # x,y,z,count
        self.sub_len = np.asarray(sub_len_kb, dtype=F32).reshape(-1, 3)   # kb
        self.sub_accu = np.asarray(sub_accu, dtype=I32).reshape(-1, 3)
        self.collector = np.asarray(collector, dtype=I32)
        self.dispatcher = np.asarray(dispatcher, dtype=I32).reshape(-1, 2)
        self.obs = obs                                                    # dense float32 W x W
        self.nfpb = F32(nfpb)
        self.limit = self.sub_id[:, 3] - 1
        self.max_rep = int((self.dispatcher[:, 1] - self.dispatcher[:, 0]).max())


def frag_geometry(slot, lv):
    """Sub-frag mid-points (kb, float32) of every frag copy, indexed by LOCAL sub position,
    in the op order of kernels3.cu:2997-3060 / 3491-3546.  Pure function of the frag's own fields."""
    id_d = slot["id_d"]
    ln = lv.sub_len[id_d]                      # (n,3)
    lim = lv.limit[id_d]
    n = id_d.shape[0]
    start_kb = slot["start_bp"].astype(F32) / F32(1000.0)
    ori = slot["ori"]
    mid = np.zeros((n, 3), dtype=F32)
    rows = np.arange(n)
    # list order: i -> local position (i if ori == 1 else limit - i)
    l0 = np.where(ori == 1, 0, lim)
    len0 = ln[rows, l0]
    acc = (start_kb + len0).astype(F32)
    mid[rows, l0] = (start_kb + len0 / F32(2.0)).astype(F32)
    for i in (1, 2):
        valid = i <= lim
        li = np.where(ori == 1, i, lim - i)
        li = np.where(valid, li, 0)
        leni = ln[rows, li]
        m = (acc + leni / F32(2.0)).astype(F32)
        r = rows[valid]
        mid[r, li[valid]] = m[valid]
        acc = np.where(valid, (acc + leni).astype(F32), acc).astype(F32)
    return mid


def _pair_expected(slot, lv, p, mid, fi, fj, o_bi, o_bj):
    """Expected 3x3 block (float32) of ONE replicate pair (fi, fj) for every pixel, indexed
    [pixel, local sub of the pixel's i bin, local sub of the pixel's j bin].
    kernels3.cu:2939-3204 (full) == :3433-3685 (delta)."""
    P = fi.shape[0]
    di, dj = slot["id_d"][fi], slot["id_d"][fj]
    acc_i = lv.sub_accu[di].astype(np.int64)         # (P,3) true accus by local pos
    acc_j = lv.sub_accu[dj].astype(np.int64)
    cis = slot["id_c"][fi] == slot["id_c"][fj]
    out = np.zeros((P, 3, 3), dtype=F32)
    # ---- trans branch (:3101-3204): quirk Q1 on the i side when ori == -1
    tr = ~cis
    if np.any(tr):
        lim_i = lv.limit[di]
        acc_iq = np.where((slot["ori"][fi] == -1)[:, None], acc_i[np.arange(P), lim_i][:, None], acc_i)
        prod = (acc_iq[tr][:, :, None] * acc_j[tr][:, None, :]).astype(I32)
        norm = prod.astype(F32) / lv.nfpb
        out[tr] = (p["v_inter"] * norm).astype(F32)
    # ---- cis branch (:2939-3100)
    if np.any(cis):
        c = np.nonzero(cis)[0]
        a_, b_ = fi[c], fj[c]
        swap = slot["pos"][a_] > slot["pos"][b_]
        first = np.where(swap, b_, a_)             # the frag closest to the contig origin
        s = np.abs(mid[b_][:, None, :] - mid[a_][:, :, None]).astype(F32)     # [c, sub_i, sub_j]
        prod = (acc_i[c][:, :, None] * acc_j[c][:, None, :]).astype(I32)
        norm = prod.astype(F32) / lv.nfpb
        circ = (slot["circ"][first] == 1)
        s_tot = slot["l_cont_bp"][first].astype(F32) / F32(1000.0)
        r = rippe_contacts(s, p)
        if np.any(circ):
            rc = rippe_contacts_circ(s[circ], s_tot[circ][:, None, None], p)
            r[circ] = rc
        out[c] = (r * norm).astype(F32)
    return out


def pixel_loglik(slot, lv, p, bi, bj, on_diag, mid=None, chunk=200000):
    """Log-likelihood of the pixels (bi[k], bj[k]) (data-bin ids, bi <= bj): the body of the
    pixel loop shared by evaluate_likelihood and sub_compute_likelihood."""
    bi = np.asarray(bi, dtype=np.int64)
    bj = np.asarray(bj, dtype=np.int64)
    on_diag = np.broadcast_to(np.asarray(on_diag, dtype=bool), bi.shape)
    if mid is None:
        mid = frag_geometry(slot, lv)
    out = np.zeros(bi.shape[0], dtype=np.float64)
    for lo in range(0, bi.shape[0], chunk):
        sl = slice(lo, min(lo + chunk, bi.shape[0]))
        out[sl] = _pixel_loglik_chunk(slot, lv, p, bi[sl], bj[sl], on_diag[sl], mid)
    return out


def _pixel_loglik_chunk(slot, lv, p, bi, bj, on_diag, mid):
    P = bi.shape[0]
    d_i, d_j = lv.dispatcher[bi], lv.dispatcher[bj]
    exp = np.zeros((P, 3, 3), dtype=F32)
    activ = slot["activ"]
    for ci in range(lv.max_rep):
        idx_i = d_i[:, 0] + ci
        ok_i = idx_i < d_i[:, 1]
        fi = lv.collector[np.where(ok_i, idx_i, d_i[:, 0])]
        ok_i &= activ[fi] == 1
        for cj in range(lv.max_rep):
            idx_j = d_j[:, 0] + cj
            ok_j = idx_j < d_j[:, 1]
            fj = lv.collector[np.where(ok_j, idx_j, d_j[:, 0])]
            ok = ok_i & ok_j & (activ[fj] == 1)
            if not np.any(ok):
                continue
            sel = np.nonzero(ok)[0]
            e = _pair_expected(slot, lv, p, mid, fi[sel], fj[sel], bi[sel], bj[sel])
            exp[sel] = (exp[sel] + e).astype(F32)
    # observed block: obs[sub_i[a], sub_j[b]] (read once, Q3; symmetric matrix, Q2)
    si, sj = lv.sub_id[bi, :3], lv.sub_id[bj, :3]
    li, lj = lv.limit[bi], lv.limit[bj]
    a = np.arange(3)
    va = a[None, :] <= li[:, None]
    vb = a[None, :] <= lj[:, None]
    valid = va[:, :, None] & vb[:, None, :]
    valid &= ~on_diag[:, None, None] | (a[None, :, None] < a[None, None, :])     # Q4: a < b on the diagonal
    rr = np.where(va, si, 0)[:, :, None]
    cc = np.where(vb, sj, 0)[:, None, :]
    obs = lv.obs[np.broadcast_to(rr, (P, 3, 3)), np.broadcast_to(cc, (P, 3, 3))]
    ll = evaluate_likelihood_double(exp.astype(np.float64), np.asarray(obs, dtype=np.float64))
    return np.where(valid, ll, 0.0).sum(axis=(1, 2))


def evaluate_likelihood(slot, lv, p, mid=None):
    """kernels3.cu:2802-3222 -> float64[N(N-1)/2 + N] (the reference's curr_likelihood vector)."""
    N = lv.n_frags
    n_up = N * (N - 1) // 2
    a, b = pix_decode(np.arange(n_up, dtype=np.int64))
    out = np.empty(n_up + N, dtype=np.float64)
    if mid is None:
        mid = frag_geometry(slot, lv)
    out[:n_up] = pixel_loglik(slot, lv, p, a, b, False, mid)
    d = np.arange(N, dtype=np.int64)
    out[n_up:] = pixel_loglik(slot, lv, p, d, d, True, mid)
    return out


def delta_pixels(lv, sub_index_no_repeats, list_rep, uniq_frags):
    """Pixel list of sub_compute_likelihood's four ranges (kernels3.cu:3356-3380) as built by
    stream_likelihood (cuda_lib_gl.py:2457-2483): returns (bi, bj, on_diag, glob_index)."""
    N = lv.n_frags
    u = np.asarray(sub_index_no_repeats, dtype=np.int64)
    r = np.asarray(list_rep, dtype=np.int64)
    q = np.asarray(uniq_frags, dtype=np.int64)
    parts = []
    m = u.shape[0]
    if m > 1:
        ia, ib = pix_decode(np.arange(m * (m - 1) // 2, dtype=np.int64))
        parts.append((np.minimum(u[ia], u[ib]), np.maximum(u[ia], u[ib]), False))
    nr = r.shape[0]
    if nr > 0:
        k = np.arange(nr * q.shape[0], dtype=np.int64)
        ti, tj = r[k // q.shape[0]], q[k % q.shape[0]]
        parts.append((np.minimum(ti, tj), np.maximum(ti, tj), False))
        if nr > 1:
            ia, ib = pix_decode(np.arange(nr * (nr - 1) // 2, dtype=np.int64))
            parts.append((np.minimum(r[ia], r[ib]), np.maximum(r[ia], r[ib]), False))
        parts.append((r, r, True))
    if not parts:
        z = np.zeros(0, dtype=np.int64)
        return z, z, np.zeros(0, dtype=bool), z
    bi = np.concatenate([x[0] for x in parts])
    bj = np.concatenate([x[1] for x in parts])
    dg = np.concatenate([np.full(x[0].shape[0], x[2], dtype=bool) for x in parts])
    glob = np.where(dg, N * (N - 1) // 2 + bi, pix_index(bi, bj))
    return bi, bj, dg, glob


def sub_compute_likelihood(slot, lv, p, curr_likelihood, sub_index_no_repeats, list_rep, uniq_frags):
    """kernels3.cu:3259-3718: sum over the touched pixels of (new - cached old), float64."""
    bi, bj, dg, glob = delta_pixels(lv, sub_index_no_repeats, list_rep, uniq_frags)
    if bi.shape[0] == 0:
        return np.float64(0.0)
    new = pixel_loglik(slot, lv, p, bi, bj, dg)
    return np.float64(np.sum(new - curr_likelihood[glob]))
