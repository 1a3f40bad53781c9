"""Sparse-equivalent restatement of the dense likelihood (TEST INFRASTRUCTURE).

The reference likelihood is dense: every zero pixel contributes ``-expected``
(kernels3.cu:204-206).  This module states the SAME sum over stored contacts plus the expected
mass of all pixels, and is proven equal to ``oracle.likelihood`` (the dense transcription) on
small levels by tests/test_oracle.py.  It is the contract of the CUDA kernels and the only oracle
for levels too large to densify.

    L(S) = sum_{contacts r<c} [ ob*log(ex_S(r,c)) - lf(ob) ]      (ex != 0)
           - [ G0 + Q(S) + B(S) ]
    G0   = sum over ALL unordered sub-frag pairs of g(a_r * a_c),   g(P) = f32(v_inter * f32(f32(P)/nfpb))
    Q(S) = sum over trans bin pairs (bi < bj), ori[bi] == -1, of the Q1-quirk correction
           g(a_i[limit] * a_j[b]) - g(a_i[a] * a_j[b])               (kernels3.cu:3155,3638)
    B(S) = sum over cis sub-frag pairs with 0 < s < d_max of  ex_S(r,c) - g(a_r * a_c)
           (pairs outside the band evaluate to the clamp value bit-exactly and contribute 0)

Unique bins only (no repeat copies); repeats go through the dense pixel path.
"""
import numpy as np
from . import likelihood as L

F32 = np.float32
I32 = np.int32


class SparseLevel:
    def __init__(self, n_frags, sub_id, sub_len_kb, sub_accu, nfpb, rows, cols, vals):
        """rows < cols: upper triangle of the prepared sub-level matrix (diagonal removed)."""
        self.n_frags = int(n_frags)
        self.sub_id = np.asarray(sub_id, dtype=I32).reshape(-1, 4)
        self.sub_len = np.asarray(sub_len_kb, dtype=F32).reshape(-1, 3)
        self.sub_accu = np.asarray(sub_accu, dtype=I32).reshape(-1, 3)
        self.nfpb = F32(nfpb)
        self.limit = self.sub_id[:, 3] - 1
        cnt = self.sub_id[:, 3]
        self.W = int(cnt.sum())
        self.sub2bin = np.repeat(np.arange(self.n_frags, dtype=I32), cnt)
        self.sub_local = (np.arange(self.W) - self.sub_id[self.sub2bin, 0]).astype(I32)
        assert np.array_equal(self.sub_id[self.sub2bin, 0] + self.sub_local, np.arange(self.W))
        self.accu_true = self.sub_accu[self.sub2bin, self.sub_local].astype(np.int64)
        self.accu_last = self.sub_accu[self.sub2bin, self.limit[self.sub2bin]].astype(np.int64)
        keep = rows != cols
        self.rows = np.asarray(rows)[keep].astype(np.int64)
        self.cols = np.asarray(cols)[keep].astype(np.int64)
        self.vals = np.asarray(vals)[keep].astype(F32)
        assert np.all(self.rows < self.cols)
        self.lf = L.log_factorial_term(self.vals)

    @classmethod
    def from_dense(cls, dense_level):
        m = np.triu(dense_level.obs, 1)
        r, c = np.nonzero(m)
        return cls(dense_level.n_frags, dense_level.sub_id, dense_level.sub_len, dense_level.sub_accu,
                   dense_level.nfpb, r, c, m[r, c])


def g_clamp(P, p, nfpb):
    """The expected value of a clamped / trans pixel: f32(v_inter * f32(f32(P) / nfpb)) as float64."""
    norm = np.asarray(P).astype(I32).astype(F32) / nfpb
    return (p["v_inter"] * norm).astype(F32).astype(np.float64)


class Geo:
    """Per-sub-frag geometry of a slot (unique bins: frag id == data bin id)."""

    def __init__(self, slot, lv):
        n = lv.n_frags
        assert slot["pos"].shape[0] == n and np.array_equal(slot["id_d"], np.arange(n)), "unique bins only"

        class _D:  # minimal DenseLevel view for frag_geometry
            sub_len, limit = lv.sub_len, lv.limit
        mid = L.frag_geometry(slot, _D)
        b, a = lv.sub2bin, lv.sub_local
        self.mid = mid[b, a]
        self.id_c = slot["id_c"][b]
        self.pos = slot["pos"][b]
        self.circ = slot["circ"][b]
        self.s_tot = slot["l_cont_bp"][b].astype(F32) / F32(1000.0)
        self.accu_q = np.where(slot["ori"][b] == 1, lv.accu_true, lv.accu_last)
        self.ori = slot["ori"][b]

    def key(self):
        """Fields whose bitwise equality makes a pair's expected value unchanged."""
        return np.stack([self.mid.view(I32), self.id_c, self.circ, self.s_tot.view(I32),
                         self.accu_q.astype(I32), self.pos], axis=1)


def pair_expected(geo, lv, p, r, c):
    """float32 expected value of sub-frag pairs (r, c), r != c, following the pixel loop of
    kernels3.cu:2939-3204: cis -> rippe / rippe_circ of |mid_c - mid_r|, trans -> v_inter, both
    times int2float(accu product)/nfpb, with the Q1 quirk on the lower data bin in trans."""
    r = np.asarray(r, dtype=np.int64)
    c = np.asarray(c, dtype=np.int64)
    br, bc = lv.sub2bin[r], lv.sub2bin[c]
    cis = geo.id_c[r] == geo.id_c[c]
    out = np.zeros(r.shape[0], dtype=F32)
    tr = ~cis
    if np.any(tr):
        lo_is_r = br[tr] < bc[tr]
        a_lo = np.where(lo_is_r, geo.accu_q[r[tr]], geo.accu_q[c[tr]])
        a_hi = np.where(lo_is_r, lv.accu_true[c[tr]], lv.accu_true[r[tr]])
        norm = (a_lo * a_hi).astype(I32).astype(F32) / lv.nfpb
        out[tr] = (p["v_inter"] * norm).astype(F32)
    if np.any(cis):
        rr, cc = r[cis], c[cis]
        s = np.abs(geo.mid[cc] - geo.mid[rr]).astype(F32)
        norm = (lv.accu_true[rr] * lv.accu_true[cc]).astype(I32).astype(F32) / lv.nfpb
        first = np.where(geo.pos[rr] > geo.pos[cc], cc, rr)
        val = L.rippe_contacts(s, p)
        circ = geo.circ[first] == 1
        if np.any(circ):
            val[circ] = L.rippe_contacts_circ(s[circ], geo.s_tot[first][circ], p)
        out[cis] = (val * norm).astype(F32)
    return out


def contact_terms(geo, lv, p, sel=None):
    """ob*log(ex) - lf(ob) per stored contact (0 where ex == 0)."""
    r, c, v, lf = lv.rows, lv.cols, lv.vals, lv.lf
    if sel is not None:
        r, c, v, lf = r[sel], c[sel], v[sel], lf[sel]
    ex = pair_expected(geo, lv, p, r, c).astype(np.float64)
    out = np.zeros(r.shape[0], dtype=np.float64)
    nz = ex != 0
    out[nz] = v[nz].astype(np.float64) * np.log(ex[nz]) - lf[nz]
    return out, ex


def g0_mass(lv, p):
    vals, cnt = np.unique(lv.accu_true, return_counts=True)
    tot = 0.0
    for i, (a, ca) in enumerate(zip(vals, cnt)):
        tot += float(ca) * (float(ca) - 1) / 2.0 * float(g_clamp(a * a, p, lv.nfpb))
        for b, cb in zip(vals[i + 1:], cnt[i + 1:]):
            tot += float(ca) * float(cb) * float(g_clamp(a * b, p, lv.nfpb))
    return tot


def quirk_mass(slot, lv, p, bins=None):
    """Q(S) restricted to pairs inside ``bins`` (all bins when None)."""
    n = lv.n_frags
    bins = np.arange(n) if bins is None else np.sort(np.asarray(bins))
    lim = lv.limit
    quirky = np.array([np.any(lv.sub_accu[b, :lim[b] + 1] != lv.sub_accu[b, lim[b]]) for b in bins])
    tot = 0.0
    for bi in bins[quirky & (slot["ori"][bins] == -1)]:
        bj = bins[(bins > bi) & (slot["id_c"][bins] != slot["id_c"][bi])]
        if bj.size == 0:
            continue
        aj = lv.sub_accu[bj]                                   # (m,3)
        vj = np.arange(3)[None, :] <= lim[bj][:, None]
        for a in range(lim[bi] + 1):
            d = g_clamp(lv.sub_accu[bi, lim[bi]] * aj, p, lv.nfpb) - g_clamp(lv.sub_accu[bi, a] * aj, p, lv.nfpb)
            tot += float(np.where(vj, d, 0.0).sum())
    return tot


def band_mass(geo, lv, p, subs=None, include_same_bin=True):
    """B(S): sum over cis sub-frag pairs (optionally restricted to ``subs``) with 0 < s < d_max of
    ex - g(true product).  Per contig the sub-frags are sorted by mid-point and pairs are enumerated
    offset by offset until every distance at that offset reaches d_max (exact: pairs beyond the band
    contribute 0)."""
    subs = np.arange(lv.W) if subs is None else np.asarray(subs)
    tot = 0.0
    idc = geo.id_c[subs]
    order = np.lexsort((geo.mid[subs], idc))
    subs, idc = subs[order], idc[order]
    bounds = np.r_[0, np.nonzero(idc[1:] != idc[:-1])[0] + 1, subs.size]
    for a, b in zip(bounds[:-1], bounds[1:]):
        m = subs[a:b]
        mids = geo.mid[m]
        for k in range(1, m.size):
            s = mids[k:] - mids[:-k]
            inb = s < p["d_max"]
            if not inb.any():
                break
            r, c = m[:-k][inb], m[k:][inb]
            keep = s[inb] > 0
            if not include_same_bin:
                keep &= lv.sub2bin[r] != lv.sub2bin[c]
            r, c = r[keep], c[keep]
            if r.size == 0:
                continue
            lo, hi = np.minimum(r, c), np.maximum(r, c)
            ex = pair_expected(geo, lv, p, lo, hi).astype(np.float64)
            tot += float((ex - g_clamp(lv.accu_true[lo] * lv.accu_true[hi], p, lv.nfpb)).sum())
    return tot


def sparse_full(slot, lv, p):
    """Full log-likelihood == oracle.likelihood.evaluate_likelihood(...).sum()."""
    geo = Geo(slot, lv)
    t, _ = contact_terms(geo, lv, p)
    return float(t.sum()) - (g0_mass(lv, p) + quirk_mass(slot, lv, p) + band_mass(geo, lv, p))


def sparse_delta(slot_new, slot_old, lv, p, bins_u, return_mass=False):
    """Delta of sub_compute_likelihood for unique bins: pairs of DISTINCT bins inside ``bins_u``
    (range 1 of kernels3.cu:3356-3362; diagonal pixels are not re-scored, Q4).
    ``return_mass``: also return sum |contact terms| + |band mass| over new and old -- the magnitude the
    delta is a difference of (the scale of the float32 noise floor of the tolerance)."""
    bins_u = np.unique(np.asarray(bins_u))
    in_u = np.zeros(lv.n_frags, dtype=bool)
    in_u[bins_u] = True
    br, bc = lv.sub2bin[lv.rows], lv.sub2bin[lv.cols]
    sel = np.nonzero(in_u[br] & in_u[bc] & (br != bc))[0]
    subs = np.nonzero(in_u[lv.sub2bin])[0]
    out, mass = 0.0, 0.0
    for sgn, slot in ((1.0, slot_new), (-1.0, slot_old)):
        geo = Geo(slot, lv)
        t, _ = contact_terms(geo, lv, p, sel)
        q = quirk_mass(slot, lv, p, bins_u)
        b = band_mass(geo, lv, p, subs, include_same_bin=False)
        out += sgn * (float(t.sum()) - q - b)
        mass += float(np.abs(t).sum()) + abs(q) + abs(b)
    return (out, mass) if return_mass else out
