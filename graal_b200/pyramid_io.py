"""Pyramid build from the reference's on-disk text triplet.

Restates the input side of /root/reference/pyramid_sparse.py for Python 3 without h5py:
``fragments_list.txt`` / ``info_contigs.txt`` / ``abs_fragments_contacts_weighted.txt`` (README.md:108-115)
-> level 0 (``init_frag_list`` :326-356, ``abs_contact_2_coo_file`` :222-264) -> coarser levels
(``subsample_data_set`` :358-569) -> one ``.npz`` per pyramid instead of the HDF5 group layout
``/<level>/{data (3 x nnz int32), nfrags}`` (``fill_sparse_pyramid_level`` :267-324).

``reference_quirks=True`` reproduces what the reference code does, which differs from what its
README documents: the base contact file is read as ONE contact per line with 1-based ids (column 3 is
ignored), and every coarsening step skips the first contact line of the level below.
``remove_problematic_fragments`` (the sparsity filter, :573-848) is restated on in-memory levels
(``remove_problematic_fragments`` below; pinned against the reference's own function text in
tests/test_reference_host_logic.py); ``save_pyramid_hdf5`` / ``load_pyramid_hdf5`` write the reference's HDF5 group layout
when h5py is importable.
"""
import os

import numpy as np

from .level import PyramidLevel, Pyramid, _derive_frag_arrays, _mean_value_trans, _coarsen

I32 = np.int32


def read_dataset(folder, reference_quirks=True):
    """-> (level-0 PyramidLevel, contig names)."""
    names, n_frags = [], []
    with open(os.path.join(folder, "info_contigs.txt")) as h:
        h.readline()
        for line in h:
            d = line.rstrip("\n").split("\t")
            if len(d) >= 3:
                names.append(d[0]); n_frags.append(int(d[2]))
    cid_of = {n: i + 1 for i, n in enumerate(names)}
    cid, st, en = [], [], []
    with open(os.path.join(folder, "fragments_list.txt")) as h:
        h.readline()
        for line in h:
            d = line.rstrip("\n").split("\t")
            if len(d) >= 5:
                cid.append(cid_of[d[1]]); st.append(int(d[2])); en.append(int(d[3]))
    cid, st, en = np.array(cid, dtype=I32), np.array(st, dtype=I32), np.array(en, dtype=I32)
    if [int((cid == c + 1).sum()) for c in range(len(names))] != n_frags:
        raise ValueError("info_contigs.txt and fragments_list.txt disagree on the fragments per contig")
    a, b, c = [], [], []
    with open(os.path.join(folder, "abs_fragments_contacts_weighted.txt")) as h:
        h.readline()
        for line in h:
            d = line.split()
            if len(d) >= 2:
                a.append(int(d[0])); b.append(int(d[1])); c.append(int(d[2]) if len(d) > 2 else 1)
    a, b, c = np.array(a, dtype=np.int64), np.array(b, dtype=np.int64), np.array(c, dtype=np.int64)
    if reference_quirks:          # abs_contact_2_coo_file: 1-based ids, one contact per line
        a, b, c = a - 1, b - 1, np.ones_like(c)
    n = cid.shape[0]
    if a.size and (min(a.min(), b.min()) < 0 or max(a.max(), b.max()) >= n):
        raise ValueError("contact ids outside 0..%d (reference_quirks=%s)" % (n - 1, reference_quirks))
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    uk, inv = np.unique(lo * n + hi, return_inverse=True)
    vals = np.bincount(inv, weights=c.astype(np.float64)).astype(np.int64) if uk.size else np.zeros(0, dtype=np.int64)
    lv = PyramidLevel(level=0, contig_id=cid, start_pos=st, end_pos=en, n_accu=np.ones(n, dtype=I32),
                      sub_low=np.arange(n, dtype=I32), sub_high=np.arange(n, dtype=I32),
                      rows=(uk // n).astype(I32), cols=(uk % n).astype(I32), vals=vals.astype(I32))
    _derive_frag_arrays(lv)
    lv.mean_value_trans = _mean_value_trans(lv)
    return lv, names


def build_pyramid(folder, n_levels, factor=3, reference_quirks=True):
    """pyramid_sparse.build (:140-218) on a dataset folder -> Pyramid (spec['contig_names'] holds the names)."""
    lv, names = read_dataset(folder, reference_quirks)
    levels = [lv]
    for k in range(1, n_levels):
        prev = levels[-1]
        if reference_quirks and prev.rows.size:      # subsample_data_set (:525-528) drops the first contact line
            order = np.lexsort((prev.cols, prev.rows))
            keep = order[1:]
            prev = PyramidLevel(level=prev.level, contig_id=prev.contig_id, start_pos=prev.start_pos, end_pos=prev.end_pos,
                                n_accu=prev.n_accu, sub_low=prev.sub_low, sub_high=prev.sub_high,
                                rows=prev.rows[keep], cols=prev.cols[keep], vals=prev.vals[keep],
                                S_o_A_frags=prev.S_o_A_frags, mean_value_trans=prev.mean_value_trans)
        levels.append(_coarsen(prev, factor, k))
    return Pyramid(levels=levels, factor=factor, spec=dict(contig_names=names, source=folder, reference_quirks=reference_quirks))


def write_dataset(folder, level0, contig_names=None, one_based_one_per_line=True):
    """Write a level-0 PyramidLevel as the text triplet (to feed the reference, or for round-trip tests).
    ``one_based_one_per_line``: the layout the reference CODE reads (see module docstring)."""
    os.makedirs(folder, exist_ok=True)
    nc = int(level0.contig_id.max())
    names = contig_names or ["contig_%d" % (c + 1) for c in range(nc)]
    with open(os.path.join(folder, "fragments_list.txt"), "w") as h:
        h.write("id\tchrom\tstart_pos\tend_pos\tsize\tgc_content\n")
        rel = level0.S_o_A_frags["pos"]
        for i in range(level0.n_frags):
            h.write("%d\t%s\t%d\t%d\t%d\t%s\n" % (rel[i] + 1, names[level0.contig_id[i] - 1], level0.start_pos[i],
                                                  level0.end_pos[i], level0.end_pos[i] - level0.start_pos[i], "0.5"))
    with open(os.path.join(folder, "info_contigs.txt"), "w") as h:
        h.write("contig\tlength_kb\tn_frags\tcumul_length\n")
        cum = 0
        for c in range(nc):
            m = level0.contig_id == c + 1
            h.write("%s\t%d\t%d\t%d\n" % (names[c], int(level0.end_pos[m].max() // 1000), int(m.sum()), cum))
            cum += int(m.sum())
    with open(os.path.join(folder, "abs_fragments_contacts_weighted.txt"), "w") as h:
        h.write("id_frag_a\tid_frag_b\tn_contact\n")
        for r, c, v in zip(level0.rows, level0.cols, level0.vals):
            if one_based_one_per_line:
                for _ in range(int(v)):
                    h.write("%d\t%d\t%d\n" % (r + 1, c + 1, 1))
            else:
                h.write("%d\t%d\t%d\n" % (r, c, v))


def save_pyramid(path, pyr):
    """One .npz per pyramid: <level>_{contig_id,start_pos,end_pos,n_accu,sub_low,sub_high,data(3 x nnz)}."""
    out = {"n_levels": len(pyr.levels), "factor": pyr.factor}
    for lv in pyr.levels:
        k = str(lv.level)
        for f in ("contig_id", "start_pos", "end_pos", "n_accu", "sub_low", "sub_high"):
            out[k + "_" + f] = getattr(lv, f)
        out[k + "_data"] = np.stack([lv.rows, lv.cols, lv.vals]).astype(I32)
    np.savez_compressed(path, **out)


def load_pyramid(path):
    z = np.load(path)
    levels = []
    for k in range(int(z["n_levels"])):
        s = str(k)
        d = z[s + "_data"]
        lv = PyramidLevel(level=k, contig_id=z[s + "_contig_id"], start_pos=z[s + "_start_pos"], end_pos=z[s + "_end_pos"],
                          n_accu=z[s + "_n_accu"], sub_low=z[s + "_sub_low"], sub_high=z[s + "_sub_high"],
                          rows=d[0], cols=d[1], vals=d[2])
        _derive_frag_arrays(lv)
        lv.mean_value_trans = _mean_value_trans(lv)
        levels.append(lv)
    return Pyramid(levels=levels, factor=int(z["factor"]), spec={})


def remove_problematic_fragments(level0, accu_frag=None):
    """pyramid_sparse.remove_problematic_fragments (:573-848) on an in-memory level 0.

    A fragment whose row of the symmetric contact matrix is too sparse -- stored neighbours / n_frags <= mean - 1.01 std
    (float32, :590-618) -- is LOCKED (the reference also flags fragments of size <= 1 bp, :716-717, but tests the
    sparsity flag alone when it decides, :726: the size has no effect): it is merged into the next unlocked fragment of its
    contig (sizes and accu counts summed; the merged fragment starts where the previous written one ended, 0 at the
    start of a contig, and ends where the unlocked fragment ends, :712-745); locked fragments left at the end of a
    contig are destroyed, and so are their contacts (:661-666, 688-691, 808-826); a contig left without fragments is
    deleted (:771-776).  Contacts are re-indexed, ordered (min, max) and summed, self contacts included.
    Returns (new level 0, thresh, old_2_new int64[n] with -1 for destroyed fragments)."""
    n = level0.n_frags
    cid = np.asarray(level0.contig_id)
    rows, cols, vals = (np.asarray(a) for a in (level0.rows, level0.cols, level0.vals))
    # stored entries per row of M + M^T (a diagonal entry counts once: scipy sums the duplicates of csr + csr.T)
    nnz = np.bincount(rows, minlength=n) + np.bincount(cols[rows != cols], minlength=n)
    spars = nnz.astype(np.float32) / np.float32(n)
    thresh = spars.mean() - 1.01 * spars.std()
    size = np.asarray(level0.end_pos, dtype=np.int64) - np.asarray(level0.start_pos, dtype=np.int64)
    accu = np.ones(n, dtype=np.int64) if accu_frag is None else np.asarray(accu_frag, dtype=np.int64)
    lock = spars <= thresh
    first = np.r_[True, cid[1:] != cid[:-1]]
    old2new = np.full(n, -1, dtype=np.int64)
    new_cid, new_st, new_en, new_accu = [], [], [], []
    pending, cur_start, cum_size, cum_accu = [], 0, 0, 0
    for i in range(n):
        if first[i]:
            # pending locked fragments of the previous contig: destroyed.  Reference quirk (:674-685): their accu count is
            # NOT cleared and leaks into the first fragment written for this contig.
            pending, cur_start, cum_size = [], 0, 0
        pending.append(i)
        cum_size += int(size[i]); cum_accu += int(accu[i])
        if not lock[i]:
            k = len(new_cid)
            old2new[pending] = k
            new_cid.append(int(cid[i])); new_st.append(cur_start); new_en.append(int(level0.end_pos[i])); new_accu.append(cum_accu)
            cur_start = int(level0.end_pos[i])
            pending, cum_size, cum_accu = [], 0, 0
    m = len(new_cid)
    new_cid = np.asarray(new_cid, dtype=np.int64)
    kept = np.unique(new_cid)                                            # deleted contigs: renumber 1..n_kept in the old order
    remap = np.zeros(int(cid.max()) + 2, dtype=np.int64)
    remap[kept] = np.arange(1, kept.size + 1)
    a, b = old2new[rows], old2new[cols]
    ok = (a >= 0) & (b >= 0)
    lo, hi, v = np.minimum(a[ok], b[ok]), np.maximum(a[ok], b[ok]), vals[ok].astype(np.int64)
    uk, inv = np.unique(lo * max(m, 1) + hi, return_inverse=True)
    sv = np.bincount(inv, weights=v.astype(np.float64)).astype(np.int64) if uk.size else np.zeros(0, dtype=np.int64)
    lv = PyramidLevel(level=0, contig_id=remap[new_cid].astype(I32), start_pos=np.asarray(new_st, dtype=I32), end_pos=np.asarray(new_en, dtype=I32),
                      n_accu=np.asarray(new_accu, dtype=I32), sub_low=np.arange(m, dtype=I32), sub_high=np.arange(m, dtype=I32),
                      rows=(uk // max(m, 1)).astype(I32), cols=(uk % max(m, 1)).astype(I32), vals=sv.astype(I32))
    _derive_frag_arrays(lv)
    lv.mean_value_trans = _mean_value_trans(lv)
    return lv, float(thresh), old2new


def save_pyramid_hdf5(path, pyr):
    """The reference's HDF5 layout (fill_sparse_pyramid_level, pyramid_sparse.py:267-324): group ``/<level>`` with datasets
    ``data`` (3 x nnz int32: rows, cols, counts) and ``nfrags`` (1 int32).  Needs h5py (not part of this image: the .npz
    container above is the default)."""
    import h5py
    with h5py.File(path, "w") as h:
        for lv in pyr.levels:
            g = h.create_group(str(lv.level))
            g.create_dataset("data", data=np.stack([lv.rows, lv.cols, lv.vals]).astype(I32))
            g.create_dataset("nfrags", data=np.array([lv.n_frags], dtype=I32))


def load_pyramid_hdf5(path, fragment_tables):
    """Contacts of every level from the reference's HDF5 file; ``fragment_tables[level]`` = (contig_id, start_pos, end_pos,
    n_accu, sub_low, sub_high) from the per-level fragments lists."""
    import h5py
    levels = []
    with h5py.File(path, "r") as h:
        for k in sorted(h.keys(), key=int):
            d = np.asarray(h[k]["data"])
            cid, st, en, na, lo, hi = fragment_tables[int(k)]
            if int(np.asarray(h[k]["nfrags"])[0]) != len(cid):
                raise ValueError("level %s: nfrags disagrees with the fragment table" % k)
            lv = PyramidLevel(level=int(k), contig_id=cid, start_pos=st, end_pos=en, n_accu=na, sub_low=lo, sub_high=hi,
                              rows=d[0], cols=d[1], vals=d[2])
            _derive_frag_arrays(lv)
            lv.mean_value_trans = _mean_value_trans(lv)
            levels.append(lv)
    return levels
