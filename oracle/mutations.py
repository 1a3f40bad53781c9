"""NumPy restatement of GRAAL's structure-mutation kernels (TEST INFRASTRUCTURE).

Each function follows one ``__global__`` kernel of the reference
(/root/reference/kernels3.cu, cited per function) as *masked writes on a
persistent destination slot*: one boolean mask per branch of the kernel, the
fields a branch leaves untouched are left untouched in ``dst``.

A slot is a dict of 14 int32 arrays of length n_new_frags, in the field order of
the reference ``frag`` struct (kernels3.cu:9-24).
"""
import numpy as np

FIELDS = ("pos", "id_c", "start_bp", "len_bp", "circ", "id", "prev", "next",
          "l_cont", "l_cont_bp", "ori", "rep", "activ", "id_d")
I32 = np.int32


def new_slot(n, ori=1, activ=1):
    """Collector-slot initial content (cuda_lib_gl.py:269-287): zeros, ori=1, activ=1."""
    s = {k: np.zeros(n, dtype=I32) for k in FIELDS}
    s["ori"][:] = ori
    s["activ"][:] = activ
    return s


def copy_slot(src):
    return {k: np.array(src[k], dtype=I32, copy=True) for k in FIELDS}


def slots_equal(a, b):
    return all(np.array_equal(a[k], b[k]) for k in FIELDS)


def _put(dst, src, mask, **kw):
    """Write every field of the bins selected by ``mask``: the value given in
    ``kw`` (scalar or full-length array) or, by default, the source value."""
    if not np.any(mask):
        return
    n = src["pos"].shape[0]
    for k in FIELDS:
        if k == "id":
            v = np.arange(n, dtype=I32)
        else:
            v = kw.get(k, src[k])
        if np.ndim(v) == 0:
            dst[k][mask] = I32(v)
        else:
            dst[k][mask] = np.asarray(v, dtype=I32)[mask]


def _all(n):
    return np.ones(n, dtype=bool)


def simple_copy(dst, src):
    """kernels3.cu:3755-3774."""
    _put(dst, src, _all(src["pos"].shape[0]))


def copy_struct(dst, src, id_contigs=None):
    """kernels3.cu:3720-3742 (commit a candidate; also mirrors id_c into id_contigs)."""
    _put(dst, src, _all(src["pos"].shape[0]))
    if id_contigs is not None:
        id_contigs[:] = src["id_c"]


def flip_frag(dst, src, id_f_flip):
    """kernels3.cu:239-279."""
    n = src["pos"].shape[0]
    ori = src["ori"].copy()
    ori[id_f_flip] = ori[id_f_flip] * I32(-1)
    _put(dst, src, _all(n), ori=ori)


def swap_activity_frag(dst, src, id_f_unactiv, max_id_contig):
    """kernels3.cu:283-326 (acts only on a repeat copy, rep == 1)."""
    n = src["pos"].shape[0]
    ids = np.arange(n)
    hit = (ids == id_f_unactiv) & (src["rep"] == 1)
    act = src["activ"]
    activ = np.where(hit, (act == 0).astype(I32), act)
    id_c = np.where(hit, src["id_c"] * (act == 1) + (max_id_contig + 1) * (act == 0), src["id_c"])
    _put(dst, src, _all(n), activ=activ, id_c=id_c)


def pop_out_frag(dst, src, pop_id_contigs, id_f_pop, max_id_contig):
    """kernels3.cu:329-563: eject bin ``id_f_pop`` into a singleton contig ``max_id_contig + 1``."""
    c = src
    n = c["pos"].shape[0]
    ids = np.arange(n, dtype=I32)
    fp = id_f_pop
    cont, pos_p, lc, len_p = c["id_c"][fp], c["pos"][fp], c["l_cont"][fp], c["len_bp"][fp]
    prev_p, next_p, circ_p = c["prev"][fp], c["next"][fp], c["circ"][fp]
    pos = c["pos"]
    if lc > 2 or lc == 2:
        in_c = c["id_c"] == cont
        lt, eq, gt = in_c & (pos < pos_p), in_c & (pos == pos_p), in_c & (pos > pos_p)
        if lc > 2:
            _put(dst, c, lt,
                 prev=np.where((ids == next_p) & (circ_p == 1), prev_p, c["prev"]),
                 next=np.where(pos == pos_p - 1, next_p, c["next"]),
                 l_cont=c["l_cont"] - 1, l_cont_bp=c["l_cont_bp"] - len_p)
            _put(dst, c, gt, pos=pos - 1, start_bp=c["start_bp"] - len_p,
                 prev=np.where(pos == pos_p + 1, prev_p, c["prev"]),
                 next=np.where((ids == prev_p) & (circ_p == 1), next_p, c["next"]),
                 l_cont=c["l_cont"] - 1, l_cont_bp=c["l_cont_bp"] - len_p)
        else:
            _put(dst, c, lt, circ=0, prev=-1, next=-1,
                 l_cont=c["l_cont"] - 1, l_cont_bp=c["l_cont_bp"] - len_p)
            _put(dst, c, gt, pos=pos - 1, start_bp=c["start_bp"] - len_p, circ=0, prev=-1, next=-1,
                 l_cont=c["l_cont"] - 1, l_cont_bp=c["l_cont_bp"] - len_p)
        _put(dst, c, eq, pos=0, id_c=max_id_contig + 1, start_bp=0, circ=0, ori=1, prev=-1, next=-1,
             l_cont=1, l_cont_bp=c["len_bp"])
        _put(dst, c, ~in_c)
        pop_id_contigs[:] = np.where(eq, max_id_contig + 1, c["id_c"])
    else:
        _put(dst, c, _all(n))
        pop_id_contigs[:] = c["id_c"]


def _ins_scalars(c, id_f_pop, id_f_ins):
    fp, fi = id_f_pop, id_f_ins
    return dict(
        len_p=c["len_bp"][fp], act_p=c["activ"][fp],
        cont=c["id_c"][fi], pos_i=c["pos"][fi], lc=c["l_cont"][fi], lcb=c["l_cont_bp"][fi],
        len_i=c["len_bp"][fi], st_i=c["start_bp"][fi], prev_i=c["prev"][fi], next_i=c["next"][fi],
        circ_i=c["circ"][fi], or_i=c["ori"][fi], act_i=c["activ"][fi])


def pop_in_frag_1(dst, src, id_f_pop, id_f_ins, max_id_contig, ori_f_pop):
    """kernels3.cu:565-812: split the host contig before ``id_f_ins`` and put ``id_f_pop`` at the
    head of the downstream piece (circular host: linearised, keeps its id)."""
    c = src
    n = c["pos"].shape[0]
    ids = np.arange(n, dtype=I32)
    s = _ins_scalars(c, id_f_pop, id_f_ins)
    if not (s["act_i"] == 1 and s["act_p"] == 1):
        _put(dst, c, _all(n))
        return
    pos, st = c["pos"], c["start_bp"]
    m_pop = ids == id_f_pop
    new_id = max_id_contig + 1
    if s["circ_i"] == 0:
        lc_new = s["lc"] - s["pos_i"] + 1
        lcb_new = s["lcb"] - s["st_i"] + s["len_p"]
        _put(dst, c, m_pop, pos=0, start_bp=0, len_bp=s["len_p"], circ=0, ori=ori_f_pop, prev=-1,
             next=id_f_ins, id_c=new_id, l_cont=lc_new, l_cont_bp=lcb_new)
    else:
        _put(dst, c, m_pop, pos=0, start_bp=0, len_bp=s["len_p"], circ=0, ori=ori_f_pop, prev=-1,
             next=id_f_ins, id_c=s["cont"], l_cont=s["lc"] + 1, l_cont_bp=s["lcb"] + s["len_p"])
    oth = ~m_pop
    in_c = oth & (c["id_c"] == s["cont"])
    lt, eq, gt = in_c & (pos < s["pos_i"]), in_c & (pos == s["pos_i"]), in_c & (pos > s["pos_i"])
    if s["circ_i"] == 0:
        _put(dst, c, lt, id_c=s["cont"], circ=0,
             next=np.where(pos == s["pos_i"] - 1, -1, c["next"]),
             l_cont=s["pos_i"], l_cont_bp=s["st_i"])
        _put(dst, c, eq, pos=1, id_c=new_id, start_bp=s["len_p"], circ=0, ori=s["or_i"],
             prev=id_f_pop, next=s["next_i"], l_cont=lc_new, l_cont_bp=lcb_new)
        _put(dst, c, gt, pos=pos - s["pos_i"] + 1, id_c=new_id,
             start_bp=st - s["st_i"] + s["len_p"], circ=0, l_cont=lc_new, l_cont_bp=lcb_new)
    else:
        lc_new, lcb_new = s["lc"] + 1, s["lcb"] + s["len_p"]
        _put(dst, c, lt, pos=s["lc"] - s["pos_i"] + pos + 1, id_c=s["cont"],
             start_bp=s["lcb"] - s["st_i"] + st + s["len_p"], circ=0,
             next=np.where(pos == s["pos_i"] - 1, -1, c["next"]),
             l_cont=lc_new, l_cont_bp=lcb_new)
        _put(dst, c, eq, pos=1, id_c=s["cont"], start_bp=s["len_p"], len_bp=s["len_i"], circ=0,
             ori=s["or_i"], prev=id_f_pop, next=s["next_i"], l_cont=lc_new, l_cont_bp=lcb_new)
        _put(dst, c, gt, pos=pos - s["pos_i"] + 1, id_c=s["cont"],
             start_bp=st - s["st_i"] + s["len_p"], circ=0,
             next=np.where(ids == s["prev_i"], -1, c["next"]),
             l_cont=lc_new, l_cont_bp=lcb_new)
    _put(dst, c, oth & ~in_c)


def pop_in_frag_2(dst, src, id_f_pop, id_f_ins, max_id_contig, ori_f_pop):
    """kernels3.cu:814-1079: append ``id_f_pop`` after ``id_f_ins`` and cut there; the downstream
    piece becomes contig ``max_id_contig + 1`` (circular host: linearised, keeps its id)."""
    c = src
    n = c["pos"].shape[0]
    ids = np.arange(n, dtype=I32)
    s = _ins_scalars(c, id_f_pop, id_f_ins)
    if not (s["act_i"] == 1 and s["act_p"] == 1):
        _put(dst, c, _all(n))
        return
    pos, st = c["pos"], c["start_bp"]
    m_pop = ids == id_f_pop
    new_id = max_id_contig + 1
    end_i = s["st_i"] + s["len_i"]
    oth = ~m_pop
    in_c = oth & (c["id_c"] == s["cont"])
    lt, eq, gt = in_c & (pos < s["pos_i"]), in_c & (pos == s["pos_i"]), in_c & (pos > s["pos_i"])
    if s["circ_i"] == 0:
        lc_up, lcb_up = s["pos_i"] + 2, end_i + s["len_p"]
        _put(dst, c, m_pop, pos=s["pos_i"] + 1, id_c=s["cont"], start_bp=end_i, len_bp=s["len_p"],
             circ=0, ori=ori_f_pop, prev=id_f_ins, next=-1, l_cont=lc_up, l_cont_bp=lcb_up)
        _put(dst, c, lt, id_c=s["cont"], circ=0, l_cont=lc_up, l_cont_bp=lcb_up)
        _put(dst, c, eq, id_c=s["cont"], circ=0, ori=s["or_i"], prev=s["prev_i"], next=id_f_pop,
             l_cont=lc_up, l_cont_bp=lcb_up)
        _put(dst, c, gt, pos=pos - (s["pos_i"] + 1), id_c=new_id, start_bp=st - end_i, circ=0,
             prev=np.where(pos == s["pos_i"] + 1, -1, c["prev"]),
             l_cont=s["lc"] - (s["pos_i"] + 1), l_cont_bp=s["lcb"] - end_i)
    else:
        sh_pos = s["lc"] - (s["pos_i"] + 1)
        sh_bp = s["lcb"] - end_i
        lc_new, lcb_new = s["lc"] + 1, s["lcb"] + s["len_p"]
        _put(dst, c, m_pop, pos=sh_pos + s["pos_i"] + 1, id_c=s["cont"], start_bp=sh_bp + end_i,
             len_bp=s["len_p"], circ=0, ori=ori_f_pop, prev=id_f_ins, next=-1,
             l_cont=lc_new, l_cont_bp=lcb_new)
        _put(dst, c, lt, pos=sh_pos + pos, id_c=s["cont"], start_bp=sh_bp + st, circ=0,
             prev=np.where(ids == s["next_i"], -1, c["prev"]), l_cont=lc_new, l_cont_bp=lcb_new)
        _put(dst, c, eq, pos=sh_pos + s["pos_i"], id_c=s["cont"], start_bp=sh_bp + s["st_i"],
             len_bp=s["len_i"], circ=0, prev=s["prev_i"], next=id_f_pop,
             l_cont=lc_new, l_cont_bp=lcb_new)
        _put(dst, c, gt, pos=pos - (s["pos_i"] + 1), id_c=s["cont"], start_bp=st - end_i, circ=0,
             prev=np.where(pos == s["pos_i"] + 1, -1, c["prev"]), l_cont=lc_new, l_cont_bp=lcb_new)
    _put(dst, c, oth & ~in_c)


def pop_in_frag_3(dst, src, id_f_pop, id_f_ins, max_id_contig, ori_f_pop):
    """kernels3.cu:1081-1265: insert ``id_f_pop`` immediately right of ``id_f_ins`` (no split)."""
    c = src
    n = c["pos"].shape[0]
    ids = np.arange(n, dtype=I32)
    s = _ins_scalars(c, id_f_pop, id_f_ins)
    if not (s["act_i"] == 1 and s["act_p"] == 1):
        _put(dst, c, _all(n))
        return
    pos, st = c["pos"], c["start_bp"]
    m_pop = ids == id_f_pop
    lc_new, lcb_new = s["lc"] + 1, s["lcb"] + s["len_p"]
    _put(dst, c, m_pop, pos=s["pos_i"] + 1, id_c=s["cont"], start_bp=s["st_i"] + s["len_i"],
         len_bp=s["len_p"], circ=s["circ_i"], ori=ori_f_pop, prev=id_f_ins, next=s["next_i"],
         l_cont=lc_new, l_cont_bp=lcb_new)
    oth = ~m_pop
    in_c = oth & (c["id_c"] == s["cont"])
    lt, eq, gt = in_c & (pos < s["pos_i"]), in_c & (pos == s["pos_i"]), in_c & (pos > s["pos_i"])
    _put(dst, c, lt, id_c=s["cont"], circ=s["circ_i"],
         prev=np.where((ids == s["next_i"]) & (s["circ_i"] == 1), id_f_pop, c["prev"]),
         l_cont=lc_new, l_cont_bp=lcb_new)
    _put(dst, c, eq, id_c=s["cont"], circ=s["circ_i"], ori=s["or_i"], next=id_f_pop,
         l_cont=lc_new, l_cont_bp=lcb_new)
    _put(dst, c, gt, pos=pos + 1, id_c=s["cont"], start_bp=st + s["len_p"], circ=s["circ_i"],
         prev=np.where(pos == s["pos_i"] + 1, id_f_pop, c["prev"]),
         l_cont=lc_new, l_cont_bp=lcb_new)
    _put(dst, c, oth & ~in_c)


def pop_in_frag_4(dst, src, id_f_pop, id_f_ins, max_id_contig, ori_f_pop):
    """kernels3.cu:1267-1448: insert ``id_f_pop`` immediately left of ``id_f_ins`` (no split)."""
    c = src
    n = c["pos"].shape[0]
    ids = np.arange(n, dtype=I32)
    s = _ins_scalars(c, id_f_pop, id_f_ins)
    if not (s["act_i"] == 1 and s["act_p"] == 1):
        _put(dst, c, _all(n))
        return
    pos, st = c["pos"], c["start_bp"]
    m_pop = ids == id_f_pop
    lc_new, lcb_new = s["lc"] + 1, s["lcb"] + s["len_p"]
    _put(dst, c, m_pop, pos=s["pos_i"], id_c=s["cont"], start_bp=s["st_i"], len_bp=s["len_p"],
         circ=s["circ_i"], ori=ori_f_pop, prev=s["prev_i"], next=id_f_ins,
         l_cont=lc_new, l_cont_bp=lcb_new)
    oth = ~m_pop
    in_c = oth & (c["id_c"] == s["cont"])
    lt, eq, gt = in_c & (pos < s["pos_i"]), in_c & (pos == s["pos_i"]), in_c & (pos > s["pos_i"])
    _put(dst, c, lt, id_c=s["cont"], circ=s["circ_i"],
         next=np.where(pos == s["pos_i"] - 1, id_f_pop, c["next"]),
         l_cont=lc_new, l_cont_bp=lcb_new)
    _put(dst, c, eq, pos=s["pos_i"] + 1, id_c=s["cont"], start_bp=s["st_i"] + s["len_p"],
         circ=s["circ_i"], ori=s["or_i"], prev=id_f_pop, next=s["next_i"],
         l_cont=lc_new, l_cont_bp=lcb_new)
    _put(dst, c, gt, pos=pos + 1, id_c=s["cont"], start_bp=st + s["len_p"], circ=s["circ_i"],
         l_cont=lc_new, l_cont_bp=lcb_new)
    _put(dst, c, oth & ~in_c)


def split_contig(dst, src, split_id_contigs, id_f_cut, upstream, max_id_contig):
    """kernels3.cu:1451-1784: cut the contig of ``id_f_cut`` before (upstream=1) or after
    (upstream=0) it; the downstream piece gets id ``max_id_contig + 1``; a circular contig is
    linearised at the cut and keeps its id."""
    c = src
    n = c["pos"].shape[0]
    ids = np.arange(n, dtype=I32)
    f = id_f_cut
    cont, pos_c, lc, lcb = c["id_c"][f], c["pos"][f], c["l_cont"][f], c["l_cont_bp"][f]
    len_c, st_c, prev_c, next_c = c["len_bp"][f], c["start_bp"][f], c["prev"][f], c["next"][f]
    circ_c, act_c = c["circ"][f], c["activ"][f]
    if not (act_c == 1 and lc > 1):
        _put(dst, c, _all(n))
        split_id_contigs[:] = c["id_c"]
        return
    pos, st = c["pos"], c["start_bp"]
    new_id = max_id_contig + 1
    in_c = c["id_c"] == cont
    lt, eq, gt = in_c & (pos < pos_c), in_c & (pos == pos_c), in_c & (pos > pos_c)
    end_c = st_c + len_c
    new_idc = c["id_c"].copy()
    if circ_c == 0:
        if upstream == 1:
            _put(dst, c, lt, id_c=cont, circ=0, next=np.where(pos == pos_c - 1, -1, c["next"]),
                 l_cont=pos_c, l_cont_bp=st_c)
            _put(dst, c, eq, pos=0, id_c=new_id, start_bp=0, len_bp=len_c, circ=0, prev=-1,
                 next=next_c, l_cont=lc - pos_c, l_cont_bp=lcb - st_c)
            _put(dst, c, gt, pos=pos - pos_c, id_c=new_id, start_bp=st - st_c, circ=0,
                 l_cont=lc - pos_c, l_cont_bp=lcb - st_c)
            new_idc[eq | gt] = new_id
        else:
            _put(dst, c, lt, id_c=cont, circ=0, l_cont=pos_c + 1, l_cont_bp=end_c)
            _put(dst, c, eq, pos=pos_c, id_c=cont, start_bp=st_c, len_bp=len_c, circ=0, prev=prev_c,
                 next=-1, l_cont=pos_c + 1, l_cont_bp=end_c)
            _put(dst, c, gt, pos=pos - (pos_c + 1), id_c=new_id, start_bp=st - end_c, circ=0,
                 prev=np.where(pos == pos_c + 1, -1, c["prev"]),
                 l_cont=lc - (pos_c + 1), l_cont_bp=lcb - end_c)
            new_idc[gt] = new_id
    else:
        if upstream == 1:
            _put(dst, c, lt, pos=lc - pos_c + pos, id_c=cont, start_bp=lcb - st_c + st, circ=0,
                 next=np.where(pos == pos_c - 1, -1, c["next"]), l_cont=lc, l_cont_bp=lcb)
            _put(dst, c, eq, pos=0, id_c=cont, start_bp=0, len_bp=len_c, circ=0, prev=-1,
                 next=next_c, l_cont=lc, l_cont_bp=lcb)
            _put(dst, c, gt, pos=pos - pos_c, id_c=cont, start_bp=st - st_c, circ=0,
                 next=np.where(ids == prev_c, -1, c["next"]), l_cont=lc, l_cont_bp=lcb)
        else:
            sh_pos, sh_bp = lc - (pos_c + 1), lcb - end_c
            _put(dst, c, lt, pos=sh_pos + pos, id_c=cont, start_bp=sh_bp + st, circ=0,
                 prev=np.where(ids == next_c, -1, c["prev"]), l_cont=lc, l_cont_bp=lcb)
            _put(dst, c, eq, pos=sh_pos + pos, id_c=cont, start_bp=sh_bp + st_c, len_bp=len_c, circ=0,
                 prev=prev_c, next=-1, l_cont=lc, l_cont_bp=lcb)
            _put(dst, c, gt, pos=pos - (pos_c + 1), id_c=cont, start_bp=st - end_c, circ=0,
                 prev=np.where(pos == pos_c + 1, -1, c["prev"]), l_cont=lc, l_cont_bp=lcb)
    _put(dst, c, ~in_c)
    split_id_contigs[:] = new_idc


def paste_contigs(dst, src, id_fA, id_fB, max_id_contig):
    """kernels3.cu:1786-2070: join contig(fB) after contig(fA) (A reversed when fA is its first
    bin, B reversed when fB is not its first bin); same contig and end-to-end -> circularise;
    same contig otherwise: the contig's bins are NOT written (persistent slot)."""
    c = src
    n = c["pos"].shape[0]
    a, b = id_fA, id_fB
    cA, pA, lA, lbA, actA = c["id_c"][a], c["pos"][a], c["l_cont"][a], c["l_cont_bp"][a], c["activ"][a]
    cB, pB, lB, lbB, actB = c["id_c"][b], c["pos"][b], c["l_cont"][b], c["l_cont_bp"][b], c["activ"][b]
    if not (actA == 1 and actB == 1):
        _put(dst, c, _all(n))
        return
    pos, st, ln = c["pos"], c["start_bp"], c["len_bp"]
    if cA != cB:
        inA, inB = c["id_c"] == cA, c["id_c"] == cB
        lc_new, lcb_new = lA + lB, lbA + lbB
        if pA == 0:
            _put(dst, c, inA, pos=lA - (pos + 1), id_c=cA, start_bp=lbA - (st + ln), circ=0,
                 ori=c["ori"] * I32(-1),
                 prev=np.where(pos == lA - 1, -1, c["next"]),
                 next=np.where(pos == pA, id_fB, c["prev"]),
                 l_cont=lc_new, l_cont_bp=lcb_new)
        else:
            _put(dst, c, inA, id_c=cA, circ=0, next=np.where(pos == pA, id_fB, c["next"]),
                 l_cont=lc_new, l_cont_bp=lcb_new)
        if pB == 0:
            _put(dst, c, inB, pos=lA + pos, id_c=cA, start_bp=lbA + st, circ=0,
                 prev=np.where(pos == pB, id_fA, c["prev"]), l_cont=lc_new, l_cont_bp=lcb_new)
        else:
            _put(dst, c, inB, pos=lA + (lB - (pos + 1)), id_c=cA, start_bp=lbA + (lbB - (st + ln)),
                 circ=0, ori=c["ori"] * I32(-1),
                 prev=np.where(pos == pB, id_fA, c["next"]),
                 next=np.where(pos == 0, -1, c["prev"]),
                 l_cont=lc_new, l_cont_bp=lcb_new)
        _put(dst, c, ~(inA | inB))
    else:
        inA = c["id_c"] == cA
        if pA == 0 and pB == lA - 1:
            _put(dst, c, inA, circ=1, prev=np.where(pos == pA, id_fB, c["prev"]),
                 next=np.where(pos == lA - 1, id_fA, c["next"]), l_cont=lA, l_cont_bp=lbA)
        elif pA == lA - 1 and pB == 0:
            _put(dst, c, inA, circ=1, prev=np.where(pos == pB, id_fA, c["prev"]),
                 next=np.where(pos == lA - 1, id_fB, c["next"]), l_cont=lA, l_cont_bp=lbA)
        # else: bins of the contig are left as they were in dst
        _put(dst, c, ~inA)


def relabel_contigs(slot):
    """The structure side-effect of gl_update_pos (kernels3.cu:3848-3851) with the host map of
    modify_gl_cuda_buffer (cuda_lib_gl.py:1697-1722): contig ids become 0..n_contigs-1 in order of
    increasing contig length (length read at the first bin carrying the id).  np.argsort's default
    is unstable; we DEFINE ties to break by increasing old id (stable sort).  Returns max_id."""
    idc_un, idx_un = np.unique(slot["id_c"], return_index=True)
    lens = slot["l_cont"][idx_un]
    order = np.argsort(lens, kind="stable")
    old_2_new = np.zeros(int(idc_un.max()) + 1, dtype=I32)
    old_2_new[idc_un[order]] = np.arange(len(idc_un), dtype=I32)
    slot["id_c"][:] = old_2_new[slot["id_c"]]
    return I32(len(idc_un) - 1)


# ----------------------------------------------------------------------------------------------
# candidate construction: which kernel / slot / max_id each of the 13 modes uses
# (cuda_lib_gl.py:841-954, new_perform_modificationS :1045-1048)
# ----------------------------------------------------------------------------------------------
N_TMP_STRUCT = 13


class Workspace:
    """The persistent slots of the reference sampler (cuda_lib_gl.py:269-360)."""

    def __init__(self, n):
        self.collector = [new_slot(n) for _ in range(N_TMP_STRUCT)]
        self.pop = new_slot(n)
        self.trans1 = new_slot(n)
        self.trans2 = new_slot(n)
        self.pop_id_contigs = np.zeros(n, dtype=I32)
        self.trans1_id_contigs = np.zeros(n, dtype=I32)
        self.trans2_id_contigs = np.zeros(n, dtype=I32)


def pop_out_pop_in(ws, cur, id_f_pop, id_f_ins, mode, max_id):
    """cuda_lib_gl.py:841-914."""
    pop_out_frag(ws.pop, cur, ws.pop_id_contigs, id_f_pop, max_id)
    max_id2 = I32(ws.pop_id_contigs.max())
    dst = ws.collector[mode]
    if mode == 0:
        simple_copy(dst, ws.pop)
    elif mode == 1:
        flip_frag(dst, cur, id_f_pop)
    elif mode in (2, 3):
        pop_in_frag_1(dst, ws.pop, id_f_pop, id_f_ins, max_id2, 1 if mode == 2 else -1)
    elif mode in (4, 5):
        pop_in_frag_2(dst, ws.pop, id_f_pop, id_f_ins, max_id2, 1 if mode == 4 else -1)
    elif mode in (6, 7):
        pop_in_frag_3(dst, ws.pop, id_f_pop, id_f_ins, max_id2, 1 if mode == 6 else -1)
    elif mode == 8:
        swap_activity_frag(dst, ws.pop, id_f_pop, max_id2)


def transloc(ws, cur, id_fA, id_fB, max_id):
    """cuda_lib_gl.py:916-954: candidates 9..12."""
    mode = 0
    for up_a in (0, 1):
        split_contig(ws.trans1, cur, ws.trans1_id_contigs, id_fA, up_a, max_id)
        for up_b in (0, 1):
            max_id1 = I32(ws.trans1_id_contigs.max())
            split_contig(ws.trans2, ws.trans1, ws.trans2_id_contigs, id_fB, up_b, max_id1)
            max_id2 = I32(ws.trans2_id_contigs.max())
            paste_contigs(ws.collector[9 + mode], ws.trans2, id_fA, id_fB, max_id2)
            mode += 1


def perform_modifications(ws, cur, id_fA, id_fB, max_id):
    """cuda_lib_gl.py:1045-1048."""
    for mode in range(9):
        pop_out_pop_in(ws, cur, id_fA, id_fB, mode, max_id)
    transloc(ws, cur, id_fA, id_fB, max_id)


def apply_mutation(ws, cur, id_fA, id_fB, mode, max_id, id_contigs=None):
    """test_copy_struct (cuda_lib_gl.py:1156-1183): rebuild the sampled candidate, commit it."""
    if mode < 9:
        pop_out_pop_in(ws, cur, id_fA, id_fB, mode, max_id)
    elif mode < 13:
        transloc(ws, cur, id_fA, id_fB, max_id)
    copy_struct(cur, ws.collector[mode], id_contigs)


def check_invariants(c):
    """The reference's own structural checks: modify_genome / explode_genome
    (cuda_lib_gl.py:1530-1537) and diagnosis (:1016-1042).  Returns a list of violations."""
    bad = []
    if np.any(c["pos"] < 0): bad.append("pos<0")
    if np.any(c["l_cont"] <= 0): bad.append("l_cont<=0")
    if np.any(c["l_cont_bp"] <= 0): bad.append("l_cont_bp<=0")
    if np.any(c["start_bp"] < 0): bad.append("start_bp<0")
    if np.any(c["l_cont_bp"] - c["start_bp"] <= 0): bad.append("l_cont_bp<=start_bp")
    if np.any((c["start_bp"] != 0) & (c["pos"] == 0)): bad.append("start!=0 at pos 0")
    if np.any((c["start_bp"] == 0) & (c["pos"] != 0)): bad.append("start==0 at pos!=0")
    if np.any(c["next"] == c["id"]): bad.append("next==id")
    if np.any(c["prev"] == c["id"]): bad.append("prev==id")
    for ele in np.nonzero(c["start_bp"] == 0)[0]:
        n = int(c["l_cont"][ele])
        cur = int(ele)
        ok = True
        for _ in range(1, n):
            cur = int(c["next"][cur])
            if cur < 0:
                ok = False
                break
        if not ok:
            bad.append("broken next chain @%d" % ele)
            continue
        extrem = cur
        if c["circ"][ele] == 1:
            if extrem != c["prev"][ele] or c["next"][extrem] != ele:
                bad.append("circular closure @%d" % ele)
        else:
            if c["next"][extrem] != -1 or c["prev"][ele] != -1:
                bad.append("open ends @%d" % ele)
        for _ in range(1, n):
            cur = int(c["prev"][cur])
        if cur != ele:
            bad.append("prev chain @%d" % ele)
        # members agree on contig id / length and positions are a permutation
        cur, seen = int(ele), []
        for k in range(n):
            seen.append(cur)
            if c["pos"][cur] != k or c["id_c"][cur] != c["id_c"][ele] or c["l_cont"][cur] != n:
                bad.append("member mismatch @%d" % cur)
                break
            cur = int(c["next"][cur])
    return bad
