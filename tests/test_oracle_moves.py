"""Pins for the mutation oracle (CPU).  The reference has no tests; these are the
substitute pins of SURVEY section 8c:
 (i)   the reference's own structure invariants after every mutation,
 (iii) the reciprocity identities of GRAALprinciple.pdf section B.3.1,
 (iv)  an independent LIST MODEL of the genome (contigs as ordered lists of (bin, orientation)) from
       which all 14 arrays are re-derived and compared with the kernel restatements -- the
       generalisation of the hand-computed 2-contig / 6-bin cases, covering linear and circular
       contigs of length 1, 2 and > 2.
"""
import numpy as np
import pytest

from oracle import mutations as M

I32 = np.int32


# ------------------------------------------------------------------------------------------------
# list model
# ------------------------------------------------------------------------------------------------
def arrays_from_lists(contigs, circ_flags, lens, ids):
    """contigs: list of lists of (bin, ori); ids: contig id per list."""
    n = len(lens)
    s = M.new_slot(n)
    for cont, circ, cid in zip(contigs, circ_flags, ids):
        L = len(cont)
        tot = sum(int(lens[b]) for b, _ in cont)
        acc = 0
        for k, (b, o) in enumerate(cont):
            s["pos"][b] = k; s["id_c"][b] = cid; s["start_bp"][b] = acc; s["len_bp"][b] = lens[b]
            s["circ"][b] = circ; s["id"][b] = b; s["l_cont"][b] = L; s["l_cont_bp"][b] = tot
            s["ori"][b] = o; s["rep"][b] = 0; s["activ"][b] = 1; s["id_d"][b] = b
            s["prev"][b] = cont[k - 1][0] if k > 0 else (cont[-1][0] if circ else -1)
            s["next"][b] = cont[k + 1][0] if k < L - 1 else (cont[0][0] if circ else -1)
            acc += int(lens[b])
    return s


def lists_from_arrays(s):
    out = {}
    for cid in np.unique(s["id_c"]):
        m = np.nonzero(s["id_c"] == cid)[0]
        m = m[np.argsort(s["pos"][m])]
        out[int(cid)] = ([(int(b), int(s["ori"][b])) for b in m], int(s["circ"][m[0]]))
    return out


def random_genome(rng, n, n_contigs, p_circ=0.3):
    perm = rng.permutation(n)
    cuts = np.sort(rng.choice(np.arange(1, n), size=n_contigs - 1, replace=False)) if n_contigs > 1 else []
    pieces = np.split(perm, cuts)
    contigs = [[(int(b), int(rng.choice([-1, 1]))) for b in p] for p in pieces]
    circ = [int(len(c) > 1 and rng.rand() < p_circ) for c in contigs]
    lens = rng.randint(100, 5000, size=n).astype(I32)
    ids = list(range(len(contigs)))
    return contigs, circ, lens, ids


def find(contigs, b):
    for ci, c in enumerate(contigs):
        for k, (x, _) in enumerate(c):
            if x == b:
                return ci, k
    raise KeyError(b)


def rev(c):
    return [(b, -o) for b, o in reversed(c)]


def model_pop_out(contigs, circ, ids, f, max_id):
    ci, k = find(contigs, f)
    if len(contigs[ci]) < 2:
        return contigs, circ, ids, max_id
    contigs = [list(c) for c in contigs]; circ = list(circ); ids = list(ids)
    contigs[ci].pop(k)
    if len(contigs[ci]) == 1:
        circ[ci] = 0
    contigs.append([(f, 1)]); circ.append(0); ids.append(max_id + 1)
    return contigs, circ, ids, max_id + 1


def model_pop_in(kind, contigs, circ, ids, f, g, max_id, ori):
    """f: a singleton bin; g: insertion partner."""
    contigs = [list(c) for c in contigs]; circ = list(circ); ids = list(ids)
    cf, _ = find(contigs, f)
    assert len(contigs[cf]) == 1
    contigs.pop(cf); circ.pop(cf); ids.pop(cf)
    cg, k = find(contigs, g)
    c = contigs[cg]
    if kind == 3:
        c.insert(k + 1, (f, ori))
    elif kind == 4:
        c.insert(k, (f, ori))
    elif kind == 1:
        if circ[cg]:
            contigs[cg] = [(f, ori)] + c[k:] + c[:k]; circ[cg] = 0
        else:
            up, down = c[:k], [(f, ori)] + c[k:]
            if up:
                contigs[cg] = up
                contigs.append(down); circ.append(0); ids.append(max_id + 1)
            else:
                contigs[cg] = down; ids[cg] = max_id + 1
    elif kind == 2:
        if circ[cg]:
            contigs[cg] = c[k + 1:] + c[:k + 1] + [(f, ori)]; circ[cg] = 0
        else:
            up, down = c[:k + 1] + [(f, ori)], c[k + 1:]
            contigs[cg] = up
            if down:
                contigs.append(down); circ.append(0); ids.append(max_id + 1)
    return contigs, circ, ids


def model_split(contigs, circ, ids, f, upstream, max_id):
    contigs = [list(c) for c in contigs]; circ = list(circ); ids = list(ids)
    ci, k = find(contigs, f)
    c = contigs[ci]
    if len(c) < 2:
        return contigs, circ, ids
    cut = k if upstream == 1 else k + 1
    if circ[ci]:
        contigs[ci] = c[cut:] + c[:cut]; circ[ci] = 0
    else:
        up, down = c[:cut], c[cut:]
        if up and down:
            contigs[ci] = up
            contigs.append(down); circ.append(0); ids.append(max_id + 1)
        elif down:            # whole contig moves to the new id (cut before its first bin)
            ids[ci] = max_id + 1
    return contigs, circ, ids


def model_paste(contigs, circ, ids, fA, fB):
    contigs = [list(c) for c in contigs]; circ = list(circ); ids = list(ids)
    ca, ka = find(contigs, fA)
    cb, kb = find(contigs, fB)
    if ca != cb:
        A = rev(contigs[ca]) if ka == 0 else contigs[ca]
        B = contigs[cb] if kb == 0 else rev(contigs[cb])
        contigs[ca] = A + B; circ[ca] = 0
        contigs.pop(cb); circ.pop(cb); ids.pop(cb)
    else:
        L = len(contigs[ca])
        if (ka == 0 and kb == L - 1) or (ka == L - 1 and kb == 0):
            circ[ca] = 1
    return contigs, circ, ids


def same_partition(a, b):
    """Equal arrays up to the contig labels; labels must induce the same partition."""
    for k in M.FIELDS:
        if k == "id_c":
            continue
        if not np.array_equal(a[k], b[k]):
            return "field %s" % k
    _, ia = np.unique(a["id_c"], return_inverse=True)
    _, ib = np.unique(b["id_c"], return_inverse=True)
    pairs = set(zip(ia.tolist(), ib.tolist()))
    if len(pairs) != len(set(ia.tolist())) or len(pairs) != len(set(ib.tolist())):
        return "partition"
    return None


GENOMES = [(12, 3), (12, 1), (9, 9), (10, 5), (7, 2), (30, 6)]


@pytest.mark.parametrize("n,nc", GENOMES)
def test_kernels_match_list_model(n, nc):
    rng = np.random.RandomState(100 * n + nc)
    for trial in range(40):
        contigs, circ, lens, ids = random_genome(rng, n, nc)
        cur = arrays_from_lists(contigs, circ, lens, ids)
        assert M.check_invariants(cur) == []
        max_id = int(cur["id_c"].max())
        fA, fB = rng.choice(n, 2, replace=False)
        fA, fB = int(fA), int(fB)
        # eject
        dst = M.new_slot(n); pid = np.zeros(n, dtype=I32)
        M.pop_out_frag(dst, cur, pid, fA, max_id)
        pc, pcirc, pids, max2 = model_pop_out(contigs, circ, ids, fA, max_id)
        exp_pop = arrays_from_lists(pc, pcirc, lens, pids)
        assert M.slots_equal(dst, exp_pop), "pop_out"
        assert int(pid.max()) == max2 and np.array_equal(pid, dst["id_c"])
        assert M.check_invariants(dst) == []
        # the four insertions from the ejected structure
        for kind, fn in ((1, M.pop_in_frag_1), (2, M.pop_in_frag_2), (3, M.pop_in_frag_3), (4, M.pop_in_frag_4)):
            for ori in (1, -1):
                out = M.new_slot(n)
                fn(out, dst, fA, fB, max2, ori)
                mc, mcirc, mids = model_pop_in(kind, pc, pcirc, pids, fA, fB, max2, ori)
                exp = arrays_from_lists(mc, mcirc, lens, mids)
                cg, kg = find(pc, fB)
                quirk = kind == 4 and pcirc[cg] and kg == 0
                if quirk:
                    # reference quirk (kernels3.cu:1393-1402): inserting left of the FIRST bin of a
                    # circular contig leaves the last bin's `next` on id_f_ins (pop_in_frag_4 is not
                    # used by step_max_likelihood); the oracle follows the reference
                    exp["next"][pc[cg][-1][0]] = fB
                why = same_partition(out, exp)
                assert why is None, "pop_in_%d ori %d: %s" % (kind, ori, why)
                if not quirk:
                    assert M.check_invariants(out) == []
        # flip
        out = M.new_slot(n)
        M.flip_frag(out, cur, fA)
        exp = M.copy_slot(cur); exp["ori"][fA] *= -1
        assert M.slots_equal(out, exp)
        # split / paste
        for up in (0, 1):
            out = M.new_slot(n); sid = np.zeros(n, dtype=I32)
            M.split_contig(out, cur, sid, fA, up, max_id)
            sc, scirc, sids = model_split(contigs, circ, ids, fA, up, max_id)
            why = same_partition(out, arrays_from_lists(sc, scirc, lens, sids))
            assert why is None, "split up=%d: %s" % (up, why)
            assert np.array_equal(sid, out["id_c"])
            assert M.check_invariants(out) == []
            # paste the two ends that the split produced back together: split o paste = identity
        if find(contigs, fA)[0] != find(contigs, fB)[0] and not circ[find(contigs, fA)[0]] and not circ[find(contigs, fB)[0]]:
            ca, ka = find(contigs, fA); cb, kb = find(contigs, fB)
            if ka in (0, len(contigs[ca]) - 1) and kb in (0, len(contigs[cb]) - 1):
                out = M.new_slot(n)
                M.paste_contigs(out, cur, fA, fB, max_id)
                mc, mcirc, mids = model_paste(contigs, circ, ids, fA, fB)
                why = same_partition(out, arrays_from_lists(mc, mcirc, lens, mids))
                assert why is None, "paste: %s" % why
                assert M.check_invariants(out) == []


def test_paste_extremities_always_covered():
    """With fA, fB at extremities of different contigs all 4 orientation cases match the model."""
    rng = np.random.RandomState(3)
    n = 14
    for trial in range(60):
        contigs, circ, lens, ids = random_genome(rng, n, 4, p_circ=0.0)
        ca, cb = rng.choice(len(contigs), 2, replace=False)
        fA = contigs[ca][0][0] if rng.rand() < 0.5 else contigs[ca][-1][0]
        fB = contigs[cb][0][0] if rng.rand() < 0.5 else contigs[cb][-1][0]
        cur = arrays_from_lists(contigs, circ, lens, ids)
        out = M.new_slot(n)
        M.paste_contigs(out, cur, fA, fB, int(cur["id_c"].max()))
        mc, mcirc, mids = model_paste(contigs, circ, ids, fA, fB)
        assert same_partition(out, arrays_from_lists(mc, mcirc, lens, mids)) is None
        assert M.check_invariants(out) == []


def test_paste_same_contig_circularises_or_leaves_slot():
    lens = np.full(5, 1000, dtype=I32)
    contigs = [[(0, 1), (1, 1), (2, -1)], [(3, 1), (4, 1)]]
    cur = arrays_from_lists(contigs, [0, 0], lens, [0, 1])
    out = M.new_slot(5)
    M.paste_contigs(out, cur, 0, 2, 1)
    exp = arrays_from_lists(contigs, [1, 0], lens, [0, 1])
    assert M.slots_equal(out, exp)
    # not end-to-end: the contig's bins are NOT written (persistent destination)
    out = M.new_slot(5)
    out["pos"][:] = 77
    M.paste_contigs(out, cur, 0, 1, 1)
    assert np.all(out["pos"][[0, 1, 2]] == 77) and np.array_equal(out["pos"][[3, 4]], cur["pos"][[3, 4]])


def test_reciprocity_identities():
    """GRAALprinciple.pdf B.3.1: flip o flip = id, eject o insert = id, split o paste = id."""
    rng = np.random.RandomState(8)
    n = 16
    for trial in range(50):
        contigs, circ, lens, ids = random_genome(rng, n, 4, p_circ=0.0)
        cur = arrays_from_lists(contigs, circ, lens, ids)
        max_id = int(cur["id_c"].max())
        f = int(rng.randint(n))
        a, b = M.new_slot(n), M.new_slot(n)
        M.flip_frag(a, cur, f); M.flip_frag(b, a, f)
        assert M.slots_equal(b, cur)
        # eject then re-insert next to the former left (or right) neighbour with the old orientation
        ci, k = find(contigs, f)
        if len(contigs[ci]) > 1:
            pid = np.zeros(n, dtype=I32)
            M.pop_out_frag(a, cur, pid, f, max_id)
            o = dict(contigs[ci])[f]
            if k > 0:
                M.pop_in_frag_3(b, a, f, contigs[ci][k - 1][0], int(pid.max()), o)
            else:
                M.pop_in_frag_4(b, a, f, contigs[ci][1][0], int(pid.max()), o)
            assert same_partition(b, cur) is None
        # split then paste back
        if len(contigs[ci]) > 1 and 0 < k:
            sid = np.zeros(n, dtype=I32)
            M.split_contig(a, cur, sid, f, 1, max_id)
            left = contigs[ci][k - 1][0]
            # paste(left piece, right piece): A = contig of `left` whose last bin is `left`
            if k - 1 != 0:      # paste reverses A when fA is its first bin; avoid that orientation
                M.paste_contigs(b, a, left, f, int(sid.max()))
                assert same_partition(b, cur) is None


def test_swap_activity_only_for_repeats():
    lens = np.full(4, 500, dtype=I32)
    cur = arrays_from_lists([[(0, 1), (1, 1)], [(2, 1)], [(3, 1)]], [0, 0, 0], lens, [0, 1, 2])
    out = M.new_slot(4)
    M.swap_activity_frag(out, cur, 2, 2)
    assert M.slots_equal(out, cur)                       # not a repeat: plain copy (Q7)
    cur["rep"][3] = 1
    M.swap_activity_frag(out, cur, 3, 2)
    assert out["activ"][3] == 0 and out["id_c"][3] == 2
    back = M.new_slot(4)
    M.swap_activity_frag(back, out, 3, 2)
    assert back["activ"][3] == 1 and back["id_c"][3] == 3


def test_relabel_orders_by_length_then_old_id():
    lens = np.full(7, 100, dtype=I32)
    contigs = [[(0, 1), (1, 1), (2, 1)], [(3, 1)], [(4, 1), (5, 1)], [(6, 1)]]
    cur = arrays_from_lists(contigs, [0, 0, 0, 0], lens, [5, 9, 2, 4])
    mx = M.relabel_contigs(cur)
    assert mx == 3
    # singletons (old ids 4, 9) first by old id, then the pair (2), then the triple (5)
    assert cur["id_c"].tolist() == [3, 3, 3, 1, 2, 2, 0]


def test_thirteen_candidates_keep_invariants(small_pyramid):
    from graal_b200.level import prepare_sampler_inputs
    inp = prepare_sampler_inputs(small_pyramid, 2)
    n = inp.n_new_frags
    cur = {k: np.array(inp.S_o_A_frags[k], dtype=I32) for k in M.FIELDS}
    ws = M.Workspace(n)
    rng = np.random.RandomState(0)
    for it in range(60):
        max_id = M.relabel_contigs(cur)
        fA, fB = rng.choice(n, 2, replace=False)
        M.perform_modifications(ws, cur, int(fA), int(fB), max_id)
        for j in range(13):
            assert M.check_invariants(ws.collector[j]) == [], "candidate %d" % j
        # candidate 8 == candidate 0 for non-repeat bins (Q7)
        assert M.slots_equal(ws.collector[8], ws.collector[0])
        M.apply_mutation(ws, cur, int(fA), int(fB), int(rng.randint(13)), max_id)
        assert M.check_invariants(cur) == []
