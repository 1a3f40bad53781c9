"""The oracle -- and the device path -- against the REFERENCE'S OWN KERNELS.

kernels3.cu is compiled for the host with g++ through a CUDA shim (oracle/ref_emu: one CUDA block at a time,
one std::thread per CUDA thread, __shared__ -> static, __syncthreads -> barrier) and launched with the
reference's grid / block shapes.
  * test_*_live: where /root/reference exists (this container): oracle vs the compiled kernels on fresh cases;
  * test_*_fixture: anywhere: oracle vs tests/golden/ref_kernels.npz (written by tests/golden/make_ref_golden.py);
  * test_device_*: on a GPU: the CUDA path vs the same fixture.
Mutation kernels are compared bit for bit.  Likelihood values go through float32 powf / expf, which glibc (the
compiled reference), NumPy (the oracle) and CUDA implement within 1-2 ulp of each other: per-pixel values are
compared to 2e-6 of their magnitude, sums to 1e-7 relative, deltas to 1e-6 |delta| + 2^-22 mass."""
import os
import sys

import numpy as np
import pytest

import helpers as H
from oracle import mutations as M, likelihood as L
from oracle import ref_emu as R
from graal_b200.level import prepare_sampler_inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_ref_golden as G                      # noqa: E402  (the fixture's pyramid / level definition)

needs_reference = pytest.mark.skipif(not R.available(), reason="reference sources not present")


@pytest.fixture(scope="module")
def pyr():
    return G.pyramid()


@pytest.fixture(scope="module")
def fixture():
    return np.load(os.path.join(GOLD, "ref_kernels.npz"))


def oracle_move(op, dst, src, a, b, aux, mx, ids):
    if op == "flip": M.flip_frag(dst, src, a)
    elif op == "swap_activity": M.swap_activity_frag(dst, src, a, mx)
    elif op == "simple_copy": M.simple_copy(dst, src)
    elif op == "split": M.split_contig(dst, src, ids, a, aux, mx)
    elif op == "paste": M.paste_contigs(dst, src, a, b, mx)
    elif op == "pop_out": M.pop_out_frag(dst, src, ids, a, mx)
    else: getattr(M, "pop_in_frag_" + op[-1])(dst, src, a, b, mx, aux)


NAMES = {v: k for k, v in R.OPS.items()}


def test_moves_fixture(fixture):
    """75 kernel calls (every mutation kernel, scrambled states incl. circular contigs): bit-exact."""
    z = fixture
    n_circ = 0
    for src, spec, out, side in zip(z["mv_src"], z["mv_spec"], z["mv_out"], z["mv_side"]):
        op, a, b, aux, mx = (int(x) for x in spec)
        s = R.unpack(src)
        n_circ += int(s["circ"].max() == 1)
        exp = R.unpack(out)
        ids = np.zeros(len(s["pos"]), np.int32)
        # every value the kernel WRITES must match (the destination starts from a marker; what a kernel leaves
        # untouched -- persistent-slot semantics -- is checked by test_moves_live and tests/test_oracle_moves.py)
        d = {k: np.full_like(exp[k], -777) for k in M.FIELDS}
        oracle_move(NAMES[op], d, s, a, b, aux, mx, ids)
        for k in M.FIELDS:
            w = d[k] != -777
            assert np.array_equal(d[k][w], exp[k][w]), (NAMES[op], k, a, b, aux)
        if NAMES[op] in ("split", "pop_out"):
            assert np.array_equal(ids, side), NAMES[op]
    assert n_circ > 0


@needs_reference
def test_moves_live(pyr):
    """Fresh cases, full persistent-slot comparison (also the bins a kernel does NOT write)."""
    o = H.make_oracle(prepare_sampler_inputs(pyr, G.LEVEL), pyr)
    rng = np.random.RandomState(99)
    n = o.n_new_frags
    for rnd in range(4):
        H.scramble(o, rng, 15)
        max_id = int(o.modify_gl_cuda_buffer())
        fA, fB = (int(x) for x in rng.choice(n, 2, replace=False))
        sentinel = M.copy_slot(o.cur)
        sentinel["pos"][:] = 4242

        def both(op, s, a=0, b=0, aux=0, mx=0, with_ids=False):
            e, g = M.copy_slot(sentinel), M.copy_slot(sentinel)
            ie, ig = (np.zeros(n, np.int32), np.zeros(n, np.int32)) if with_ids else (None, None)
            oracle_move(op, e, s, a, b, aux, mx, ie)
            R.move(op, g, s, a, b, aux=aux, max_id=mx, ids=ig)
            assert H.slots_diff(e, g) == [], (op, a, b, aux)
            if with_ids:
                assert np.array_equal(ie, ig)
            return e, ie

        both("flip", o.cur, fA); both("swap_activity", o.cur, fA, mx=max_id); both("simple_copy", o.cur)
        both("split", o.cur, fA, aux=0, mx=max_id, with_ids=True); both("split", o.cur, fA, aux=1, mx=max_id, with_ids=True)
        both("paste", o.cur, fA, fB, mx=max_id)
        pop, pid = both("pop_out", o.cur, fA, mx=max_id, with_ids=True)
        for k in (1, 2, 3, 4):
            for ori in (1, -1):
                both("pop_in_%d" % k, pop, fA, fB, aux=ori, mx=int(pid.max()))
    # copy_struct (commit) incl. its id_contigs side array
    e, g = M.copy_slot(sentinel), M.copy_slot(sentinel)
    ie, ig = np.zeros(n, np.int32), np.zeros(n, np.int32)
    M.copy_struct(e, o.cur, ie)
    R.move("copy_struct", g, o.cur, ids=ig)
    assert H.slots_diff(e, g) == [] and np.array_equal(ie, ig)


def test_scalar_functions_fixture(fixture):
    z = fixture
    p = dict(zip(("kuhn", "lm", "c1", "slope", "d", "d_max", "fact", "v_inter"), z["sc_params"]))
    s = z["sc_s"]
    assert np.allclose(L.rippe_contacts(s, p), z["sc_rippe"], rtol=5e-7, atol=0)
    assert np.allclose(L.rippe_contacts_circ(s, np.full_like(s, 2000.0), p), z["sc_rippe_circ"], rtol=5e-7, atol=0)


def _likelihood_cases(z, tag, pyr):
    kw = {"allow_repeats": True} if tag == "r" else {}
    o = H.make_oracle(prepare_sampler_inputs(pyr, G.LEVEL, **kw), pyr)
    return o, z["ll_%s_states" % tag], z["ll_%s_full" % tag], z["ll_%s_props" % tag], z["ll_%s_deltas" % tag]


def oracle_delta_and_mass(o, cur_ll, fA, fB, j):
    no_rep, rep = o.candidate_index_sets(fA, fB)
    bi, bj, dg, glob = L.delta_pixels(o.lv, no_rep, rep, o.uniq_frags)
    new = L.pixel_loglik(o.ws.collector[j], o.lv, o.param_simu, bi, bj, dg)
    old = cur_ll[glob]
    return float(np.sum(new - old)), float(np.abs(new).sum() + np.abs(old).sum())


@pytest.mark.parametrize("tag", ["u", "r"])
def test_likelihood_fixture(fixture, pyr, tag):
    """evaluate_likelihood per pixel and sub_compute_likelihood for 13 candidates x 9 proposals, unique bins ("u")
    and a level with duplicated bins ("r"), as computed by the reference kernels."""
    o, states, fulls, props, deltas = _likelihood_cases(fixture, tag, pyr)
    cur_lls = []
    for st, ref in zip(states, fulls):
        slot = R.unpack(st)
        mine = L.evaluate_likelihood(slot, o.lv, o.param_simu)
        assert mine.shape == ref.shape
        assert np.all(np.abs(mine - ref) <= 2e-6 * (np.abs(ref) + 1.0) + 1e-3), float(np.max(np.abs(mine - ref)))
        assert abs(mine.sum() - ref.sum()) <= 1e-7 * abs(ref.sum())
        cur_lls.append(mine)
    for (st, fA, fB, max_id), ref in zip(props, deltas):
        for k in M.FIELDS:
            o.cur[k][:] = R.unpack(states[st])[k]
        M.perform_modifications(o.ws, o.cur, int(fA), int(fB), int(max_id))
        for j in range(13):
            d, mass = oracle_delta_and_mass(o, cur_lls[st], int(fA), int(fB), j)
            assert abs(d - ref[j]) <= 1e-6 * abs(ref[j]) + H.MASS_FLOOR * mass + 1e-9, (tag, st, fA, fB, j, d, ref[j])


@needs_reference
def test_likelihood_live(pyr):
    o = H.make_oracle(prepare_sampler_inputs(pyr, G.LEVEL), pyr)
    rng = np.random.RandomState(123)
    n = o.n_new_frags
    H.scramble(o, rng, 20)
    max_id = o.modify_gl_cuda_buffer()
    ref = R.evaluate_likelihood(o.cur, o.lv, o.param_simu)
    mine = L.evaluate_likelihood(o.cur, o.lv, o.param_simu)
    assert abs(mine.sum() - ref.sum()) <= 1e-7 * abs(ref.sum())
    fA, fB = (int(x) for x in rng.choice(n, 2, replace=False))
    M.perform_modifications(o.ws, o.cur, fA, fB, max_id)
    no_rep, rep = o.candidate_index_sets(fA, fB)
    # the reference's own fill_sub_index kernels give the same member list (order included)
    cA, cB = o.cur["id_c"][fA], o.cur["id_c"][fB]
    lA, lB = int(o.cur["l_cont"][fA]), int(o.cur["l_cont"][fB])
    si = R.fill_sub_index(o.cur, cA, cB, lA)[: lA + (lB if cB != cA else 0)]
    assert np.array_equal(np.sort(si), np.sort(np.concatenate([no_rep, rep[rep >= 0]]) if len(rep) else no_rep))
    for j in range(13):
        a = R.sub_compute_likelihood(o.ws.collector[j], o.lv, o.param_simu, ref, no_rep, rep, o.uniq_frags)
        d, mass = oracle_delta_and_mass(o, mine, fA, fB, j)
        assert abs(d - a) <= 1e-6 * abs(a) + H.MASS_FLOOR * mass + 1e-9, (j, d, a)


@needs_reference
def test_likelihood_live_circular_contig(pyr):
    """A contig closed into a circle (rippe_contacts_circ inside both likelihood kernels) and flipped bins
    (the trans accumulation quirk of the lower data bin)."""
    o = H.make_oracle(prepare_sampler_inputs(pyr, G.LEVEL), pyr)
    rng = np.random.RandomState(7)
    n = o.n_new_frags
    H.scramble(o, rng, 10, modes=[1, 1, 2, 5])                     # flips and a few insertions
    max_id = int(o.modify_gl_cuda_buffer())
    for c in np.unique(o.cur["id_c"]):
        bins = np.nonzero(o.cur["id_c"] == c)[0]
        if bins.size >= 4:
            head = int(bins[np.argmin(o.cur["pos"][bins])]); tail = int(bins[np.argmax(o.cur["pos"][bins])])
            new = M.copy_slot(o.cur)
            R.move("paste", new, o.cur, tail, head, max_id=max_id)          # the reference kernel itself closes the circle
            if new["circ"][head] == 1:
                for k in M.FIELDS:
                    o.cur[k][:] = new[k]
                break
    assert int((o.cur["circ"] == 1).sum()) >= 4 and int((o.cur["ori"] == -1).sum()) > 0
    max_id = o.modify_gl_cuda_buffer()
    ref = R.evaluate_likelihood(o.cur, o.lv, o.param_simu)
    mine = L.evaluate_likelihood(o.cur, o.lv, o.param_simu)
    assert np.all(np.abs(mine - ref) <= 2e-6 * (np.abs(ref) + 1.0) + 1e-3)
    assert abs(mine.sum() - ref.sum()) <= 1e-7 * abs(ref.sum())
    circ_bins = np.nonzero(o.cur["circ"] == 1)[0]
    for fA in (int(circ_bins[0]), int(circ_bins[-1]), int(rng.randint(n))):
        fB = int(rng.choice(np.setdiff1d(np.arange(n), [fA])))
        M.perform_modifications(o.ws, o.cur, fA, fB, max_id)
        no_rep, rep = o.candidate_index_sets(fA, fB)
        for j in range(13):
            a = R.sub_compute_likelihood(o.ws.collector[j], o.lv, o.param_simu, ref, no_rep, rep, o.uniq_frags)
            d, mass = oracle_delta_and_mass(o, mine, fA, fB, j)
            assert abs(d - a) <= 1e-6 * abs(a) + H.MASS_FLOOR * mass + 1e-9, (fA, fB, j, d, a)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["u", "r"])
def test_device_vs_reference_kernels(fixture, pyr, tag):
    """The CUDA path against the reference kernels' own numbers: full likelihood and the 13 deltas."""
    from graal_b200.sampler import sampler, CUR
    o, states, fulls, props, deltas = _likelihood_cases(fixture, tag, pyr)
    kw = {"allow_repeats": True} if tag == "r" else {}
    g = sampler.from_inputs(prepare_sampler_inputs(pyr, G.LEVEL, **kw), rng=np.random.RandomState(1))
    p, dm = H.default_params(pyr)
    g.set_parameters(p, dm)
    cur_lls = []
    for st, ref in zip(states, fulls):
        slot = R.unpack(st)
        g.slot_from_host(CUR, slot)
        got = g.eval_likelihood()
        assert abs(got - ref.sum()) <= 1e-7 * abs(ref.sum()), (tag, got, ref.sum())
        cur_lls.append(L.evaluate_likelihood(slot, o.lv, o.param_simu))
    for (st, fA, fB, max_id), ref in zip(props, deltas):
        slot = R.unpack(states[st])
        g.slot_from_host(CUR, slot)
        g.modify_gl_cuda_buffer()
        for k in M.FIELDS:
            o.cur[k][:] = slot[k]
        M.perform_modifications(o.ws, o.cur, int(fA), int(fB), int(max_id))
        g.score_neighbours(int(fA), [int(fB)])
        got = g._fetch()[16:29].copy()
        for j in range(13):
            _, mass = oracle_delta_and_mass(o, cur_lls[st], int(fA), int(fB), j)
            assert abs(got[j] - ref[j]) <= 1e-6 * abs(ref[j]) + H.MASS_FLOOR * mass + 1e-9, (tag, st, fA, fB, j, got[j], ref[j])
    g.free_gpu()
