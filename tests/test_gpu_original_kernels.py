"""The reference's ORIGINAL kernels on the B200 (baseline/ref_gpu.py: kernels3.cu compiled for sm_100a, driven through the
CUDA driver API in the reference's own launch sequences) against the device path and the NumPy oracle, on BASELINE config C1
(dense sizes, N < 4,609): integer state of the 13 candidates bit-exact, full likelihood 1e-7, deltas within the float32-libm
tolerance.  Skipped where the cubin was not built (it is built where /root/reference exists and travels with the snapshot)."""
import numpy as np
import pytest

from graal_b200.level import prepare_sampler_inputs
from oracle import mutations as M, likelihood as L
import helpers as H

pytestmark = pytest.mark.gpu


def _three(pyr, level):
    from baseline import ref_gpu
    from graal_b200.sampler import sampler
    if not ref_gpu.available():
        pytest.skip("baseline/_ref/kernels3_sm100a.cubin not built")
    inp = prepare_sampler_inputs(pyr, level)
    o = H.make_oracle(inp, pyr)
    g = sampler.from_inputs(inp, rng=np.random.RandomState(1000))
    p, dm = H.default_params(pyr)
    g.set_parameters(p, dm)
    r = ref_gpu.RefGPU(inp, np.array(list(g.param_simu[0]), dtype=np.float32))
    return inp, o, g, r


@pytest.mark.parametrize("level", [3, 2, 1])
def test_original_kernels_on_b200(yeast_pyramid, level):
    from baseline.ref_gpu import CUR as RCUR, CAND0 as RC0
    from graal_b200.sampler import CUR, CAND0
    inp, o, g, r = _three(yeast_pyramid, level)
    rng = np.random.RandomState(40 + level)
    n = o.n_new_frags
    for state in ("assembled", "scrambled"):
        if state == "scrambled":
            H.scramble(o, rng, 25, g)
        max_id = int(o.modify_gl_cuda_buffer()); g.modify_gl_cuda_buffer()
        cur = g.slot_to_host(CUR)
        assert H.slots_diff(o.cur, cur) == []
        r.slot_from_host(RCUR, cur)
        fo, fg, fr = o.eval_likelihood(), g.eval_likelihood(), r.evaluate_likelihood(RCUR)
        assert abs(fr - fo) <= 1e-7 * abs(fo) and abs(fr - fg) <= 1e-7 * abs(fr), (level, state, fo, fg, fr)
        for it in range(2 if level == 1 else 3):
            fA, fB = (int(x) for x in rng.choice(n, 2, replace=False))
            if it == 0:
                fB = int(min(max(fA + 1, 0), n - 1)) if fA + 1 < n else fA - 1
            M.perform_modifications(o.ws, o.cur, fA, fB, max_id)
            _, dr = r.score_step(fA, [fB], max_id)
            g.score_neighbours(fA, [fB])
            dg = g._fetch()[16:29].copy()
            no_rep, rep = o.candidate_index_sets(fA, fB)
            bi, bj, dgm, glob = L.delta_pixels(o.lv, no_rep, rep, o.uniq_frags)
            for j in range(13):
                ref_slot = r.slot_to_host(RC0 + j)
                assert H.slots_diff(o.ws.collector[j], ref_slot) == [], (level, state, fA, fB, j, "original kernel vs oracle")
                assert H.slots_diff(ref_slot, g.slot_to_host(CAND0 + j)) == [], (level, state, fA, fB, j, "original kernel vs device")
                new = L.pixel_loglik(o.ws.collector[j], o.lv, o.param_simu, bi, bj, dgm)
                old = o.curr_likelihood[glob]
                do, mass = float(np.sum(new - old)), float(np.abs(new).sum() + np.abs(old).sum())
                assert H.delta_check(dr[j], do, mass)[0], (level, state, fA, fB, j, "original kernel vs oracle", dr[j], do)
                assert abs(dg[j] - dr[j]) <= 1e-6 * abs(dr[j]) + 2 * H.MASS_FLOOR * mass + 1e-9, (level, state, fA, fB, j, dg[j], dr[j])
    g.free_gpu()
