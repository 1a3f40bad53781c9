"""Quick device-vs-oracle comparison (development aid; the real tests are tests/test_gpu_*.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from graal_b200.level import yeast_shaped_pyramid, prepare_sampler_inputs
from graal_b200.sampler import sampler, CUR, CAND0
from oracle import mutations as M, likelihood as L
import helpers as H

pyr = yeast_shaped_pyramid()
for level in (3, 2):
    inp = prepare_sampler_inputs(pyr, level)
    o = H.make_oracle(inp, pyr)
    g = sampler.from_inputs(inp, rng=np.random.RandomState(1000))
    p, dm = H.default_params(pyr)
    g.set_parameters(p, dm)
    print("level", level, "N", inp.n_frags, "W", inp.init_n_sub_frags, "E", g.n_contacts)
    rng = np.random.RandomState(5)
    H.scramble(o, rng, 80, g)
    print("  state diff after scramble:", H.slots_diff(o.cur, g.slot_to_host(CUR)))
    mo = o.modify_gl_cuda_buffer(); mg = g.modify_gl_cuda_buffer()
    print("  relabel max_id", mo, mg, "diff", H.slots_diff(o.cur, g.slot_to_host(CUR)))
    fo = o.eval_likelihood(); fg = g.eval_likelihood()
    print("  full oracle %.6f gpu %.6f rel %.3e" % (fo, fg, abs(fo - fg) / abs(fo)))
    n = o.n_new_frags
    worst = 0
    for it in range(6):
        fA, fB = int(rng.randint(n)), int(rng.randint(n))
        if fA == fB: continue
        M.perform_modifications(o.ws, o.cur, fA, fB, mo)
        g.perform_modifications(fA, fB)
        for j in range(13):
            d = H.slots_diff(o.ws.collector[j], g.slot_to_host(CAND0 + j))
            if d: print("   cand", j, "differs in", d)
        no_rep, rep = o.candidate_index_sets(fA, fB)
        g.score_neighbours(fA, [fB])
        dg = g._fetch()[16:29].copy()
        for j in range(13):
            do = L.sub_compute_likelihood(o.ws.collector[j], o.lv, o.param_simu, o.curr_likelihood, no_rep, rep, o.uniq_frags)
            err = abs(do - dg[j])
            worst = max(worst, err / max(1.0, abs(do)))
            print("   %4d %4d cand %2d oracle %16.6f gpu %16.6f err %.3e" % (fA, fB, j, do, dg[j], err))
    print("  worst rel err", worst)
    # trajectory
    o.rng = np.random.RandomState(77); g.rng = np.random.RandomState(77)
    same = True
    for it in range(40):
        fA = int(np.random.RandomState(it).randint(n))
        ro = o.step_max_likelihood(fA, 3); rg = g.step_max_likelihood(fA, 3)
        ok = (ro[5], ro[6]) == (rg[5], rg[6]) and not H.slots_diff(o.cur, g.slot_to_host(CUR))
        if not ok or it % 10 == 0:
            print("   step", it, "oracle", ro[:8], "\n          gpu   ", rg[:8], "OK" if ok else "MISMATCH")
        same &= ok
        if not ok: break
    print("  trajectory identical:", same)
    bins_o, mean_o, sums_o, cnts_o = __import__("oracle.sampler", fromlist=["x"]).distance_histogram(inp.S_o_A_sub_frags, o.hic_matrix, 500.0, 10.0)
    bins_g, mean_g, sums_g, cnts_g = g.distance_histogram(500.0, 10.0)
    print("  histogram equal:", np.array_equal(cnts_o, cnts_g), np.allclose(sums_o, sums_g))
    g.free_gpu()
