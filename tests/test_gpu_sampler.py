"""The sampler surface end to end: identical accepted-move trajectories under identical RNG draws."""
import os

import numpy as np
import pytest

from graal_b200.driver import start_EM, replay_simu, load_mutations
from graal_b200.level import prepare_sampler_inputs
from oracle import mutations as M
import helpers as H

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gpu_sampler(pyr, level, seed, **kw):
    from graal_b200.sampler import sampler
    inp = prepare_sampler_inputs(pyr, level, **kw)
    g = sampler.from_inputs(inp, rng=np.random.RandomState(seed))
    p, dm = H.default_params(pyr)
    g.set_parameters(p, dm)
    return inp, g


def test_live_trajectory_vs_oracle(small_pyramid):
    """Oracle and device driven by the same schedule and the same RandomState seed: every returned
    9-tuple agrees (integers exactly, scores to 1e-9 relative) and the genomes stay bit-identical."""
    from graal_b200.sampler import CUR
    inp, g = gpu_sampler(small_pyramid, 2, 77, blacklist_contigs=(5,))
    o = H.make_oracle(inp, small_pyramid, seed=77)
    rows_o, rows_g = [], []
    tr_o = start_EM(o, 3, 3, scrambled=True, max_steps=150, on_step=lambda it, tr: rows_o.append(M.copy_slot(o.cur)) if it % 50 == 0 else None)
    tr_g = start_EM(g, 3, 3, scrambled=True, max_steps=150, on_step=lambda it, tr: rows_g.append(g.slot_to_host(CUR)) if it % 50 == 0 else None)
    assert np.array_equal(tr_o.mutations(), tr_g.mutations())
    assert tr_o.n_contigs == tr_g.n_contigs
    assert np.allclose(tr_o.likelihood, tr_g.likelihood, rtol=1e-7, atol=0)
    assert np.allclose(tr_o.mean_len, tr_g.mean_len, rtol=1e-12) and np.allclose(tr_o.dist_from_init_genome, tr_g.dist_from_init_genome, rtol=0, atol=1e-12)
    for a, b in zip(rows_o, rows_g):
        assert H.slots_diff(a, b) == []
    assert -1 in tr_g.op_sampled                               # blacklisted bins are skipped (op = -1)
    g.free_gpu()


def test_golden_trajectory_10k_steps(yeast_pyramid, tmp_path):
    """BASELINE: 'under identical RNG draws the accepted-move trajectory must match for the first 10^4
    steps' -- against the frozen oracle trajectory of tests/golden/traj_c1_l3.npz."""
    from graal_b200.sampler import CUR
    z = np.load(os.path.join(GOLD, "traj_c1_l3.npz"))
    inp, g = gpu_sampler(yeast_pyramid, int(z["level"]), int(z["seed"]))
    n_steps = z["mutations"].shape[0]
    tr = start_EM(g, n_steps // g.n_new_frags + 1, 3, scrambled=True, max_steps=n_steps)
    got = tr.mutations()
    same = np.all(got == z["mutations"], axis=1)
    first_bad = int(np.argmin(same)) if not same.all() else -1
    margin = z["margins"][first_bad] if first_bad >= 0 else None
    assert first_bad < 0, "diverged at step %d (draw-to-boundary margin %s)" % (first_bad, margin)
    assert np.allclose(tr.likelihood, z["likelihood"], rtol=1e-7, atol=0)
    assert np.array_equal(np.array(tr.n_contigs), z["n_contigs"])
    final = g.slot_to_host(CUR)
    for k in M.FIELDS:
        assert np.array_equal(final[k], z["state_" + k]), k
    # export the trace (main_gl.py:321-342) and rebuild the genome from it (replay_simu :140-207)
    tr.save_behaviour_to_txt(str(tmp_path))
    muts = load_mutations(os.path.join(str(tmp_path), "list_mutations.txt"))
    assert len(muts) == n_steps
    inp2, g2 = gpu_sampler(yeast_pyramid, int(z["level"]), 0)
    replay_simu(g2, muts, scrambled=True)
    again = g2.slot_to_host(CUR)
    g.modify_gl_cuda_buffer(); g2.modify_gl_cuda_buffer()
    assert H.slots_diff(g.slot_to_host(CUR), g2.slot_to_host(CUR)) == []
    g.free_gpu(); g2.free_gpu()


def test_golden_trajectory_with_nuisance_parameters(yeast_pyramid):
    z = np.load(os.path.join(GOLD, "traj_c1_l2_nuis.npz"))
    inp, g = gpu_sampler(yeast_pyramid, int(z["level"]), int(z["seed"]))
    g.bins = np.arange(10.0, 510.0, 10.0)
    n_steps = z["mutations"].shape[0]
    tr = start_EM(g, n_steps // g.n_new_frags + 1, 3, sample_param=True, scrambled=True, max_steps=n_steps)
    assert np.array_equal(tr.mutations(), z["mutations"]), int(np.argmin(np.all(tr.mutations() == z["mutations"], axis=1)))
    assert np.array_equal(np.array(tr.success), z["success"])
    for k in ("fact", "slope", "d_max", "d_nuc"):
        assert np.allclose(np.array(getattr(tr, k), dtype=np.float64), z[k], rtol=1e-6), k
    assert np.allclose(tr.likelihood, z["likelihood"], rtol=1e-7, atol=0)
    g.free_gpu()


def test_surface_attributes(small_pyramid):
    inp, g = gpu_sampler(small_pyramid, 2, 1)
    assert g.n_tmp_struct == 13 and len(g.modification_str) >= 13 and int(g.n_new_frags) == inp.n_new_frags
    g.setup_texture()
    g.init_likelihood()
    assert np.isfinite(g.likelihood_t)
    g.gpu_vect_frags.copy_from_gpu()
    assert np.array_equal(g.gpu_vect_frags.pos, inp.S_o_A_frags["pos"]) and np.all(g.gpu_vect_frags.ori == 1)
    content = g.genome_content()
    assert sum(len(v) for v in content.values()) == inp.n_new_frags
    assert g.gpu_launches > 0
    g.free_gpu()


def test_repeat_levels_are_refused_loudly(small_pyramid):
    from graal_b200.sampler import sampler, GraalError
    inp = prepare_sampler_inputs(small_pyramid, 2, allow_repeats=True)
    if inp.n_new_frags == inp.n_frags:
        pytest.skip("no coverage outlier in this pyramid")
    with pytest.raises(GraalError, match="repeat"):
        sampler.from_inputs(inp)
