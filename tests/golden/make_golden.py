"""Generate the frozen fixtures of tests/golden/ from the NumPy oracle (run here, committed).

The reference ships no golden vectors and cannot be imported (Python 2 +
PyCUDA + OpenGL), so the fixtures are outputs of oracle/ -- itself pinned by tests/test_oracle_*.py --
on the synthetic yeast-shaped pyramid (BASELINE config C1, generator seed 20141217):

  like_c1_l{2,3}.npz   scrambled state, full log-likelihood, the 13 candidate deltas of 6 proposals
  traj_c1_l3.npz       the first 10^4 steps of start_EM from the exploded genome (3 neighbours,
                       RandomState(20141217)): list_mutations (id_fA, id_fB, id_mutation), scores
  traj_c1_l2_nuis.npz  600 steps at level 2 with nuisance-parameter sampling
  traj_c1_l1.npz       one whole cycle (1,672 steps) at level 1 -- 1,672 bins, 5,000 sub-frags, the size the original
                       kernels are benchmarked on (about 40 minutes of the dense oracle)

usage: python tests/golden/make_golden.py [like|traj3|traj2|traj1|all]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from graal_b200.level import yeast_shaped_pyramid, prepare_sampler_inputs   # noqa: E402
from graal_b200.driver import start_EM                                      # noqa: E402
from oracle import mutations as M, likelihood as L                          # noqa: E402
import helpers as H                                                         # noqa: E402

SEED = 20141217


def state_arrays(cur):
    return {"state_" + k: cur[k] for k in M.FIELDS}


def make_like(pyr, level):
    inp = prepare_sampler_inputs(pyr, level)
    o = H.make_oracle(inp, pyr, seed=SEED)
    rng = np.random.RandomState(5)
    moves = H.scramble(o, rng, 80)
    max_id = o.modify_gl_cuda_buffer()
    full = o.eval_likelihood()
    n = o.n_new_frags
    pairs, deltas, masses, cands = [], [], [], []
    while len(pairs) < 6:
        fA, fB = int(rng.randint(n)), int(rng.randint(n))
        if fA == fB:
            continue
        M.perform_modifications(o.ws, o.cur, fA, fB, max_id)
        no_rep, rep = o.candidate_index_sets(fA, fB)
        bi, bj, dg, glob = L.delta_pixels(o.lv, no_rep, rep, o.uniq_frags)
        for j in range(13):
            new = L.pixel_loglik(o.ws.collector[j], o.lv, o.param_simu, bi, bj, dg)
            old = o.curr_likelihood[glob]
            deltas.append(float(np.sum(new - old)))
            masses.append(float(np.sum(np.abs(new)) + np.sum(np.abs(old))))
            cands.append(np.stack([o.ws.collector[j][k] for k in M.FIELDS]))
        pairs.append((fA, fB))
    np.savez_compressed(os.path.join(HERE, "like_c1_l%d.npz" % level),
                        moves=np.array(moves), max_id=max_id, full=full, pairs=np.array(pairs),
                        deltas=np.array(deltas).reshape(-1, 13), masses=np.array(masses).reshape(-1, 13),
                        candidates=np.array(cands).reshape(len(pairs), 13, len(M.FIELDS), -1),
                        params=L.params_to_array(o.param_simu), **state_arrays(o.cur))
    print("like level", level, "full", full)


def make_traj(pyr, level, n_steps, sample_param, name):
    inp = prepare_sampler_inputs(pyr, level)
    o = H.make_oracle(inp, pyr, seed=SEED)
    if sample_param:
        o.bins = np.arange(10.0, 510.0, 10.0)
    margins, deltas = [], []

    def on_step(it, tr):
        margins.append(-1.0 if o.sample_margin is None else o.sample_margin)
        if it % 500 == 0:
            print(name, it, "steps", time.time() - t0, "s", flush=True)
    t0 = time.time()
    n_cycles = n_steps // o.n_new_frags + 1
    tr = start_EM(o, n_cycles, 3, sample_param=sample_param, scrambled=True, max_steps=n_steps, on_step=on_step)
    np.savez_compressed(os.path.join(HERE, name),
                        mutations=tr.mutations(), likelihood=np.array(tr.likelihood, dtype=np.float64),
                        n_contigs=np.array(tr.n_contigs), dist=np.array(tr.dist_from_init_genome),
                        success=np.array(tr.success), margins=np.array(margins),
                        fact=np.array(tr.fact, dtype=np.float64), slope=np.array(tr.slope, dtype=np.float64),
                        d_max=np.array(tr.d_max, dtype=np.float64), d_nuc=np.array(tr.d_nuc, dtype=np.float64),
                        params=L.params_to_array(H.make_oracle(inp, pyr).param_simu), seed=SEED, level=level,
                        **state_arrays(o.cur))
    print(name, "done", time.time() - t0, "s; min margin", np.min([m for m in margins if m >= 0]))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    pyr = yeast_shaped_pyramid()
    if what in ("like", "all"):
        make_like(pyr, 3)
        make_like(pyr, 2)
    if what in ("traj3", "all"):
        make_traj(pyr, 3, 10000, False, "traj_c1_l3.npz")
    if what in ("traj2", "all"):
        make_traj(pyr, 2, 600, True, "traj_c1_l2_nuis.npz")
    if what in ("traj1", "all"):
        make_traj(pyr, 1, 1672, False, "traj_c1_l1.npz")
