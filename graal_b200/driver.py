"""Headless MCMC schedule and trace writers: the structure-sampling loop of
/root/reference/main_gl.py (window.start_EM :210-283, save_behaviour_to_txt :321-342,
replay_simu :140-207) without GLUT / wx.  Works on any object with the reference ``sampler``
surface (the CUDA sampler of this package; the tests also drive the NumPy oracle through it)."""
import os
import numpy as np


class Trace:
    """The lists ``window`` collects per step (main_gl.py:250-278)."""
    FILES = {"mean_len": "list_mean_len.txt", "n_contigs": "list_n_contigs.txt",
             "dist_from_init_genome": "list_dist_init_genome.txt", "likelihood": "list_likelihood.txt",
             "fact": "list_fact.txt", "slope": "list_slope.txt", "d_max": "list_d_max.txt",
             "d_nuc": "list_d_nuc.txt", "success": "list_success.txt"}

    def __init__(self):
        for k in list(self.FILES) + ["full_likelihood", "op_sampled", "id_fB_sampled", "id_fA_sampled",
                                     "d", "likelihood_nuisance"]:
            setattr(self, k, [])

    def mutations(self):
        return np.array([self.id_fA_sampled, self.id_fB_sampled, self.op_sampled], dtype=np.int64).T

    def save_behaviour_to_txt(self, folder, prefix=""):
        """main_gl.py:321-342: one value per line, plus list_mutations.txt (id_fA, id_fB, id_mutation)."""
        os.makedirs(folder, exist_ok=True)
        for k, fname in self.FILES.items():
            with open(os.path.join(folder, prefix + fname), "w") as h:
                for item in getattr(self, k):
                    h.write("%s\n" % item)
        with open(os.path.join(folder, prefix + "list_mutations.txt"), "w") as h:
            h.write("%s\t%s\t%s\n" % ("id_fA", "id_fB", "id_mutation"))
            for a, b, m in zip(self.id_fA_sampled, self.id_fB_sampled, self.op_sampled):
                h.write("%s\t%s\t%s\n" % (a, b, m))


def load_mutations(path):
    """Read a list_mutations.txt (main_gl.py:335-342)."""
    out = []
    with open(path) as h:
        h.readline()
        for line in h:
            a, b, m = line.split()
            out.append((int(a), int(b), int(m)))
    return out


def start_EM(sampler, n_cycles, n_neighbours, sample_param=False, scrambled=False, max_steps=None, dt=0.0,
             trace=None, on_step=None):
    """window.start_EM (main_gl.py:210-283).  RNG draws come from ``sampler.rng`` in the reference's
    order: one shuffle per cycle, then per bin the proposal draw, the candidate draw and (optionally)
    the nuisance-parameter draws."""
    rng = sampler.rng
    trace = Trace() if trace is None else trace
    delta = np.ones((n_cycles,), dtype=np.int32) * n_neighbours
    sampler.init_likelihood()
    sampler.modify_gl_cuda_buffer(0, dt)
    if scrambled:
        sampler.explode_genome(dt)
    list_frags = np.arange(0, sampler.n_new_frags, dtype=np.int32)
    n_iter = np.float32(n_cycles)
    it = 0
    for j in range(n_cycles):
        rng.shuffle(list_frags)
        for i in list_frags:
            o, n_contigs, min_len, mean_len, max_len, op_sampled, id_f_sampled, dist, temp = \
                sampler.step_max_likelihood(int(i), delta[j], 512, dt, np.float32(j), n_iter)
            trace.full_likelihood.append(sampler.likelihood_t)
            trace.likelihood.append(o)
            trace.n_contigs.append(n_contigs)
            trace.mean_len.append(mean_len)
            trace.op_sampled.append(int(op_sampled))
            trace.id_fB_sampled.append(int(id_f_sampled))
            trace.id_fA_sampled.append(int(i))
            trace.dist_from_init_genome.append(dist)
            if sample_param:
                fact, d, d_max, d_nuc, slope, likeli, success, y_eval = \
                    sampler.step_nuisance_parameters(dt, np.float32(j), n_iter)
            else:
                success = 1
                p = sampler.param_simu
                if isinstance(p, dict):
                    slope, d, d_max, fact, d_nuc = p["slope"], p["d"], p["d_max"], p["fact"], p["v_inter"]
                else:
                    kuhn, lm, c1, slope, d, d_max, fact, d_nuc = np.copy(p)[0]
                likeli = o
            trace.fact.append(fact); trace.d.append(d); trace.d_max.append(d_max); trace.d_nuc.append(d_nuc)
            trace.slope.append(slope); trace.likelihood_nuisance.append(likeli); trace.success.append(success)
            it += 1
            if on_step is not None:
                on_step(it, trace)
            if max_steps is not None and it >= max_steps:
                return trace
    return trace


def replay_simu(sampler, mutations, scrambled=False, dt=0.0):
    """window.replay_simu (main_gl.py:140-207): re-apply a saved list of (id_fA, id_fB, id_mutation)."""
    sampler.modify_gl_cuda_buffer(0, dt)
    if scrambled:
        sampler.explode_genome(dt)
    for a, b, m in mutations:
        if m >= 0:
            sampler.apply_replay_simu(a, b, m, dt)
