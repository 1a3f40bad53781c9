"""Headless ``simulation`` + ``window``: what the reference's GUI does between "load a data set" and "genome.fasta"
(/root/reference/simulation_loader.py class simulation :39-125, 781-786; main_gl.py class window :22-138, 210-283, 321-342),
without wx / GLUT / OpenGL.

    sim = simulation(pyramid, name, level, n_iterations, False, None, output_folder, fasta_file, candidates_blacklist, allow_repeats)
    trace = sim.start_EM(n_neighbours, sample_param=False, scrambled=False)      # window.start_EM + save_behaviour_to_txt
    sim.export_new_fasta()                                                       # genome.fasta + info_frags.txt
    sim.release()

or from a shell:  python -m graal_b200 <data set folder> --level 2 --cycles 3 --neighbours 5 --out <folder> [--fasta genome.fa]

Constructor arguments keep the reference's order (``gl_window`` is accepted and ignored).  The level is prepared by
graal_b200.level.prepare_sampler_inputs (select_repeated_frags, modify_vect_frags, blacklist_contig: pinned to the
reference's lines by tests/test_reference_host_logic.py), the model parameters are estimated as simulation.__init__ does
(:110-122: histogram of cis distances up to the mean contig length, bins of the mean bin length, Rippe fit)."""
import os

import numpy as np

from .driver import Trace, start_EM as _start_EM, replay_simu as _replay_simu, load_mutations
from .level import prepare_sampler_inputs


class simulation:
    def __init__(self, pyramid, name, level, n_iterations, is_simu=False, gl_window=None, output_folder=".", fasta_file=None,
                 candidates_blacklist=(), allow_repeats=False, device=0, rng=None):
        from .sampler import sampler as sampler_lib
        if is_simu:
            raise ValueError("is_simu (simulate_rippe_contacts: curand data simulation of the GUI) is not part of this package")
        if level < 1:
            raise ValueError("level must be >= 1: the sampler scores level `level` on the contacts of level `level - 1`")
        self.name = name
        self.use_rippe = True
        self.str_sub_level, self.str_level = str(level - 1), str(level)
        self.allow_repeats = allow_repeats
        self.hic_pyr = pyramid
        self.output_folder = output_folder
        os.makedirs(output_folder, exist_ok=True)
        self.new_fasta = os.path.join(output_folder, "genome.fasta")
        self.info_frags = os.path.join(output_folder, "info_frags.txt")
        self.fasta_file = fasta_file
        self.n_iterations = n_iterations
        self.level = pyramid.get_level(level)
        self.sub_level = pyramid.get_level(level - 1)
        names = list(pyramid.spec.get("contig_names", [])) if isinstance(getattr(pyramid, "spec", None), dict) else []
        self.contig_names = names or ["contig_%d" % (c + 1) for c in range(int(np.max(self.level.contig_id)))]
        # blacklist_contig (:129-163) takes contig NAMES in the GUI; ids (1-based) are accepted too
        black = []
        for c in candidates_blacklist or ():
            black.append(self.contig_names.index(c) + 1 if isinstance(c, str) and c in self.contig_names else int(c))
        self.inputs = prepare_sampler_inputs(pyramid, level, allow_repeats=allow_repeats, blacklist_contigs=tuple(black))
        self.n_frags = int(self.inputs.n_new_frags)
        self.init_n_frags = int(self.inputs.n_frags)
        self.sampler = sampler_lib.from_inputs(self.inputs, device=device, rng=rng)
        self.sampler.n_iterations = n_iterations
        self.sampler.setup_texture()
        # :110-122
        self.sampler.gpu_vect_frags.copy_from_gpu()
        v = self.sampler.gpu_vect_frags
        id_start = np.nonzero(v.start_bp == 0)[0]
        mean_dist_kb = v.l_cont_bp[id_start].mean() / 1000.
        size_bin_kb = v.len_bp.mean() / 1000.0
        self.sampler.estimate_parameters(mean_dist_kb, size_bin_kb)
        self.trace = Trace()

    # ------------------------------------------------------------------ main_gl.window
    def start_EM(self, n_neighbours, sample_param=False, scrambled=False, max_steps=None, id_exp=0, on_step=None):
        """window.start_EM (main_gl.py:210-283) for ``n_iterations`` cycles, then save_behaviour_to_txt (:321-342) into the
        output folder with the reference's file names (``<id_exp>list_*.txt``)."""
        _start_EM(self.sampler, int(self.n_iterations), int(n_neighbours), sample_param=sample_param, scrambled=scrambled,
                  max_steps=max_steps, trace=self.trace, on_step=on_step)
        self.trace.save_behaviour_to_txt(self.output_folder, prefix=str(id_exp))
        return self.trace

    def replay_simu(self, folder_res, scrambled=False, id_exp=0):
        """window.replay_simu (main_gl.py:140-207): re-apply ``<id_exp>list_mutations.txt`` of an earlier run."""
        _replay_simu(self.sampler, load_mutations(os.path.join(folder_res, str(id_exp) + "list_mutations.txt")), scrambled=scrambled)

    # ------------------------------------------------------------------ simulation_loader.simulation
    def export_new_fasta(self):
        """:781-783: genome.fasta + info_frags.txt of the current genome.  Without a FASTA of the initial contigs only
        info_frags.txt (the layout) is written."""
        from .export import generate_new_fasta, read_fasta
        self.sampler.gpu_vect_frags.copy_from_gpu()
        if self.fasta_file:
            sequences = read_fasta(self.fasta_file)
            return generate_new_fasta(self.sampler.gpu_vect_frags, self.level, self.contig_names, sequences, self.new_fasta, self.info_frags)
        return write_info_frags(self.sampler.gpu_vect_frags, self.level, self.contig_names, self.info_frags)

    def release(self):
        """:785-786."""
        self.sampler.free_gpu()


def write_info_frags(vect_frags, level, contig_names, info_frags):
    """The info_frags.txt half of level.generate_new_fasta (pyramid_sparse.py:1430-1488): per contig whose bins are all
    active, its bins in position order as (init_contig, id_frag, orientation, start, end)."""
    get = (lambda k: vect_frags[k]) if isinstance(vect_frags, dict) else (lambda k: getattr(vect_frags, k))
    id_c, pos, ori, activ, id_d = (np.asarray(get(k)) for k in ("id_c", "pos", "ori", "activ", "id_d"))
    done = []
    with open(info_frags, "w") as hi:
        for id_cont in np.unique(id_c):
            frags = np.nonzero(id_c == id_cont)[0]
            if not np.all(activ[frags] == 1):
                continue
            done.append(int(id_cont))
            hi.write("%s\n" % (">3C-assembly|contig_" + str(id_cont)))
            hi.write("%s\t%s\t%s\t%s\t%s\n" % ("init_contig", "id_frag", "orientation", "start", "end"))
            for f in frags[np.argsort(pos[frags], kind="stable")]:
                k = int(id_d[f])
                hi.write("%s\t%s\t%s\t%s\t%s\n" % (contig_names[int(level.contig_id[k]) - 1], k, int(ori[f]),
                                                   int(level.start_pos[k]), int(level.end_pos[k])))
    return done


def main(argv=None):
    """python -m graal_b200: build the pyramid of a GRAAL data set folder (fragments_list.txt, info_contigs.txt,
    abs_fragments_contacts_weighted.txt), sample, write the traces and the genome."""
    import argparse
    from .pyramid_io import build_pyramid
    ap = argparse.ArgumentParser(prog="python -m graal_b200", description=main.__doc__)
    ap.add_argument("folder", help="data set folder (GRAAL text triplet)")
    ap.add_argument("--levels", type=int, default=4, help="pyramid size (main_window: size_pyramid)")
    ap.add_argument("--factor", type=int, default=3, help="sub-sampling factor between two levels")
    ap.add_argument("--level", type=int, default=2, help="level to sample")
    ap.add_argument("--cycles", type=int, default=3, help="cycles over all bins (n_iterations)")
    ap.add_argument("--neighbours", type=int, default=5, help="proposal partners per bin and cycle")
    ap.add_argument("--out", default="graal_out", help="output folder")
    ap.add_argument("--fasta", default=None, help="FASTA of the initial contigs (for genome.fasta)")
    ap.add_argument("--blacklist", default="", help="comma-separated contig names / ids to blacklist")
    ap.add_argument("--allow-repeats", action="store_true")
    ap.add_argument("--scrambled", action="store_true", help="start from the exploded genome (every bin its own contig)")
    ap.add_argument("--sample-param", action="store_true", help="sample the nuisance parameters after every step")
    ap.add_argument("--seed", type=int, default=None, help="RandomState seed (default: the global np.random, as the reference)")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--max-steps", type=int, default=None)
    a = ap.parse_args(argv)
    pyr = build_pyramid(a.folder, a.levels, a.factor)
    rng = np.random.RandomState(a.seed) if a.seed is not None else None
    sim = simulation(pyr, os.path.basename(os.path.normpath(a.folder)), a.level, a.cycles, False, None, a.out, a.fasta,
                     [c for c in a.blacklist.split(",") if c], a.allow_repeats, device=a.device, rng=rng)
    tr = sim.start_EM(a.neighbours, sample_param=a.sample_param, scrambled=a.scrambled, max_steps=a.max_steps)
    sim.export_new_fasta()
    print("steps %d  log-likelihood %.6f  contigs %d  -> %s" % (len(tr.likelihood), tr.likelihood[-1], tr.n_contigs[-1], a.out))
    sim.release()
    return 0
