"""The headless simulation / command line (graal_b200/simulation.py: simulation_loader.simulation + main_gl.window without
the GUI): a GRAAL data set folder in, traces + info_frags.txt + genome.fasta out, and the run reproducible from its own
list_mutations.txt."""
import os

import numpy as np
import pytest

from graal_b200 import pyramid_io as P
from graal_b200.export import read_fasta, write_fasta
import helpers as H

pytestmark = pytest.mark.gpu


def test_command_line_end_to_end(small_pyramid, tmp_path):
    from graal_b200.simulation import main, simulation
    from graal_b200.sampler import CUR
    l0 = small_pyramid.get_level(0)
    folder = str(tmp_path / "dataset")
    P.write_dataset(folder, l0, one_based_one_per_line=True)
    names = ["contig_%d" % (c + 1) for c in range(int(l0.contig_id.max()))]
    rng = np.random.RandomState(0)
    seqs = {nm: "".join(rng.choice(list("ACGT"), size=int(l0.end_pos[l0.contig_id == c + 1].max()))) for c, nm in enumerate(names)}
    fasta = str(tmp_path / "genome_in.fa")
    write_fasta(fasta, seqs)
    out = str(tmp_path / "out")
    assert main([folder, "--levels", "3", "--level", "2", "--cycles", "1", "--neighbours", "3", "--out", out, "--fasta", fasta,
                 "--seed", "5", "--scrambled"]) == 0
    for f in ("0list_mutations.txt", "0list_likelihood.txt", "0list_n_contigs.txt", "info_frags.txt", "genome.fasta"):
        assert os.path.getsize(os.path.join(out, f)) > 0, f
    lik = np.loadtxt(os.path.join(out, "0list_likelihood.txt"))
    assert np.all(np.isfinite(lik)) and lik.max() > lik[0]        # the exploded genome is being re-assembled
    # every base of the input is in the output exactly once (bins are moved and flipped, never lost)
    got = read_fasta(os.path.join(out, "genome.fasta"))
    assert sum(len(s) for s in got.values()) == sum(len(s) for s in seqs.values())
    # the run is reproducible from its own list of mutations (window.replay_simu)
    pyr = P.build_pyramid(folder, 3)
    sim = simulation(pyr, "dataset", 2, 1, False, None, str(tmp_path / "again"), fasta, [], False, rng=np.random.RandomState(5))
    sim.replay_simu(out, scrambled=True)
    sim.export_new_fasta()
    again = read_fasta(os.path.join(str(tmp_path / "again"), "genome.fasta"))
    assert again == got
    sim.release()
