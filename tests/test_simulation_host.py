"""Host-only pieces of graal_b200/simulation.py: the info_frags.txt writer against generate_new_fasta's, the command line's
arguments, and that a simulation without a GPU fails loudly (no CPU fallback)."""
import numpy as np
import pytest

from graal_b200 import simulation as S
from graal_b200.export import generate_new_fasta
from graal_b200.level import prepare_sampler_inputs
from oracle import mutations as M
import helpers as H


def test_info_frags_writer_matches_generate_new_fasta(small_pyramid, tmp_path):
    inp = prepare_sampler_inputs(small_pyramid, 2)
    o = H.make_oracle(inp, small_pyramid, seed=3)
    H.scramble(o, np.random.RandomState(1), 40)
    level = small_pyramid.get_level(2)
    names = ["contig_%d" % (c + 1) for c in range(int(level.contig_id.max()))]
    seqs = {nm: "A" * int(level.end_pos[level.contig_id == c + 1].max()) for c, nm in enumerate(names)}
    a, b = str(tmp_path / "a.txt"), str(tmp_path / "b.txt")
    generate_new_fasta(o.cur, level, names, seqs, str(tmp_path / "g.fa"), a)
    done = S.write_info_frags(o.cur, level, names, b)
    assert open(a).read() == open(b).read() and len(done) > 0


def test_command_line_arguments_and_no_cpu_fallback(small_pyramid, tmp_path):
    import torch
    from graal_b200 import pyramid_io as P
    from graal_b200._lib import GraalError
    with pytest.raises(SystemExit):
        S.main(["--help"])
    if torch.cuda.is_available():
        pytest.skip("GPU present: the end-to-end run is tests/test_gpu_simulation.py")
    folder = str(tmp_path / "d")
    P.write_dataset(folder, small_pyramid.get_level(0), one_based_one_per_line=True)
    with pytest.raises(GraalError):
        S.main([folder, "--levels", "3", "--level", "2", "--out", str(tmp_path / "o")])
