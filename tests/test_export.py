"""genome.fasta / info_frags.txt export (pyramid_sparse.py:1430-1488) on a synthetic FASTA."""
import os

import numpy as np

from graal_b200 import export as E
from graal_b200.level import prepare_sampler_inputs
from oracle import mutations as M
import helpers as H


def test_export_round_trip(small_pyramid, tmp_path):
    rng = np.random.RandomState(0)
    lv0 = small_pyramid.levels[0]
    names = ["chr%d" % (c + 1) for c in range(int(lv0.contig_id.max()))]
    seqs = {n: "".join(rng.choice(list("ACGTacgt"), size=int(lv0.end_pos[lv0.contig_id == c + 1].max())))
            for c, n in enumerate(names)}
    E.write_fasta(str(tmp_path / "genome_in.fasta"), seqs)
    assert E.read_fasta(str(tmp_path / "genome_in.fasta")) == seqs
    level = small_pyramid.levels[2]
    inp = prepare_sampler_inputs(small_pyramid, 2)
    o = H.make_oracle(inp, small_pyramid)
    # initial genome: every contig comes back unchanged
    out = E.generate_new_fasta(o.cur, level, names, seqs, str(tmp_path / "g0.fasta"), str(tmp_path / "i0.txt"))
    assert sorted(out.values(), key=len) == sorted(seqs.values(), key=len)
    # scrambled genome: total length conserved, every bin appears once, reversed bins are reverse-complemented
    H.scramble(o, rng, 60)
    o.modify_gl_cuda_buffer()
    out = E.generate_new_fasta(o.cur, level, names, seqs, str(tmp_path / "g1.fasta"), str(tmp_path / "i1.txt"))
    assert sum(len(s) for s in out.values()) == sum(len(s) for s in seqs.values())
    rows = [l.split("\t") for l in open(tmp_path / "i1.txt") if not l.startswith(">") and not l.startswith("init_contig")]
    assert sorted(int(r[1]) for r in rows) == list(range(inp.n_frags))
    f = int(np.nonzero(o.cur["ori"] == -1)[0][0])
    cid = int(o.cur["id_c"][f])
    members = np.nonzero(o.cur["id_c"] == cid)[0]
    members = members[np.argsort(o.cur["pos"][members])]
    off = sum(int(level.end_pos[m] - level.start_pos[m]) for m in members[:list(members).index(f)])
    piece = seqs[names[int(level.contig_id[f]) - 1]][int(level.start_pos[f]):int(level.end_pos[f])]
    assert out[cid][off:off + len(piece)] == piece[::-1].translate(str.maketrans("TAGCtagc", "ATCGATCG"))
    # the FASTA file holds the same sequences in 61-column lines (a 1-base tail is dropped, as the reference does)
    back = E.read_fasta(str(tmp_path / "g1.fasta"))
    for k, s in out.items():
        got = back["3C-assembly|contig_%d" % k]
        assert got == s or (len(s) % 61 == 1 and got == s[:-1])
    lines = open(tmp_path / "g1.fasta").read().split("\n")
    assert max(len(l) for l in lines if not l.startswith(">")) == 61
