"""Pyramid build from the reference's text triplet (pyramid_sparse.py:140-569)."""
import numpy as np

from graal_b200 import pyramid_io as P
from graal_b200.level import build_synthetic_pyramid, prepare_sampler_inputs


def test_text_round_trip_and_quirks(tmp_path):
    pyr = build_synthetic_pyramid([30_000, 20_000, 4_000, 300], 60, 3, seed=5, cis_rowsum=40.0, v_inter=0.02)
    l0 = pyr.levels[0]
    # documented format (README): 0-based ids with counts
    P.write_dataset(str(tmp_path / "doc"), l0, one_based_one_per_line=False)
    back = P.build_pyramid(str(tmp_path / "doc"), 3, reference_quirks=False)
    for a, b in zip(pyr.levels, back.levels):
        assert np.array_equal(a.contig_id, b.contig_id) and np.array_equal(a.start_pos, b.start_pos)
        assert np.array_equal(a.rows, b.rows) and np.array_equal(a.cols, b.cols) and np.array_equal(a.vals, b.vals)
        assert np.array_equal(a.n_accu, b.n_accu) and abs(a.mean_value_trans - b.mean_value_trans) < 1e-12
        for k in a.S_o_A_frags:
            assert np.array_equal(a.S_o_A_frags[k], b.S_o_A_frags[k]), k
    assert back.spec["contig_names"] == ["contig_1", "contig_2", "contig_3", "contig_4"]
    # what the reference CODE reads (Q13): 1-based ids, one contact per line, first contact line dropped per level
    P.write_dataset(str(tmp_path / "ref"), l0, one_based_one_per_line=True)
    q = P.build_pyramid(str(tmp_path / "ref"), 3, reference_quirks=True)
    assert np.array_equal(q.levels[0].rows, l0.rows) and np.array_equal(q.levels[0].vals, l0.vals)
    first = int(l0.vals[np.lexsort((l0.cols, l0.rows))[0]])
    assert int(q.levels[1].vals.sum()) == int(l0.vals.sum()) - first
    second = q.levels[1]
    first1 = int(second.vals[np.lexsort((second.cols, second.rows))[0]])
    assert int(q.levels[2].vals.sum()) == int(second.vals.sum()) - first1
    # npz persistence + the sampler inputs derive from a loaded pyramid
    P.save_pyramid(str(tmp_path / "pyr.npz"), back)
    again = P.load_pyramid(str(tmp_path / "pyr.npz"))
    i1, i2 = prepare_sampler_inputs(back, 2), prepare_sampler_inputs(again, 2)
    assert np.array_equal(i1.np_sub_frags_id, i2.np_sub_frags_id) and i1.mean_value_trans == i2.mean_value_trans
    assert np.array_equal(i1.sub_coo[2], i2.sub_coo[2])
