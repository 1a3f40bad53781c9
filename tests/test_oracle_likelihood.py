"""Pins for the likelihood oracle (CPU): closed-form values of the model helpers, the reference's own
cross-check full(candidate) == likelihood_t + delta (cuda_lib_gl.py:2109-2292), and the equality of
the sparse formulation (the device contract) with the dense transcription."""
import math

import numpy as np
import pytest

from graal_b200.level import prepare_sampler_inputs
from oracle import likelihood as L, mutations as M, sparse as S
import helpers as H

F32 = np.float32


def test_factorial_and_poisson_closed_forms():
    f = L.factorial_f32(np.array([0, 1, 2, 5, 9, 9.7], dtype=F32))
    assert f.tolist() == [1.0, 1.0, 2.0, 120.0, 362880.0, 362880.0]
    # n >= 10: float32 Stirling (kernels3.cu:90)
    n = F32(12.0)
    st = F32(np.power(n, n)) * F32(np.exp(-n)) * F32(np.sqrt(F32(2 * np.pi * 12.0)))
    assert L.factorial_f32(np.array([12.0], dtype=F32))[0] == st
    assert abs(float(st) / math.factorial(12) - 1) < 0.01
    # Poisson log-pmf branches (kernels3.cu:191-210)
    ex = np.array([0.0, 2.5, 2.5, 2.5, 30.0])
    ob = np.array([3.0, 0.0, 4.0, 20.0, 40.0])
    got = L.evaluate_likelihood_double(ex, ob)
    assert got[0] == 0.0
    assert got[1] == -2.5
    assert abs(got[2] - (4 * math.log(2.5) - 2.5 - math.log(24.0))) < 1e-12
    stir = lambda o: o * math.log(o) - o + math.log(math.sqrt(o * 2.0 * math.pi))
    assert abs(got[3] - (20 * math.log(2.5) - 2.5 - stir(20.0))) < 1e-12
    assert abs(got[4] - (40 * math.log(30.0) - 30.0 - stir(40.0))) < 1e-12


def test_rippe_contacts_hand_values():
    p = L.make_params(1.0, 9.6, -1.5, 3.0, 100.0, 500.0, 0.02)
    assert p["c1"] == F32(0.53 * float(F32(9.6)) ** -1.5)
    s = np.array([0.0, 1.0, 10.0, 499.9, 500.0, 600.0], dtype=F32)
    r = L.rippe_contacts(s, p)
    ref = lambda x: 0.53 * (9.6 * x) ** -1.5 * math.exp(1.0 / ((9.6 * x) ** 2 + 3.0)) * 100.0
    assert r[0] == F32(0.02)                                  # s == 0 -> clamp
    assert abs(r[1] / ref(1.0) - 1) < 1e-6 and abs(r[2] / ref(10.0) - 1) < 1e-6
    assert r[3] == max(F32(0.02), r[3]) and r[4] == F32(0.02) and r[5] == F32(0.02)   # s >= d_max -> v_inter
    # circular: symmetric in s <-> s_tot - s up to the linear normalisation, NaN-free for s > s_tot
    rc = L.rippe_contacts_circ(np.array([10.0, 90.0, 150.0], dtype=F32), F32(100.0), p)
    assert np.all(np.isfinite(rc)) and rc[2] == F32(0.02)


def test_pixel_index_round_trip():
    idx = np.arange(0, 200000, 7, dtype=np.int64)
    a, b = L.pix_decode(idx)
    assert np.all(a < b) and np.array_equal(L.pix_index(a, b), idx)
    big = np.array([10619135, 2 ** 40 + 12345], dtype=np.int64)      # the reference's float32 decode fails here (F2)
    a, b = L.pix_decode(big)
    assert np.array_equal(L.pix_index(a, b), big)


@pytest.mark.parametrize("level", [1, 2])
def test_sparse_equals_dense(small_pyramid, level):
    """L_sparse == sum(evaluate_likelihood) and sparse delta == sub_compute_likelihood on scrambled
    states with flips, circular contigs and (level 2) non-uniform accu (quirk Q1)."""
    inp = prepare_sampler_inputs(small_pyramid, level)
    o = H.make_oracle(inp, small_pyramid)
    slv = S.SparseLevel.from_dense(o.lv)
    rng = np.random.RandomState(3 + level)
    H.scramble(o, rng, 50)
    assert M.check_invariants(o.cur) == []
    dense = o.eval_likelihood()
    sp = S.sparse_full(o.cur, slv, o.param_simu)
    assert abs(dense - sp) <= 1e-10 * abs(dense)
    n = o.n_new_frags
    for it in range(4):
        max_id = o.modify_gl_cuda_buffer()
        o.eval_likelihood()
        fA, fB = (int(x) for x in rng.choice(n, 2, replace=False))
        M.perform_modifications(o.ws, o.cur, fA, fB, max_id)
        no_rep, rep = o.candidate_index_sets(fA, fB)
        for j in range(13):
            d_dense = L.sub_compute_likelihood(o.ws.collector[j], o.lv, o.param_simu, o.curr_likelihood,
                                               no_rep, rep, o.uniq_frags)
            d_sp = S.sparse_delta(o.ws.collector[j], o.cur, slv, o.param_simu, no_rep)
            assert abs(d_dense - d_sp) <= 1e-9 * max(1.0, abs(d_dense)), (fA, fB, j)


def test_delta_plus_current_equals_full_of_candidate(small_pyramid):
    """debug_step_max_likelihood's check (cuda_lib_gl.py:2196-2220) at level 1 (uniform accu): the
    only disagreement is the float32 noise of the diagonal pixels the delta never re-scores (Q4)."""
    inp = prepare_sampler_inputs(small_pyramid, 1)
    o = H.make_oracle(inp, small_pyramid)
    rng = np.random.RandomState(2)
    H.scramble(o, rng, 30)
    n = o.n_new_frags
    for it in range(3):
        max_id = o.modify_gl_cuda_buffer()
        like = o.eval_likelihood()
        fA, fB = (int(x) for x in rng.choice(n, 2, replace=False))
        M.perform_modifications(o.ws, o.cur, fA, fB, max_id)
        no_rep, rep = o.candidate_index_sets(fA, fB)
        for j in range(13):
            d = L.sub_compute_likelihood(o.ws.collector[j], o.lv, o.param_simu, o.curr_likelihood, no_rep, rep, o.uniq_frags)
            full = L.evaluate_likelihood(o.ws.collector[j], o.lv, o.param_simu).sum()
            assert abs(like + d - full) < 1e-6 * abs(full)
            # and exactly (to float64 summation order) once the diagonal pixels are discounted
            N = o.lv.n_frags
            diag_new = L.evaluate_likelihood(o.ws.collector[j], o.lv, o.param_simu)[N * (N - 1) // 2:].sum()
            diag_old = o.curr_likelihood[N * (N - 1) // 2:].sum()
            assert abs((like + d) - (full - diag_new + diag_old)) < 1e-9 * abs(full)


def test_repeat_pixels_sum_over_active_copies(small_pyramid):
    """H3: with duplicated bins the expected value of a pixel is the float32 sum over ACTIVE copy
    pairs; de-activating a copy (mode 8) removes its share."""
    inp = prepare_sampler_inputs(small_pyramid, 2, allow_repeats=True)
    if inp.n_new_frags == inp.n_frags:
        pytest.skip("no coverage outlier in this pyramid")
    o = H.make_oracle(inp, small_pyramid)
    base = o.eval_likelihood()
    rep = int(np.nonzero(o.cur["rep"] == 1)[0][0])
    max_id = o.modify_gl_cuda_buffer()
    M.apply_mutation(o.ws, o.cur, rep, 0, 8, max_id, o.id_contigs)
    assert o.cur["activ"][rep] == 0
    off = o.eval_likelihood()
    assert off != base
    max_id = o.modify_gl_cuda_buffer()
    M.apply_mutation(o.ws, o.cur, rep, 0, 8, max_id, o.id_contigs)
    assert o.cur["activ"][rep] == 1
    assert abs(o.eval_likelihood() - base) < 1e-9 * abs(base)
