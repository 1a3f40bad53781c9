"""CPU restatement of GRAAL's MCMC scoring hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline /
``--impl reference`` legs may import this package.  The product
(``graal_b200``) never does: it fails loudly when the CUDA library is missing.

Parity status: the reference (koszullab/GRAAL) ships no tests, golden vectors
or fixtures for this path and cannot be imported here
(Python 2 + PyCUDA + OpenGL).  The oracle is therefore pinned by
 (i)   the structure invariants the reference itself checks
       (cuda_lib_gl.py:1016-1042, 1530-1537),
 (ii)  the reference's own cross-check ``likelihood_t + delta == full(candidate)``
       (cuda_lib_gl.py:2109-2292, debug_step_max_likelihood),
 (iii) the reciprocity identities of GRAALprinciple.pdf section B.3.1,
 (iv)  hand-computed small cases for every mutation,
and frozen fixtures under ``tests/golden/`` produced by ``tests/golden/make_golden.py``.
"PARITY UNPINNED BY THE REFERENCE" in the sense of the task statement: no
reference-held vector exists; see DESIGN.md section "Oracle".
"""
