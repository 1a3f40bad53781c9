"""CPU restatement of GRAAL's MCMC scoring hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline /
``--impl reference`` legs may import this package.  The product
(``graal_b200``) never does: it fails loudly when the CUDA library is missing.

Parity status: PINNED AGAINST THE REFERENCE'S OWN KERNELS.  The reference (koszullab/GRAAL) ships no tests,
golden vectors or fixtures for this path and its Python cannot be imported here
(Python 2 + PyCUDA + OpenGL), but its kernel file compiles for the HOST: ``oracle/ref_emu`` builds
``/root/reference/kernels3.cu`` where it lies with g++ through a small CUDA shim (one block at a time, one
std::thread per CUDA thread, ``__shared__`` -> static, ``__syncthreads`` -> barrier) into ``oracle/_ref/`` and
launches the kernels with the reference's grid / block shapes.  ``tests/test_reference_kernels.py`` checks
 (a) every mutation kernel (+ copy_struct, fill_sub_index) bit for bit, linear and circular contigs,
 (b) evaluate_likelihood per pixel and sub_compute_likelihood for the 13 candidates, with and without
     duplicated bins, to float32-libm accuracy,
live where the reference is present and through ``tests/golden/ref_kernels.npz`` (written by
``tests/golden/make_ref_golden.py`` from the compiled kernels) everywhere else -- the device path is checked
against the same vectors.  Independent anchors kept from before:
 (i)   the structure invariants the reference itself checks
       (cuda_lib_gl.py:1016-1042, 1530-1537),
 (ii)  the reference's own cross-check ``likelihood_t + delta == full(candidate)``
       (cuda_lib_gl.py:2109-2292, debug_step_max_likelihood),
 (iii) the reciprocity identities of GRAALprinciple.pdf section B.3.1,
 (iv)  an independent list model of the genome for every mutation (tests/test_oracle_moves.py),
and the frozen trajectories under ``tests/golden/`` produced by ``tests/golden/make_golden.py``.
The HOST logic of a step (candidate draw, return_neighbours, setup_distri_frags, dist_inter_genome) is checked
against the reference's own Python lines executed under Python 3 (tests/test_reference_host_logic.py).
"""
