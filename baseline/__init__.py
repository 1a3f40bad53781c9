"""Second baseline: the reference's ORIGINAL kernels (kernels3.cu) compiled for sm_100a and driven on the B200 through
the CUDA driver API -- reference code executed as a baseline and cross-check, never product source.  See build_ref.py
(the cubin goes to the git-ignored baseline/_ref/) and ref_gpu.py (loader + the reference's launch sequences)."""
