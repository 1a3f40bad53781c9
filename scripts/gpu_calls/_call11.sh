set -x
timeout 900 python scripts/bench_delta.py c2 30 GRAAL_BAND_FAST=1 > gpurun_out/r2k_delta_c2.log 2>&1; echo "rc=$?"; cut -c1-330 gpurun_out/r2k_delta_c2.log | tail -3
timeout 900 python scripts/bench_delta.py c4 12 GRAAL_BAND_FAST=1 > gpurun_out/r2k_delta_c4.log 2>&1; echo "rc=$?"; cut -c1-330 gpurun_out/r2k_delta_c4.log | tail -3
timeout 900 python -m pytest tests/test_gpu_likelihood.py tests/test_gpu_sampler.py -m gpu -x -q -k "not c4" > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2k_tests.log
