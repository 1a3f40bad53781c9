set -x
timeout 600 python scripts/bench_delta.py c2 60 GRAAL_DELTA_SPLIT=1,GRAAL_BAND_SPLIT=1 GRAAL_DELTA_SPLIT=8,GRAAL_BAND_SPLIT=1 GRAAL_DELTA_SPLIT=1,GRAAL_BAND_SPLIT=16 GRAAL_DELTA_SPLIT=8,GRAAL_BAND_SPLIT=16 GRAAL_DELTA_SPLIT=4,GRAAL_BAND_SPLIT=8 GRAAL_DELTA_SPLIT=8,GRAAL_BAND_SPLIT=32 > gpurun_out/r2x_ab.log 2>&1; echo rc=$?; tail -20 gpurun_out/r2x_ab.log
timeout 900 python -m pytest tests/test_gpu_likelihood.py tests/test_gpu_bench_configs.py tests/test_gpu_sampler.py -m gpu -x -q > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2x_tests.log | cut -c1-300
