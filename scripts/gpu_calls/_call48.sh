set -x
timeout 900 python scripts/bench_delta.py c2 60 GRAAL_DELTA_SPLIT=8 GRAAL_DELTA_SPLIT=8 > gpurun_out/r2au_ab.log 2>&1; echo rc=$?; grep -v Warn gpurun_out/r2au_ab.log | cut -c1-330 | tail -3
timeout 900 python -m pytest tests/test_gpu_likelihood.py tests/test_gpu_sampler.py -m gpu -x -q > gpurun_out/r2au_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2au_tests.log | cut -c1-200
