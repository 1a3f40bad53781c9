set -x
timeout 900 python scripts/bench_delta.py c2 60 GRAAL_DELTA_BATCHED=0 GRAAL_DELTA_BATCHED=1 > gpurun_out/r2ac_ab.log 2>&1; echo rc=$?; grep -v Warn gpurun_out/r2ac_ab.log | cut -c1-330 | tail -3
GRAAL_DELTA_BATCHED=1 timeout 900 python -m pytest tests/test_gpu_likelihood.py tests/test_gpu_bench_configs.py -m gpu -x -q > gpurun_out/r2ac_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2ac_tests.log | cut -c1-300
