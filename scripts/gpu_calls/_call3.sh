set -x
timeout 900 python scripts/bench_full.py c2 20 > gpurun_out/r2c_full_c2.log 2>&1; echo "rc=$?"; cut -c1-200 gpurun_out/r2c_full_c2.log | tail -20
timeout 900 python scripts/bench_full.py c4 10 > gpurun_out/r2c_full_c4.log 2>&1; echo "rc=$?"; cut -c1-200 gpurun_out/r2c_full_c4.log | tail -20
timeout 1500 python -m pytest tests/test_gpu_bench_configs.py -m gpu -x -q -s --durations=10 > gpurun_out/r2c_tests_bench.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r2c_tests_bench.log
