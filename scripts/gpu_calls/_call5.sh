set -x
timeout 900 python scripts/bench_delta.py c2 30 > gpurun_out/r2e_delta_c2.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2e_delta_c2.log
timeout 900 python scripts/bench_delta.py c4 12 > gpurun_out/r2e_delta_c4.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2e_delta_c4.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2e_tests.log
