set -x
timeout 900 python scripts/bench_delta.py c2 30 > gpurun_out/r2f_delta_c2.log 2>&1; echo "rc=$?"; cut -c1-330 gpurun_out/r2f_delta_c2.log | tail -5
timeout 900 python scripts/bench_delta.py c4 12 > gpurun_out/r2f_delta_c4.log 2>&1; echo "rc=$?"; cut -c1-330 gpurun_out/r2f_delta_c4.log | tail -5
timeout 1500 python -m pytest tests -m gpu -x -q -k "not c4" > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2f_tests.log
