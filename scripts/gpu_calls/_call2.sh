set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "not c4_scale" > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2b_tests.log
timeout 900 python scripts/bench_full.py c2 20 > gpurun_out/r2b_full_c2.log 2>&1; echo "rc=$?"; cat gpurun_out/r2b_full_c2.log | tail -20
timeout 900 python scripts/bench_full.py c4 10 > gpurun_out/r2b_full_c4.log 2>&1; echo "rc=$?"; cat gpurun_out/r2b_full_c4.log | tail -20
