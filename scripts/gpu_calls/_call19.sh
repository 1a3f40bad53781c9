set -x
timeout 900 python -m pytest tests/test_gpu_likelihood.py -m gpu -x -q -k degenerate > gpurun_out/r2s_tests_deg.log 2>&1; echo "deg rc=$?"; tail -12 gpurun_out/r2s_tests_deg.log | cut -c1-250
V="0:8:4:2 0:8:4:4 1:8:4:2 1:8:4:4 1:8:5:2 1:8:5:4 1:8:6:2 1:4:4:2 1:4:5:2 1:4:6:2 1:4:6:4"
timeout 900 python scripts/bench_full.py c2 20 $V > gpurun_out/r2s_full_c2.log 2>&1; echo "rc=$?"
timeout 900 python scripts/bench_full.py c4 10 $V > gpurun_out/r2s_full_c4.log 2>&1; echo "rc=$?"
