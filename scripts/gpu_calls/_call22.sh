set -x
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2v_tests.log | cut -c1-300
