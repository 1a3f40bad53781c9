set -x
timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -q > gpurun_out/r2y_tests.log 2>&1; echo "tests rc=$?"; grep -E "^E |passed|failed" gpurun_out/r2y_tests.log | cut -c1-300 | head -30
timeout 600 python scripts/bench_delta.py c2 60 GRAAL_DELTA_MINB=2 GRAAL_DELTA_MINB=3 GRAAL_DELTA_MINB=4 > gpurun_out/r2y_ab.log 2>&1; echo rc=$?; grep -v Warn gpurun_out/r2y_ab.log | cut -c1-330 | tail -5
