set -x
timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -q > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"; grep -E "^E |passed|failed" gpurun_out/r2z_tests.log | cut -c1-300 | head -20
timeout 600 python scripts/bench_full.py c2 30 8:4:2:0 8:4:2:1 8:4:4:0 8:4:4:1 > gpurun_out/r2z_full_c2.log 2>&1; echo rc=$?; grep -v Warn gpurun_out/r2z_full_c2.log | cut -c1-400 | tail -5
timeout 900 python scripts/bench_full.py c4 20 8:4:2:0 8:4:2:1 8:4:4:0 8:4:4:1 > gpurun_out/r2z_full_c4.log 2>&1; echo rc=$?; grep -v Warn gpurun_out/r2z_full_c4.log | cut -c1-400 | tail -5
