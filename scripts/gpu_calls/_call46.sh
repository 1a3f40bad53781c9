set -x
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2as_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2as_smoke.log | cut -c1-300
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2as_tests.log 2>&1; echo "tests rc=$?"; grep -E "^E |passed|failed" gpurun_out/r2as_tests.log | cut -c1-300 | head
timeout 1200 python bench.py > gpurun_out/r2as_bench.json 2> gpurun_out/r2as_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2as_bench.err | cut -c1-200
