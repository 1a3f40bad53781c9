set -x
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2ak_tests.log 2>&1; echo "tests rc=$?"; grep -E "^E |passed|failed" gpurun_out/r2ak_tests.log | cut -c1-300 | head -20
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ak_bench.json 2> gpurun_out/r2ak_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2ak_bench.err | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 0 python scripts/sanitize_step.py > gpurun_out/r2_sanitizer_memcheck_c1l2.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2_sanitizer_memcheck_c1l2.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 0 python scripts/sanitize_step.py > gpurun_out/r2_sanitizer_racecheck_c1l2.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2_sanitizer_racecheck_c1l2.log
