set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2ai_tests.log 2>&1; echo "tests rc=$?"; grep -E "^E |passed|failed" gpurun_out/r2ai_tests.log | cut -c1-300 | head -20
timeout 900 python bench.py --steps 20 --warmup 5 --no-c4 --no-original --no-cpu-baseline > gpurun_out/r2ai_bench.json 2> gpurun_out/r2ai_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2ai_bench.err
GRAAL_PREPARE_PROLOGUE=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-c4 --no-original --no-cpu-baseline > gpurun_out/r2ai_bench_noprep.json 2> gpurun_out/r2ai_bench_noprep.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 60 --warmup 5 --no-c4 --no-original --no-cpu-baseline > gpurun_out/r2ai_bench60.json 2> gpurun_out/r2ai_bench60.err; echo "bench rc=$?"
