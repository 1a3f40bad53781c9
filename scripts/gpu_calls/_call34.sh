set -x
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2ag_tests.log 2>&1; echo "tests rc=$?"; grep -E "^E |passed|failed" gpurun_out/r2ag_tests.log | cut -c1-300 | head -20
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ag_bench.json 2> gpurun_out/r2ag_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2ag_bench.err | cut -c1-300
timeout 900 python bench.py --chains-per-gpu 4 --steps 20 --warmup 5 --no-c4 --no-original --no-cpu-baseline > gpurun_out/r2ag_bench_4chains.json 2> gpurun_out/r2ag_bench_4chains.err; echo "bench4 rc=$?"; tail -2 gpurun_out/r2ag_bench_4chains.err | cut -c1-300
