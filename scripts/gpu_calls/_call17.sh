set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2q_tests.log | cut -c1-300
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2q_bench.err
timeout 900 python bench.py --steps 60 --warmup 5 --no-c4 --no-original --no-cpu-baseline > gpurun_out/r2q_bench60.json 2> gpurun_out/r2q_bench60.err; echo "bench rc=$?"
