set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2l_tests.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2l_bench.err
timeout 1500 python bench.py --config c5 --chains-per-gpu 8 --steps 6 --warmup 3 > gpurun_out/r2l_bench_c5.json 2> gpurun_out/r2l_bench_c5.err; echo "bench c5 rc=$?"; tail -5 gpurun_out/r2l_bench_c5.err; nvidia-smi --query-gpu=memory.used --format=csv
