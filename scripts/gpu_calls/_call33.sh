set -x
timeout 1500 python -m pytest tests/test_gpu_sampler.py tests/test_replica.py tests/test_gpu_variants.py -m gpu -x -q > gpurun_out/r2af_tests.log 2>&1; echo "tests rc=$?"; grep -E "^E |passed|failed" gpurun_out/r2af_tests.log | cut -c1-300 | head -20
timeout 900 python bench.py --steps 20 --warmup 5 --no-c4 --no-original --no-cpu-baseline > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2af_bench.err
GRAAL_DEVICE_DRAW=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-c4 --no-original --no-cpu-baseline > gpurun_out/r2af_bench_hostdraw.json 2> gpurun_out/r2af_bench_hostdraw.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 60 --warmup 5 --no-c4 --no-original --no-cpu-baseline > gpurun_out/r2af_bench60.json 2> gpurun_out/r2af_bench60.err; echo "bench rc=$?"
