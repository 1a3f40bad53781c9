set -x
timeout 900 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_moves.py tests/test_gpu_mh.py -m gpu -x -q > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2p_tests.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 --no-c4 --no-original --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2p_bench.err; head -c 700 gpurun_out/r2p_bench.json
GRAAL_FUSED_PROLOGUE=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-c4 --no-original --no-cpu-baseline > gpurun_out/r2p_bench_nofuse.json 2> gpurun_out/r2p_bench_nofuse.err; echo "bench rc=$?"; head -c 700 gpurun_out/r2p_bench_nofuse.json
timeout 900 python bench.py --steps 60 --warmup 5 --no-c4 --no-original --no-cpu-baseline > gpurun_out/r2p_bench60.json 2> gpurun_out/r2p_bench60.err; echo "bench rc=$?"; head -c 700 gpurun_out/r2p_bench60.json
