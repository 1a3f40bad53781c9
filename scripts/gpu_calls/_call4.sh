set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2d_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2d_bench.err; head -c 6000 gpurun_out/r2d_bench.json
timeout 600 python bench.py --steps 20 --warmup 5 --chains-per-gpu 4 --no-c4 --no-cpu-baseline > gpurun_out/r2d_bench_ch4.json 2> gpurun_out/r2d_bench_ch4.err; echo "bench ch4 rc=$?"; tail -3 gpurun_out/r2d_bench_ch4.err; head -c 1500 gpurun_out/r2d_bench_ch4.json
timeout 900 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/r2d_ref.json 2> gpurun_out/r2d_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/r2d_ref.err; head -c 2500 gpurun_out/r2d_ref.json
