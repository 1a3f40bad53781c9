set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests0.log 2>&1; echo "tests rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 0 python __graft_entry__.py smoke > gpurun_out/r2a_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 0 python __graft_entry__.py smoke > gpurun_out/r2a_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
timeout 900 python bench.py --config c4 --steps 20 --warmup 3 > gpurun_out/bench_c4_r2a.json 2> gpurun_out/bench_c4_r2a.err; echo "bench c4 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_full_contacts|k_delta_contacts|k_band_delta" -s 9 -c 8 -o gpurun_out/prof_c4_r2a -f python bench.py --config c4 --profile-only --steps 2 --warmup 1 > gpurun_out/ncu_c4_r2a.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2_tests0.log
tail -5 gpurun_out/r2a_sanitizer_memcheck.log
tail -5 gpurun_out/r2a_sanitizer_racecheck.log
cat gpurun_out/bench_c4_r2a.json | head -c 3000
