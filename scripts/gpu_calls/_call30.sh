set -x
export GRAAL_DELTA_BATCHED=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_delta_contacts -s 9 -c 4 -o gpurun_out/prof_delta_batched_r2 -f python bench.py --steps 3 --warmup 3 --no-c4 --no-original --no-cpu-baseline > gpurun_out/r2ad_ncu.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2ad_ncu.log | cut -c1-200
ls -la gpurun_out/prof_delta_batched_r2.ncu-rep
