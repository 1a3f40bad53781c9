set -x
V="0:8:4:4 1:8:4:2 1:8:4:4 1:8:3:2 1:8:3:4 1:8:3:8 1:4:4:2 1:4:4:4"
timeout 900 python scripts/bench_full.py c4 10 $V > gpurun_out/r2j_full_c4.log 2>&1; echo "rc=$?"; cut -c1-200 gpurun_out/r2j_full_c4.log | tail -14
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_delta_contacts_rows|k_band_delta_fast" -s 4 -c 3 -o gpurun_out/prof_c2_r2j -f python bench.py --config c2 --profile-only --steps 2 --warmup 2 > gpurun_out/ncu_c2_r2j.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_c2_r2j.log
