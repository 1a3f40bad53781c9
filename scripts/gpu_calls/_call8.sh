set -x
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_full_contacts_win|k_delta_contacts_rows|k_band_delta_fast" -s 12 -c 4 -o gpurun_out/prof_c2_r2h -f python bench.py --config c2 --profile-only --steps 2 --warmup 2 > gpurun_out/ncu_c2_r2h.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_c2_r2h.log
timeout 900 python -m pytest tests/test_gpu_original_kernels.py -m gpu -x -q -s > gpurun_out/r2h_tests_orig.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2h_tests_orig.log
