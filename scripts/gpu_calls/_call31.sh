set -x
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2ae_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2ae_tests.log | cut -c1-300
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2ae_bench.err | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2ae_ref.json 2> gpurun_out/r2ae_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/launches_r2_final2.csv python bench.py --config c2 --profile-only --steps 3 --warmup 3 > gpurun_out/ncu_launches_r2b.log 2>&1; echo "ncu launches rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_full_contacts_win" -s 6 -c 3 -o gpurun_out/prof_c4_r2_stab -f python bench.py --config c4 --profile-only --steps 2 --warmup 2 > gpurun_out/ncu_c4_r2_stab.log 2>&1; echo "ncu c4 rc=$?"; tail -2 gpurun_out/ncu_c4_r2_stab.log
