set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 171 --csv --log-file gpurun_out/launches_r2_e2e.csv python scripts/profile_host_step.py > gpurun_out/ncu_launches_e2e.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_launches_e2e.log | cut -c1-200
