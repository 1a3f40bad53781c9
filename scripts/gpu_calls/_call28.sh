set -x
timeout 900 python scripts/bench_delta.py c2 60 GRAAL_DELTA_PIPE=0 GRAAL_DELTA_PIPE=1 GRAAL_DELTA_GROUPS=2 GRAAL_DELTA_GROUPS=3 GRAAL_DELTA_GROUPS=2,GRAAL_DELTA_GROUP_WAVES=2 > gpurun_out/r2ab_ab.log 2>&1; echo rc=$?; grep -v Warn gpurun_out/r2ab_ab.log | cut -c1-330 | tail -6
