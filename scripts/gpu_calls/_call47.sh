set -x
timeout 900 python scripts/bench_delta.py c2 60 GRAAL_DELTA_STAB=0 GRAAL_DELTA_STAB=1 GRAAL_DELTA_STAB=0 GRAAL_DELTA_STAB=1 > gpurun_out/r2at_ab.log 2>&1; echo rc=$?; grep -v Warn gpurun_out/r2at_ab.log | cut -c1-330 | tail -5
