set -x
timeout 600 python scripts/profile_host_step.py > gpurun_out/r2ah_hostprof.log 2>&1; echo rc=$?; grep -v Warn gpurun_out/r2ah_hostprof.log | head -45 | cut -c1-160
