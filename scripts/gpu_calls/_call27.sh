set -x
timeout 900 python scripts/bench_full.py c4 20 8:4:4:1 8:4:8:1 8:3:4:1 8:5:4:1 8:6:4:1 4:4:4:1 4:6:4:1 > gpurun_out/r2aa_full_c4.log 2>&1; echo rc=$?; grep -v Warn gpurun_out/r2aa_full_c4.log | cut -c1-200 | tail -9
