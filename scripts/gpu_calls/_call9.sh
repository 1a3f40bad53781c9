set -x
V="0:8:4:2 0:8:4:4 1:4:4:2 1:4:4:4 1:4:3:2 1:4:3:4 1:8:4:2 1:8:3:2 1:8:3:4 1:2:4:2 1:4:5:2"
timeout 900 python scripts/bench_full.py c2 20 $V > gpurun_out/r2i_full_c2.log 2>&1; echo "rc=$?"; cut -c1-200 gpurun_out/r2i_full_c2.log | tail -14
timeout 900 python scripts/bench_full.py c4 10 $V > gpurun_out/r2i_full_c4.log 2>&1; echo "rc=$?"; cut -c1-200 gpurun_out/r2i_full_c4.log | tail -14
