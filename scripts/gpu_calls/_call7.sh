set -x
V="GRAAL_DELTA_REL=0,GRAAL_BAND_FAST=0 GRAAL_DELTA_REL=1,GRAAL_BAND_FAST=0 GRAAL_DELTA_REL=0,GRAAL_BAND_FAST=1 GRAAL_DELTA_REL=1,GRAAL_BAND_FAST=1"
timeout 900 python scripts/bench_delta.py c2 30 $V > gpurun_out/r2g_delta_c2.log 2>&1; echo "rc=$?"; cut -c1-330 gpurun_out/r2g_delta_c2.log | tail -5
timeout 900 python scripts/bench_delta.py c4 12 $V > gpurun_out/r2g_delta_c4.log 2>&1; echo "rc=$?"; cut -c1-330 gpurun_out/r2g_delta_c4.log | tail -5
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2g_tests.log
