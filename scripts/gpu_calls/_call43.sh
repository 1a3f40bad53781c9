set -x
timeout 900 python bench.py --config c5 --chains-per-gpu 8 --steps 6 --warmup 3 > gpurun_out/r2ap_c5.json 2> gpurun_out/r2ap_c5.err; echo "c5 rc=$?"; tail -c 600 gpurun_out/r2ap_c5.json; tail -2 gpurun_out/r2ap_c5.err
