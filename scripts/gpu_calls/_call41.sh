set -x
GRAAL_DEVICE_DRAW=0 timeout 900 python -m pytest tests/test_gpu_sampler.py tests/test_replica.py -m gpu -x -q > gpurun_out/r2an_hostdraw.log 2>&1; echo "hostdraw rc=$?"; tail -1 gpurun_out/r2an_hostdraw.log
GRAAL_PUBLISH=0 timeout 900 python -m pytest tests/test_gpu_sampler.py -m gpu -x -q > gpurun_out/r2an_nopublish.log 2>&1; echo "nopublish rc=$?"; tail -1 gpurun_out/r2an_nopublish.log
GRAAL_WIN_STAB=1 timeout 900 python -m pytest tests/test_gpu_likelihood.py tests/test_gpu_sampler.py tests/test_gpu_bench_configs.py -m gpu -x -q > gpurun_out/r2an_stab.log 2>&1; echo "stab rc=$?"; tail -1 gpurun_out/r2an_stab.log
GRAAL_LANES=1 GRAAL_GRAPHS=0 timeout 900 python -m pytest tests/test_gpu_sampler.py -m gpu -x -q > gpurun_out/r2an_serial.log 2>&1; echo "serial rc=$?"; tail -1 gpurun_out/r2an_serial.log
