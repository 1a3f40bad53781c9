set -x
timeout 900 python -m pytest tests/test_gpu_mh.py -m gpu -x -q > gpurun_out/r2n_tests_mh.log 2>&1; echo "mh rc=$?"; tail -30 gpurun_out/r2n_tests_mh.log
timeout 900 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_likelihood.py tests/test_gpu_moves.py -m gpu -x -q -k "not c4" > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2n_tests.log
