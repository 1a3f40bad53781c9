set -x
timeout 900 python -m pytest tests/test_gpu_mh.py -m gpu -x -q > gpurun_out/r2o_tests_mh.log 2>&1; echo "mh rc=$?"; tail -12 gpurun_out/r2o_tests_mh.log | cut -c1-400
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 0 python __graft_entry__.py smoke > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 0 python __graft_entry__.py smoke > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 python scripts/sanitize_step.py > gpurun_out/r2_sanitizer_memcheck_c1l2.log 2>&1; echo "memcheck c1l2 rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_c1l2.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 0 python scripts/sanitize_step.py > gpurun_out/r2_sanitizer_racecheck_c1l2.log 2>&1; echo "racecheck c1l2 rc=$?"; tail -4 gpurun_out/r2_sanitizer_racecheck_c1l2.log
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 0 python scripts/sanitize_step.py > gpurun_out/r2_sanitizer_initcheck_c1l2.log 2>&1; echo "initcheck c1l2 rc=$?"; tail -4 gpurun_out/r2_sanitizer_initcheck_c1l2.log
