set -x
export GRAAL_WIN_STAB=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 python scripts/sanitize_step.py > gpurun_out/r2_sanitizer_memcheck_c1l2.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_c1l2.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 0 python scripts/sanitize_step.py > gpurun_out/r2_sanitizer_racecheck_c1l2.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_racecheck_c1l2.log
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 0 python scripts/sanitize_step.py > gpurun_out/r2_sanitizer_initcheck_c1l2.log 2>&1; echo "initcheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_initcheck_c1l2.log
