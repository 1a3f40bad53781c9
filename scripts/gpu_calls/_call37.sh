set -x
timeout 900 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_variants.py tests/test_gpu_mh.py -m gpu -x -q > gpurun_out/r2aj_tests.log 2>&1; echo "tests rc=$?"; grep -E "^E |passed|failed" gpurun_out/r2aj_tests.log | cut -c1-300 | head -20
for v in "GRAAL_PUBLISH=1 GRAAL_PREPARE_PROLOGUE=1" "GRAAL_PUBLISH=0 GRAAL_PREPARE_PROLOGUE=0" "GRAAL_PUBLISH=1 GRAAL_PREPARE_PROLOGUE=0" "GRAAL_PUBLISH=1 GRAAL_PREPARE_PROLOGUE=1"; do
  env $v timeout 900 python bench.py --steps 40 --warmup 5 --no-c4 --no-original --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$v', 'value %.0f (%.4f ms) e2e %.0f (%.4f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
" >> gpurun_out/r2aj_ab.log
done
cat gpurun_out/r2aj_ab.log
