set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2am_tests.log 2>&1; echo "tests rc=$?"; grep -E "^E |passed|failed" gpurun_out/r2am_tests.log | cut -c1-300 | head -20
rm -f gpurun_out/r2am_ab.log
for v in "GRAAL_DEVICE_DRAW=1" "GRAAL_DEVICE_DRAW=0" "GRAAL_DEVICE_DRAW=1"; do
  env $v timeout 900 python bench.py --steps 40 --warmup 5 --no-c4 --no-original --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$v', 'value %.0f (%.4f ms) e2e %.0f (%.4f ms) launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches']))
" >> gpurun_out/r2am_ab.log
done
cat gpurun_out/r2am_ab.log
