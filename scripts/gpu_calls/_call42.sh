set -x
timeout 900 python -m pytest tests/test_gpu_simulation.py tests/test_simulation_host.py -x -q > gpurun_out/r2ao_sim.log 2>&1; echo "rc=$?"; grep -E "^E |passed|failed|Error" gpurun_out/r2ao_sim.log | cut -c1-300 | head -20
