set -x
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2al_bench8.json 2> gpurun_out/r2al_bench8.err; echo "bench8 rc=$?"; tail -c 900 gpurun_out/r2al_bench8.json
