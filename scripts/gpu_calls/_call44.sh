set -x
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --config c5 --chains-per-gpu 8 --steps 6 --warmup 3 > gpurun_out/r2aq_c5_8x8.json 2> gpurun_out/r2aq_c5_8x8.err; echo "c5 rc=$?"; tail -c 400 gpurun_out/r2aq_c5_8x8.json
