set -x
nvidia-smi --query-gpu=index,name --format=csv | head -10
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2t_bench_n8.json 2> gpurun_out/r2t_bench_n8.err; echo "bench n8 rc=$?"; tail -3 gpurun_out/r2t_bench_n8.err; tail -c 1200 gpurun_out/r2t_bench_n8.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --impl reference --steps 3 --warmup 1 > gpurun_out/r2t_ref_n8.json 2> gpurun_out/r2t_ref_n8.err; echo "ref n8 rc=$?"; tail -c 600 gpurun_out/r2t_ref_n8.json
