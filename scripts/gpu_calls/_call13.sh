set -x
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2m_bench_n2.err; tail -c 1500 gpurun_out/r2m_bench_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 120 --warmup 5 --exchange-every 20 --chains-per-gpu 2 > gpurun_out/r2m_bench_n2c2.json 2> gpurun_out/r2m_bench_n2c2.err; echo "bench n2c2 rc=$?"; tail -3 gpurun_out/r2m_bench_n2c2.err; tail -c 1500 gpurun_out/r2m_bench_n2c2.json
