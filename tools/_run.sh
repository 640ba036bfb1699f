mkdir -p gpurun_out
nvidia-smi -L
echo "== multi_gpu_check world=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py 2>&1 | tail -8
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_c2_n2.json 2> gpurun_out/r2_bench_c2_n2.err
tail -c 3000 gpurun_out/r2_bench_c2_n2.json; tail -5 gpurun_out/r2_bench_c2_n2.err
