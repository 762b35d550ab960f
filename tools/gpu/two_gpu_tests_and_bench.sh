mkdir -p gpurun_out
python -m pytest tests/test_distributed_gpu.py -x -q -m gpu > gpurun_out/t_dist.log 2>&1; tail -3 gpurun_out/t_dist.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -3 gpurun_out/bench_2gpu.err; cut -c1-300 gpurun_out/bench_2gpu.json
