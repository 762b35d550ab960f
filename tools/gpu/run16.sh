mkdir -p gpurun_out
N=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu_b.json 2> gpurun_out/bench_${N}gpu_b.err
tail -2 gpurun_out/bench_${N}gpu_b.err | cut -c1-300; cut -c1-200 gpurun_out/bench_${N}gpu_b.json
