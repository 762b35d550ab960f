# usage: bash tools/gpu/multi_gpu_bench.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
tail -2 gpurun_out/r02_bench_${N}gpu.err | cut -c1-300; cut -c1-400 gpurun_out/r02_bench_${N}gpu.json
