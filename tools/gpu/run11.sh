mkdir -p gpurun_out
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -2 gpurun_out/bench_${N}gpu.err | cut -c1-300; cut -c1-200 gpurun_out/bench_${N}gpu.json
done
