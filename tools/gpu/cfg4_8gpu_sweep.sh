mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 tools/run_sweep.py --config cfg4_15Gbp_wheat --thresholds 10,30,100,300,1000 > gpurun_out/cfg4_8gpu_sweep.json 2> gpurun_out/cfg4_8gpu_sweep.err
tail -2 gpurun_out/cfg4_8gpu_sweep.err | cut -c1-300; grep '^{' gpurun_out/cfg4_8gpu_sweep.json | cut -c1-1800
