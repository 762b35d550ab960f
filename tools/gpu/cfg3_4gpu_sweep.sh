mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 tools/run_sweep.py --config cfg3_4.5Gbp_tobacco --thresholds 100,30,300,1000 > gpurun_out/cfg3_4gpu_sweep.json 2> gpurun_out/cfg3_4gpu_sweep.err
tail -2 gpurun_out/cfg3_4gpu_sweep.err | cut -c1-300; grep '^{' gpurun_out/cfg3_4gpu_sweep.json | cut -c1-1500
