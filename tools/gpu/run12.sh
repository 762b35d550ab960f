mkdir -p gpurun_out
python -m pytest tests/test_gpu.py -x -q -m gpu -k "probe_filters or full_path or skewed" 2>&1 | tail -3
V="probe_filter=1;probe_filter=0;probe_filter=2;probe_filter=2,probe_fp=1"
timeout 600 python tools/kbench.py --genome-size 500000000 --what matrix --variants "$V" 2>&1 | tee gpurun_out/kbench_probe4.log | cut -c1-330
python tools/run_config.py --config cfg3_4.5Gbp_tobacco --low-memory > gpurun_out/cfg3_smem2.log 2>&1; tail -1 gpurun_out/cfg3_smem2.log | cut -c1-1200
