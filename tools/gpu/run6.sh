mkdir -p gpurun_out
python -m pytest tests/test_gpu.py -x -q -m gpu -k "full_path or degenerate or skewed or known" 2>&1 | tail -3
V="probe_filter=0;probe_filter=1;probe_filter=1,filter_bits=1048576;probe_filter=0,probe_batch=8"
timeout 600 python tools/kbench.py --genome-size 500000000 --what matrix --variants "$V" 2>&1 | tee gpurun_out/kbench_probe2.log | cut -c1-330
