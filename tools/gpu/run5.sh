mkdir -p gpurun_out
python -m pytest tests/test_gpu.py -x -q -m gpu 2>&1 | tail -5
V="part_mode=1;part_mode=2,keep_min=3;part_mode=2,keep_min=3,smem_variant=2;part_mode=2,keep_min=3,smem_variant=1"
timeout 600 python tools/kbench.py --genome-size 500000000 --what count --variants "$V" 2>&1 | tee gpurun_out/kbench_smem3.log | cut -c1-330
V="probe_filter=0;probe_filter=1;probe_filter=1,filter_bits=1048576;probe_filter=1,filter_bits=524288;probe_filter=0,rep_load_inv=8;probe_filter=1,rep_load_inv=2;probe_filter=1,rep_load_inv=4"
timeout 600 python tools/kbench.py --genome-size 500000000 --what matrix --variants "$V" 2>&1 | tee gpurun_out/kbench_probe1.log | cut -c1-330
