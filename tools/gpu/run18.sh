mkdir -p gpurun_out
python -m pytest tests/test_gpu.py -x -q -m gpu 2>&1 | tail -3
V="part_mode=2,keep_min=3,smem_variant=0;part_mode=2,keep_min=3,smem_variant=4;part_mode=2,keep_min=3,smem_variant=1;part_mode=2,keep_min=3,smem_variant=2;part_mode=2,keep_min=1,smem_variant=0"
timeout 600 python tools/kbench.py --genome-size 500000000 --what count --variants "$V" 2>&1 | tee gpurun_out/kbench_smem6.log | cut -c1-230
V="probe_filter=1;probe_filter=2;probe_filter=0"
timeout 600 python tools/kbench.py --genome-size 500000000 --what matrix --variants "$V" 2>&1 | tee gpurun_out/kbench_probe5.log | cut -c1-330
