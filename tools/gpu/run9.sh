mkdir -p gpurun_out
python -m pytest tests/test_gpu.py -x -q -m gpu -k "shared_memory" 2>&1 | tail -3
V="part_mode=2,keep_min=3,scatter_variant=0;part_mode=2,keep_min=3,scatter_variant=1;part_mode=2,keep_min=3,scatter_variant=1,smem_target=4096;part_mode=2,keep_min=3,scatter_variant=0,smem_target=4096"
timeout 600 python tools/kbench.py --genome-size 500000000 --what count --variants "$V" 2>&1 | tee gpurun_out/kbench_smem4.log | cut -c1-250
