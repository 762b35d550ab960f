mkdir -p gpurun_out
python -m pytest tests/test_gpu.py -x -q -m gpu -k "shared_memory or partitioned or counts_bit_exact or skewed" 2>&1 | tail -5
V="part_mode=1;part_mode=2,smem_variant=0,keep_min=1;part_mode=2,smem_variant=0,keep_min=3;part_mode=2,smem_variant=4,keep_min=3;part_mode=2,smem_variant=3,keep_min=3;part_mode=2,smem_variant=1,keep_min=3;part_mode=2,smem_variant=2,keep_min=3;part_mode=2,smem_variant=0,keep_min=3,smem_target=4096;part_mode=2,smem_variant=0,keep_min=3,smem_target=5000;part_mode=2,smem_variant=0,keep_min=3,smem_target=0,hash_mode=1"
timeout 600 python tools/kbench.py --genome-size 500000000 --what count --variants "$V" 2>&1 | tee gpurun_out/kbench_smem2.log | cut -c1-330
