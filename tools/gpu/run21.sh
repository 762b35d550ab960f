V="part_mode=2,keep_min=3,scatter_tile=8192;part_mode=2,keep_min=3,scatter_tile=4096;part_mode=2,keep_min=3,scatter_tile=4096,smem_pb=8"
timeout 600 python tools/kbench.py --genome-size 500000000 --what count --variants "$V" 2>&1 | cut -c1-230
