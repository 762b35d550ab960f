V="part_mode=2,keep_min=3,hash_mode=0;part_mode=2,keep_min=3,hash_mode=1"
timeout 600 python tools/kbench.py --genome-size 500000000 --what count --variants "$V" 2>&1 | cut -c1-400
