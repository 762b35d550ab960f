python -m pytest tests/test_gpu.py -x -q -m gpu 2>&1 | tail -2
V="part_mode=2,keep_min=3,scatter_pack=0;part_mode=2,keep_min=3,scatter_pack=1;part_mode=2,keep_min=3,scatter_pack=1,scatter_tile=4096"
timeout 600 python tools/kbench.py --genome-size 500000000 --what count --variants "$V" 2>&1 | cut -c1-230
