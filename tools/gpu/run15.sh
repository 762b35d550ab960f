mkdir -p gpurun_out
python -m pytest tests/test_gpu.py -x -q -m gpu -k "shared_memory or partitioned or skewed" 2>&1 | tail -2
V="part_mode=2,keep_min=3;part_mode=2,keep_min=3,smem_pb=8;part_mode=2,keep_min=3,smem_pb=10;part_mode=1"
timeout 600 python tools/kbench.py --genome-size 500000000 --what count --variants "$V" 2>&1 | tee gpurun_out/kbench_smem5.log | cut -c1-250
python tools/run_config.py --config cfg3_4.5Gbp_tobacco --low-memory > gpurun_out/cfg3_smem4.log 2>&1; grep '^{' gpurun_out/cfg3_smem4.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['k'], d['pass0_s'], d['pass1_s'], d['stats']['distinct'], d['stats']['n_selected'], d['stats']['nnz'], d['stage_ms'])"
