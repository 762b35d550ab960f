mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/run_config.py --config cfg3_4.5Gbp_tobacco --low-memory --k-list 23,15,31 > gpurun_out/cfg3_smem3.log 2>&1; grep '^{' gpurun_out/cfg3_smem3.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['k'], d['pass0_s'], d['pass1_s'], d['stats']['distinct'], d['stats']['n_selected'], d['stats']['nnz'], d['stage_ms'])"
