mkdir -p gpurun_out
for PB in 10 9 12; do
python tools/run_config.py --config cfg3_4.5Gbp_tobacco --low-memory --tune smem_pb=$PB > gpurun_out/cfg3_pb$PB.log 2>&1; tail -1 gpurun_out/cfg3_pb$PB.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print($PB, d['pass1_s'], {k:v for k,v in d['stage_ms'].items() if k in ('part_hist','part_scatter','part2','count')})"
done
