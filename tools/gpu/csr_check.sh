mkdir -p gpurun_out
python -m pytest tests/test_gpu.py -x -q -m gpu -k "rows_to_csr or full_path or probe_filters or degenerate or thresholds or cfg2" 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_csr.json 2> gpurun_out/bench_csr.err; tail -2 gpurun_out/bench_csr.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_csr.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['timed_attempts'])
print({k:round(v,3) for k,v in d['roofline']['stage_ms'].items() if v>0.01})
PY
