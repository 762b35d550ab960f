#!/usr/bin/env python
"""Readable digest of a bench.py JSON line (argv[1])."""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print("N=%d value %.2f Gbp/s  %.2f ms/step | e2e %.2f Gbp/s %.2f ms/step | attempts %s" % (
    d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'] if d.get('e2e') else -1,
    d['e2e']['ms_per_step'] if d.get('e2e') else -1, [a['ms_per_step'] for a in d.get('timed_attempts', [])]))
print("parity", d.get('parity_check') and d['parity_check'].get('ok'), "| roofline", d['roofline']['kernel'][:40], round(d['roofline']['frac'], 3))
print("stage_ms", {k: round(v, 2) for k, v in d['roofline']['stage_ms'].items() if v > 0.01})
if d['result'].get('phase_ms_rank0'):
    print("phases", d['result']['phase_ms_rank0'])
if d.get('iso_work'):
    i = d['iso_work']; print("iso_work: %.2f ms/step %.2f Gbp/s n_selected %s" % (i['ms_per_step'], i['value'], i['n_selected']))
print("n_selected", d['result']['n_selected'], "nnz_rank0", d['result']['nnz_rank0'], "overflowed", d['result']['overflowed_sub_buckets'], "numa", d.get('numa_binding'))
