mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/overlap_probe.py 2>&1 | grep -E "alone|ce " 
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -3 gpurun_out/bench3.err
