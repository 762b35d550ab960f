mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/t2.log
cat gpurun_out/t2.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
tail -3 gpurun_out/bench2.err; cat gpurun_out/bench2.json
