# Round-2 evidence: GPU test-suite, smoke, both bench arms, launch list, one ncu --set full capture of every hot kernel of a
# bench step and of the Gram kernel, Gram bench.  Everything lands in gpurun_out/ (copied into profiles/ by hand).
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tee gpurun_out/r02_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_smoke.txt
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -1 gpurun_out/r02_bench_1gpu.err | cut -c1-200
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; cut -c1-300 gpurun_out/r02_bench_reference_arm.json
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
$CMD > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_default.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-200
PCMD="python tools/prof_step.py --genome-size 500000000 --passes 2"
$PCMD > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on \
  -k regex:'k_count_skm|k_probe_roll_filt|k_row_bitmap_csr|k_skm_extract|k_rec_scatter|k_tfidf|k_pack|k_select_list|k_count_rec_table' \
  --launch-skip 10 --launch-count 10 -o gpurun_out/r02_full_step $PCMD > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-200
python tools/gram_bench.py > gpurun_out/r02_gram_bench.log 2>&1; grep '^{' gpurun_out/r02_gram_bench.log | cut -c1-600
ncu --set full --clock-control none -k regex:k_gram_tf32 --launch-skip 10 --launch-count 1 -o gpurun_out/r02_gram python tools/gram_bench.py > gpurun_out/ncu_gram.log 2>&1; tail -1 gpurun_out/ncu_gram.log | cut -c1-200
