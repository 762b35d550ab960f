mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_full.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_count_smem|k_sub_scatter|k_probe_stream_filt|k_tile_scatter|k_extract|k_sub_hist" -s 18 -c 6 -o gpurun_out/r01_bench_full $CMD > gpurun_out/ncu_full3.log 2>&1
tail -2 gpurun_out/ncu_full3.log | cut -c1-200
