mkdir -p gpurun_out
CMD="python tools/prof_step.py --genome-size 200000000 --passes 1"
$CMD > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_count_smem|k_sub_scatter|k_sub_hist|k_probe_stream|k_tile_scatter|k_extract|k_row_sort" -c 7 -o gpurun_out/r01_smem_full $CMD > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/prof_plain.log; tail -5 gpurun_out/ncu_full.log
