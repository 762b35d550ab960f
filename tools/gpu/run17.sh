mkdir -p gpurun_out
CMD="python tools/prof_step.py --genome-size 200000000 --passes 1"
$CMD > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_count_smem|k_sub_scatter|k_sub_hist|k_probe_stream_filt|k_tile_scatter|k_extract|k_row_sort|k_row_emit|k_tfidf|k_pack" -c 10 -o gpurun_out/r01_final_full $CMD > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/prof_plain2.log | cut -c1-600; tail -3 gpurun_out/ncu_full2.log | cut -c1-300
