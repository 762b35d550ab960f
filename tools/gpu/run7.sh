mkdir -p gpurun_out
V="probe_filter=1,filter_bits=1048576,probe_threads=1024;probe_filter=1,filter_bits=1048576,probe_threads=512;probe_filter=1,filter_bits=786432,probe_threads=512;probe_filter=1,filter_bits=786432,probe_threads=1024;probe_filter=1,filter_bits=901120,probe_threads=512;probe_filter=1,filter_bits=1310720,probe_threads=1024;probe_filter=1,filter_bits=524288,probe_threads=512;probe_filter=1,filter_bits=1744896,probe_threads=1024"
timeout 600 python tools/kbench.py --genome-size 500000000 --what matrix --variants "$V" 2>&1 | tee gpurun_out/kbench_probe3.log | cut -c1-330
