# Documentation run on one B200: bench lines of every workload, the ncu launch lists (tf32 / f16) of one forward and full captures of
# the attention and fused Mlp-tail kernels (exported to CSV on the box: gpurun_out/ is limited to 64 MiB).
python bench.py > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err
tail -c 200 gpurun_out/r02_bench7.err
for w in lite_1080p base_4k base_vimeo_b32 stream_1080p; do python bench.py --workload $w --no-cpu --no-extra > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; tail -c 200 gpurun_out/r02_bench_$w.err; done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_tf32_1080p.csv python tools/one_forward.py base 1088 1920 tf32 > gpurun_out/ncu_l1.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_f16_1080p.csv python tools/one_forward.py base 1088 1920 f16 > gpurun_out/ncu_l2.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:window_attention_tc -c 6 -o /tmp/att -f python tools/one_forward.py base 1088 1920 tf32 > gpurun_out/ncu_l3.log 2>&1
ncu -i /tmp/att.ncu-rep --page raw --csv > gpurun_out/r02_attention_tf32_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:mlp_tail -c 6 -o /tmp/mt -f python tools/one_forward.py base 1088 1920 tf32 > gpurun_out/ncu_l4.log 2>&1
ncu -i /tmp/mt.ncu-rep --page raw --csv > gpurun_out/r02_mlp_tail_tf32_raw.csv 2>/dev/null
du -sh gpurun_out
