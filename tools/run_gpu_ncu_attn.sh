mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd_kernel -s 4 -c 1 -f -o gpurun_out/prof_attn_bwd_v2 python tools/attn_raw.py 8 1000 16 > gpurun_out/ncu_attn_bwd.log 2>&1; tail -2 gpurun_out/ncu_attn_bwd.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd_kernel -s 4 -c 1 -f -o gpurun_out/prof_attn_fwd_v2 python tools/attn_raw.py 8 1000 16 > gpurun_out/ncu_attn_fwd.log 2>&1; tail -2 gpurun_out/ncu_attn_fwd.log
