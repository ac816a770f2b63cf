# final-build evidence of round 2: GPU tests, the bench line, ncu launch lists (training step, decode steps) and ncu --set full
# captures of the decode kernels.  Outputs under gpurun_out/ (scratch); summaries are copied into profiles/ by hand.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu_all.log 2>&1; tail -3 gpurun_out/pytest_gpu_all.log
timeout 600 python tools/decode_bench.py 8 16 24 --kind=linear 2>&1 | grep -v Warning | tee gpurun_out/ds_bench_linear_small.log
VG_BENCH_GEMM_TABLE=1 timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
bash tools/run_gpu_profile.sh > /dev/null 2>&1; head -12 gpurun_out/r02_launch_summary.md
for cfg in "256 layerwise" "64 linear" "1 step"; do
  set -- $cfg
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dec_launches_$1_$2.csv python tools/decode_profile.py $1 $2 400 > gpurun_out/dec_prof_$1_$2.log 2>&1
  python tools/decode_launch_summary.py gpurun_out/dec_launches_$1_$2.csv > gpurun_out/dec_summary_$1_$2.md
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_decode_stream --launch-skip 40 --launch-count 1 -o gpurun_out/prof_ad_stream -f python tools/attn_decode_bench.py 402 256 > gpurun_out/ncu_ad.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_decode_stream --launch-skip 40 --launch-count 1 -o gpurun_out/prof_ad_stream_b64 -f python tools/attn_decode_bench.py 402 64 > gpurun_out/ncu_ad64.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:decode_step_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/prof_decode_step_b1 -f python tools/decode_profile.py 1 step 400 > gpurun_out/ncu_ds1.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:decode_linear_kernel --launch-skip 200 --launch-count 1 -o gpurun_out/prof_decode_linear_b64 -f python tools/decode_profile.py 64 linear 400 > gpurun_out/ncu_dl64.log 2>&1
timeout 200 python tools/attn_decode_bench.py 402 1 8 32 64 128 256 2>&1 | grep "splits=None" | tee gpurun_out/ad_bench_final.log
ls -la gpurun_out/*.ncu-rep | tail -5
