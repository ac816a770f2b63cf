# ncu launch list of ONE eager, serial training step of the current build (cold-cache, serialised: compare SHARES)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2500 --launch-count 2600 --csv --log-file gpurun_out/r02_launches_eager.csv python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline --no-decode --no-shapes --no-gpu-reference > gpurun_out/ncu_eager.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_eager.csv | tee gpurun_out/r02_launch_summary.md | head -60
