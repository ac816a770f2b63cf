# ncu launch list of ONE eager, serial training step of the current build (cold-cache, serialised: compare SHARES).
# bench.py brackets its instrumented serial step with spin_kernel markers when VG_BENCH_MARK=1.
mkdir -p gpurun_out
VG_BENCH_MARK=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_eager.csv python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline --no-decode --no-shapes --no-gpu-reference > gpurun_out/ncu_eager.log 2>&1
grep -c spin_kernel gpurun_out/r02_launches_eager.csv
python tools/launch_summary.py gpurun_out/r02_launches_eager.csv | tee gpurun_out/r02_launch_summary.md | head -70
