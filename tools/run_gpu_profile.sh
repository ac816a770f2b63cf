# ncu launch list of ONE eager training step (skip the warm-up launches), for profiles/.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2500 --launch-count 2600 --csv --log-file gpurun_out/launches_eager.csv \
  python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline --no-decode > gpurun_out/ncu_eager.log 2>&1; echo "rc=$?" >> gpurun_out/ncu_eager.log
tail -3 gpurun_out/ncu_eager.log; wc -l gpurun_out/launches_eager.csv
