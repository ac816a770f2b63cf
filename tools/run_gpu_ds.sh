mkdir -p gpurun_out
for args in "--small --batch 3 --prefix 0" "--batch 1 --prefix 2,4,0" "--batch 5 --prefix 0" "--batch 33 --prefix 2,0" "--batch 64 --prefix 0" "--batch 65 --prefix 0" "--batch 130 --prefix 0" "--batch 256 --prefix 0"; do
  echo "=== ds_debug $args"
  timeout 200 python tools/ds_debug.py --mode 1 $args 2>&1 | grep -v "Warning\|return float" | tail -5
done 2>&1 | tee gpurun_out/ds_debug.log
( timeout 120 python tools/ds_trace.py 1 1 2>&1 | grep -v Warn
  timeout 120 python tools/ds_trace.py 64 1 2>&1 | grep -v Warn ) | tee gpurun_out/ds_trace.log
VG_DS_BARRIER=1 timeout 600 python tools/decode_bench.py 1 2 4 8 16 32 64 128 256 --kind=step 2>&1 | grep -v Warning | tee gpurun_out/ds_bench_step.log
