mkdir -p gpurun_out
( timeout 120 python tools/ds_trace.py 1 1 2>&1 | grep -v Warn ) | tee gpurun_out/ds_trace.log
