for p in 0 1 0 1; do echo -n "VG_TRAIN_PDL=$p  "; VG_TRAIN_PDL=$p timeout 900 python bench.py --no-decode --no-shapes --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"; done
