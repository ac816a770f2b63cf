for v in base new base new; do
if [ $v = base ]; then export VGSLM_LIB=$PWD/vae_gslm_b200/libvgslm_base.so; else unset VGSLM_LIB; fi
echo -n "$v  "; timeout 900 python bench.py --no-decode --no-shapes --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"; done
