for p in 0 1; do echo "== VG_TRAIN_PDL=$p"; VG_TRAIN_PDL=$p timeout 900 python bench.py --no-decode --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'], d['ms_per_step'], r['achieved'], d['loss'], {k:(v['value'],v['ms_per_step']) for k,v in d['shapes'].items()})"; done
VG_TRAIN_PDL=1 timeout 600 python -m pytest tests/test_model_gpu.py -q -x -k "train_step or gradient_accumulation or first_writer" 2>&1 | tail -2
