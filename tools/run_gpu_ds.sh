for ns in 4 6 9; do echo "== B=1 nsplit $ns"; VG_DS_NSPLIT=$ns timeout 300 python tools/decode_bench.py 1 --kind=step 2>&1 | grep -v Warning | cut -c1-80; done
for ns in 2 4; do echo "== B=2 nsplit $ns"; VG_DS_NSPLIT=$ns timeout 300 python tools/decode_bench.py 2 --kind=step 2>&1 | grep -v Warning | cut -c1-80; done
echo "== B=8 warp mode nsplit 8 / 4"; VG_DS_COOP=0 VG_DS_NSPLIT=8 timeout 300 python tools/decode_bench.py 8 --kind=step 2>&1 | grep -v Warning | cut -c1-80
VG_DS_COOP=0 VG_DS_NSPLIT=4 timeout 300 python tools/decode_bench.py 8 --kind=step 2>&1 | grep -v Warning | cut -c1-80
echo "== B=8 direct (no late merge)"; VG_DS_LATE_MERGE=0 timeout 300 python tools/decode_bench.py 8 --kind=step 2>&1 | grep -v Warning | cut -c1-80
