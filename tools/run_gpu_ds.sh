echo "== coop default"; timeout 600 python tools/decode_bench.py 1 4 8 --kind=step 2>&1 | grep -v Warning | cut -c1-100
echo "== warp mode"; VG_DS_COOP=0 timeout 600 python tools/decode_bench.py 1 4 8 --kind=step 2>&1 | grep -v Warning | cut -c1-100
