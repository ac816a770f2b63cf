timeout 600 python tools/decode_bench.py 80 96 128 --kind=linear --force-engine 2>&1 | grep -v Warning | cut -c1-130
