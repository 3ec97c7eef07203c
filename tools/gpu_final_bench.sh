#!/bin/bash
# Round-2 evidence, part B (after tools/make_traffic.py): every bench line on the final build.
O=gpurun_out/final_bench; mkdir -p $O; rm -f $O/*
sha256sum dantzig_b200/libdantzig_b200.so | cut -c1-16 > $O/lib_sha16.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err
for wl in c2 c3 c4; do timeout 400 python bench.py --workload $wl --steps 3 --warmup 3 > $O/bench_$wl.json 2> $O/bench_$wl.err; done
for wl in c2 c5; do timeout 400 python bench.py --workload $wl --numerics fast --steps 3 --warmup 3 > $O/bench_${wl}_fast.json 2> $O/bench_${wl}_fast.err; done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_c5_reference.json 2> $O/bench_c5_reference.err
for f in $O/*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(' ', d.get('value'), d.get('unit'), 'e2e', d.get('e2e',{}).get('value'), 'frac', d.get('roofline',{}).get('frac'), 'traffic', d.get('roofline',{}).get('traffic'), 'parity', d.get('parity'))
"; done
tail -2 $O/*.err | tail -20
