#!/bin/bash
# A/B of whole-step throughput: bench.py with each library in ab_tmp/, interleaved, same box (power-capped clocks differ per box)
mkdir -p gpurun_out
{
for rep in 1 2; do for lib in "$@"; do
  AVB_LIB=$PWD/ab_tmp/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        j = json.loads(l)
        print('$lib', round(j['value'], 1), 'ds/s', j['ms_per_step'], 'ms', j.get('clocks'), j.get('kernel_ms_per_step'))
"
done; done
} > gpurun_out/bench_ab.log 2>&1
cat gpurun_out/bench_ab.log
