#!/bin/bash
# A/B of an experimental FFN build (ab_tmp/lib_lne2.so) against the default library
export AVB_LIB=/root/repo/ab_tmp/lib_lne2.so
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ffn_tc" --timeout 120 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "bf16 or stage or golden" --timeout 300 2>&1 | tail -3
AVB_LIB=/root/repo/ab_tmp/lib_trace.so timeout 120 python scripts/trace_ffn.py 2>&1 | grep -A5 "ffn M\|== ln\|== mma\|== epi" | head -30
for v in lne2 base lne2 base; do
  if [ $v == lne2 ]; then export AVB_LIB=/root/repo/ab_tmp/lib_lne2.so; else unset AVB_LIB; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']
print('$v', round(d['value'],1), 'out', k['out'], 'ffn', k['ffn1'], 'qkv', k['qkv'], d['clocks']['sm_mhz'])"
done
