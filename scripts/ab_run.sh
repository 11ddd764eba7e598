#!/bin/bash
# A/B of the resident-weight FFN pair kernel against the ring version, then the parity subset on the default
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ffn_tc" --timeout 120 2>&1 | tail -3
AVB_LIB=/root/repo/ab_tmp/lib_trace.so timeout 120 python scripts/trace_ffn.py 2>&1 | grep -A9 "ffn M\|== ln\|== mma\|== epi" | head -44
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "bf16 or stage or golden" --timeout 300 2>&1 | tail -3
for v in 1 0; do
  AVB_FFN_RESIDENT=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']
print('resident=$v', round(d['value'],1), 'out', k['out'], 'ffn', k['ffn1'], 'qkv', k['qkv'], d['clocks']['sm_mhz'])"
done
