#!/bin/bash
for v in tok base tok base; do
  if [ $v == tok ]; then export AVB_LIB=/root/repo/ab_tmp/lib_tok.so; else unset AVB_LIB; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']
print('$v', round(d['value'],1), 'out', k['out'], 'ffn', k['ffn1'], 'qkv', k['qkv'], d['clocks']['sm_mhz'])"
done
export AVB_LIB=/root/repo/ab_tmp/lib_tok.so
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "bf16 or stage or golden" --timeout 300 2>&1 | tail -3
