#!/bin/bash
# A/B of an experimental build against the default library on the GPU box.
#   build here:  AVB_LIB=/root/repo/ab_tmp/lib_x.so AVB_NVCC_EXTRA="-DAVB_FFN_DECOUPLE=1" python -c "import avici_b200._lib as l; l.build(force=True)"
#   run:         gpurun --timeout 900 -- 'bash scripts/ab_run.sh /root/repo/ab_tmp/lib_x.so'
# (ab_tmp/ is git-ignored but travels with the snapshot; gpurun_out/ does not.)
X=${1:?path of the experimental .so}
timeout 120 python scripts/ab_ffn_e2.py $X 2>&1 | tail -4
AVB_LIB=$X timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ffn_tc" --timeout 120 2>&1 | tail -2
AVB_LIB=$X timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "bf16 or stage or golden" --timeout 300 2>&1 | tail -2
for v in exp base exp base; do
  if [ $v == exp ]; then export AVB_LIB=$X; else unset AVB_LIB; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']
print('$v', round(d['value'],1), 'out', k['out'], 'ffn', k['ffn'], 'qkv', k['qkv'], 'attn_n', k['attn_n'], d['clocks']['sm_mhz'])"
done
