#!/bin/bash
# quick iteration: tc unit tests + bf16 e2e parity + short bench
mkdir -p gpurun_out
( timeout 300 python scripts/tc_debug.py ${1:-attn} > gpurun_out/dbg.log 2>&1; echo "exit=$?" >> gpurun_out/dbg.log ); tail -n 12 gpurun_out/dbg.log
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_tc or gemm_tc or ffn_tc or bf16 or stage" --timeout 300 > gpurun_out/pytest_quick.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_quick.log ); tail -n 8 gpurun_out/pytest_quick.log
( timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "exit=$?" >> gpurun_out/bench_quick.err ); python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
    print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],1),'roofline',round(d['roofline']['frac'],3), d['kernel_ms_per_step'], d['clocks'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_quick.err').read()[-2000:])
PY
