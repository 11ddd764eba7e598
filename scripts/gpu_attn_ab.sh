#!/bin/bash
# A/B of attention builds: unit tests with each library, then interleaved timings of the n-axis launch (B = 64)
mkdir -p gpurun_out
{
for lib in "$@"; do
  echo "=== $lib: tests"; AVB_LIB=$PWD/ab_tmp/$lib timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "attention_tc" 2>&1 | grep -E "passed|failed|FAILED"
done
for rep in 1 2; do for lib in "$@"; do
  echo "=== $lib: timing"; AVB_LIB=$PWD/ab_tmp/$lib AVB_ATTN_PREPACKED=1 AVB_ATTN_MODE=fast timeout 120 python scripts/prof_attn.py 64 | tail -2
done; done
} > gpurun_out/attn_ab.log 2>&1
cat gpurun_out/attn_ab.log
