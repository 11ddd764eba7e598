#!/bin/bash
# dattn_tc_kernel iteration: unit tests (correctness) + timing (production lib and any ab_tmp/ libs named in $2..) + role trace
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "dattn" 2>&1 | grep -E "passed|failed|FAILED"
echo "=== timing (production lib)"; timeout 120 python scripts/trace_dattn.py 64000 100 1 2>/dev/null | head -3
for lib in "${@:2}"; do
  echo "=== $lib"; AVB_LIB=$PWD/ab_tmp/$lib timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "dattn" 2>&1 | grep -E "passed|failed|FAILED"
  AVB_LIB=$PWD/ab_tmp/$lib timeout 120 python scripts/trace_dattn.py 64000 100 1 2>/dev/null | head -3
done
echo "=== trace"; AVB_LIB=$PWD/ab_tmp/lib_trace.so timeout 120 python scripts/trace_dattn.py 16000 100 1 ${1:-40}
} > gpurun_out/dattn_iter.log 2>&1
grep -v '^ step' gpurun_out/dattn_iter.log
