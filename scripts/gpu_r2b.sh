#!/bin/bash
# round-2 GPU pass B: where does dattn_tc_kernel's time go (trace build), what do the exponentials cost (NOEXP builds)
mkdir -p gpurun_out
L=$PWD/ab_tmp
{
echo "=== dattn trace (fuse_out=1)"; AVB_LIB=$L/lib_trace.so timeout 120 python scripts/trace_dattn.py 16000 100 1 40
echo "=== dattn trace (fuse_out=0)"; AVB_LIB=$L/lib_trace.so timeout 120 python scripts/trace_dattn.py 16000 100 0 40 | head -60
for v in ab nobw noexp; do
  echo "=== dattn timing lib_$v"; AVB_LIB=$L/lib_$v.so timeout 120 python scripts/trace_dattn.py 64000 100 1 | head -3
done
for v in ab noexp; do
  echo "=== attention timing lib_$v (prepacked, fast kernel only)"
  AVB_LIB=$L/lib_$v.so AVB_ATTN_PREPACKED=1 AVB_ATTN_MODE=fast timeout 120 python scripts/prof_attn.py 64
done
} > gpurun_out/r2b_trace.log 2>&1
cat gpurun_out/r2b_trace.log
