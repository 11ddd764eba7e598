#!/bin/bash
# round-2 GPU pass A: fused d-axis kernel unit tests first (debug dump on failure), then the full GPU suite with the
# stated-size parity tests, FFN A/B builds, config-5 timing, bench
mkdir -p gpurun_out
OK=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "dattn_fused" 2>&1 | tail -15 > gpurun_out/r2a_dattn_unit.log || OK=0
grep -q "passed" gpurun_out/r2a_dattn_unit.log && ! grep -q "failed\|error" gpurun_out/r2a_dattn_unit.log || OK=0
cat gpurun_out/r2a_dattn_unit.log
if [ $OK == 0 ]; then
  echo "=== dattn unit test failed: debug dumps"
  for fo in 0 1; do
    AVB_LIB=$PWD/ab_tmp/lib_dattn_debug.so timeout 120 python scripts/debug_dattn.py 300 100 $fo > gpurun_out/r2a_dattn_debug_$fo.log 2>&1
    tail -25 gpurun_out/r2a_dattn_debug_$fo.log
  done
  export AVB_LIB=$PWD/ab_tmp/lib_ab.so AVB_DATTN_MODE=0
  echo "=== continuing with the unfused d-axis path (AVB_DATTN_MODE=0)"
fi
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -s --deselect tests/test_gpu_parity.py::test_dattn_fused_sublayer_against_torch 2>&1 | tail -80 > gpurun_out/r2a_pytest_gpu.log
tail -12 gpurun_out/r2a_pytest_gpu.log
timeout 300 python scripts/ab_ffn_e2.py ab_tmp/lib_ffn_E1X64.so ab_tmp/lib_ffn_DECOUPLE.so ab_tmp/lib_ffn_BOTH.so > gpurun_out/r2a_ffn_ab.log 2>&1
cat gpurun_out/r2a_ffn_ab.log
timeout 300 python scripts/time_config5.py > gpurun_out/r2a_config5.json 2> gpurun_out/r2a_config5.err
cat gpurun_out/r2a_config5.json
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
cat gpurun_out/r2a_bench.json
if [ $OK == 1 ]; then  # A/B of the d-axis modes in the bench (AB build honours AVB_DATTN_MODE)
  for m in 0 2 1; do
    AVB_LIB=$PWD/ab_tmp/lib_ab.so AVB_DATTN_MODE=$m timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']
print('dattn_mode $m', round(d['value'],1), k, d['clocks']['sm_mhz'])" | tee -a gpurun_out/r2a_dattn_modes.log
  done
fi
