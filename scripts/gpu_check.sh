#!/bin/bash
# First-contact GPU check: each group in its own process (a trapped kernel poisons the CUDA context),
# each under its own timeout so a hung kernel cannot hold the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit=$?" >> gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
run dbg_gemm python scripts/tc_debug.py gemm
run dbg_attn python scripts/tc_debug.py attn
run t_fp32 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fp32 or stage or sample" --timeout 300 -x
run t_gemm python -m pytest tests/test_gpu_parity.py -q -m gpu -k "gemm_tc" --timeout 120
run t_attn python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_tc" --timeout 120
run t_bf16 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "bf16 and not attention_tc and not gemm_tc" --timeout 300
