#!/bin/bash
# Full GPU pass: parity suite, smoke, bench, then the ncu launch list of the bench command.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu.log ); tail -n 15 gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit=$?" >> gpurun_out/smoke.log ); tail -n 5 gpurun_out/smoke.log
( timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit=$?" >> gpurun_out/bench.err ); cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
  timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 180 -c 170 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
  echo "ncu exit=$?"; tail -n 3 gpurun_out/ncu.log
fi
