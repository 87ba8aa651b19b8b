#!/bin/bash
# new-feature tests first (fast fail), then the whole GPU suite and a bench line
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "mel_power or gather_cols or not_a_multiple" > gpurun_out/new_ops.log 2>&1
a=$?; tail -5 gpurun_out/new_ops.log
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -x -s -k "mel_none" > gpurun_out/new_model.log 2>&1
b=$?; grep "^\[" gpurun_out/new_model.log | tail -12; tail -5 gpurun_out/new_model.log
if [ $a -ne 0 ] || [ $b -ne 0 ]; then echo "NEW TESTS FAILED ($a, $b)"; exit 1; fi
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/full.log 2>&1
c=$?; tail -4 gpurun_out/full.log
timeout 240 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 1500 gpurun_out/bench_final.json
exit $c
