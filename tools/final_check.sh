#!/bin/bash
# end-of-round check on one B200: GPU test suite, smoke(), bench line, then the ncu launch list of the same bench command
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/full.log 2>&1; a=$?; tail -3 gpurun_out/full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1; b=$?; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; c=$?; tail -c 900 gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 700 --csv --log-file gpurun_out/r01_launches_bench_cfg2_v6.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "rc tests=$a smoke=$b bench=$c"
