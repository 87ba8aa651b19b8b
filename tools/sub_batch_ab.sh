#!/bin/bash
# A/B of the sub-batched forward (engine._sub_batch) on the bench workload: one pass vs automatic vs fixed sizes
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "sub_batched or full_size_cfg2" 2>&1 | tail -3
for sb in 0 auto 8 16; do
  if [ "$sb" = auto ]; then unset WFL_SUB_BATCH; else export WFL_SUB_BATCH=$sb; fi
  for b in ${BATCHES:-32}; do
    python bench.py --no-cpu-baseline --batch $b 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('sub_batch', '$sb', 'batch', $b, 'ms/step', d['ms_per_step'], 'audio-s/s', d['value'], 'e2e', d['e2e']['value'], 'conv31 TF', d['roofline']['achieved'], 'launches', d['gpu_launches'], 'clk', d['clocks']['sm_mhz'])"
  done
done 2>&1 | tee gpurun_out/sub_batch_ab.log
