# bench.py over a few per-GPU batch sizes (cfg2 unless WFL_BENCH_WORKLOAD is set): ms/step, audio-s/s, e2e, conv31 TF, GEMM-family TF
for b in ${BATCHES:-8 16 32}; do
  python bench.py --no-cpu-baseline --batch $b 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('batch', $b, 'ms/step', d['ms_per_step'], 'audio-s/s', d['value'], 'e2e', d['e2e']['value'], 'conv31 TF', d['roofline']['achieved'], 'gemm family TF', d['roofline']['gemm_family']['achieved'], 'share', d['roofline']['gemm_family']['share_of_step'], 'p50 ms', d['latency_p50_ms']['value'])"
done
