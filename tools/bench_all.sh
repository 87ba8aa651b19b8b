# one bench line per BASELINE config on one GPU (cfg2 is the headline; the others are parity-test cases timed for the record)
for w in ${WORKLOADS:-cfg1 cfg2 cfg3 cfg4 cfg5}; do
  WFL_BENCH_WORKLOAD=$w python bench.py --no-cpu-baseline --steps ${STEPS:-5} --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('$w', 'ms/step', d['ms_per_step'], 'audio-s/s', d['value'], 'e2e', d['e2e']['value'], '| dominant', r['kernel'], r['achieved'], 'TF | gemm family', r['gemm_family']['achieved'], 'TF share', r['gemm_family']['share_of_step'], '| p50 ms', d['latency_p50_ms']['value'])"
done
