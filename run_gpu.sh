mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_r1r.json 2> gpurun_out/bench_r1r.err; tail -3 gpurun_out/bench_r1r.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1r.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value']); print('at_scale',d.get('roofline_at_scale'))
PY
HS_SWEEP_C=8 timeout 300 python tools/sweep.py 16384 > gpurun_out/sweep_r1r_c8.jsonl 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err; cat gpurun_out/sweep_r1r_c8.jsonl | cut -c1-600
