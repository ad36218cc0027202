mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "host_buffer" 2>&1 | grep -E "Error|error|passed|failed|FAILED|off \(" | head
timeout 600 python bench.py > gpurun_out/bench_r1w.json 2> gpurun_out/bench_r1w.err; tail -3 gpurun_out/bench_r1w.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1w.json').read().strip().splitlines()[-1])
print('value',d['value']); print('e2e',d['e2e']); print('e2e_python_env',d.get('e2e_python_env',{}).get('value'))
PY
