mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fused_predictor" 2>&1 | tail -4
for v in 3 4; do HS_TP_VARIANT=$v timeout 300 python tools/sweep.py 4096 8192 16384 65536 1048576 > gpurun_out/sweep_r1s_v$v.jsonl 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err; python -c "
import json
for l in open('gpurun_out/sweep_r1s_v$v.jsonl'):
    r=json.loads(l); print('v$v E',r['E'],'tick us',round(r['tick']['us_per_launch'],1),'tp us',round(r['tp_fill']['us_per_launch'],1),'both Menv/s',round(r['tick_plus_tp']['env_steps_per_s']/1e6,1))
"; done
