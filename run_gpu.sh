mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fused_predictor" 2>&1 | tail -4
HS_TP_VARIANT=3 timeout 300 python tools/sweep.py 4096 8192 16384 > gpurun_out/sweep_r1p_v3.jsonl 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err; python -c "
import json
for l in open('gpurun_out/sweep_r1p_v3.jsonl'):
    r=json.loads(l); print('v3 E',r['E'],'tick us',round(r['tick']['us_per_launch'],1),'tp us',round(r['tp_fill']['us_per_launch'],1),'both Menv/s',round(r['tick_plus_tp']['env_steps_per_s']/1e6,1))
"
HS_TP_VARIANT=3 timeout 600 ncu --set full --import-source on --clock-control none -k regex:hs_tp_fill_tcn -s 20 -c 1 -o gpurun_out/tcn4096_r1p -f python tools/sweep.py 4096 > gpurun_out/ncu_tcn.log 2>&1; tail -2 gpurun_out/ncu_tcn.log
