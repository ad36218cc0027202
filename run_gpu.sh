mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -E "Error|error|passed|failed|FAILED|off \(" | head -20
timeout 300 python tools/sweep.py 4096 8192 16384 65536 1048576 > gpurun_out/sweep_r1v_auto.jsonl 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err; python -c "
import json
for l in open('gpurun_out/sweep_r1v_auto.jsonl'):
    r=json.loads(l); print('auto E',r['E'],'tick us',round(r['tick']['us_per_launch'],1),'tp us',round(r['tp_fill']['us_per_launch'],1),'both Menv/s',round(r['tick_plus_tp']['env_steps_per_s']/1e6,1))
"
HS_SWEEP_C=8 timeout 300 python tools/sweep.py 16384 > gpurun_out/sweep_r1v_c8.jsonl 2> gpurun_out/sweep.err; cut -c1-400 gpurun_out/sweep_r1v_c8.jsonl
HS_TP_VARIANT=4 timeout 600 ncu --set full --import-source on --clock-control none -k regex:hs_tp_fill_tcw -s 4 -c 1 -o gpurun_out/tcw65536_r1v -f python tools/sweep.py 65536 > gpurun_out/ncu_tcw.log 2>&1; tail -2 gpurun_out/ncu_tcw.log
