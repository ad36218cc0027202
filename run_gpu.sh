mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6
timeout 300 python tools/sweep.py 4096 16384 65536 262144 1048576 > gpurun_out/sweep_r1k_auto.jsonl 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err; python -c "
import json
for l in open('gpurun_out/sweep_r1k_auto.jsonl'):
    r=json.loads(l); print('auto E',r['E'],'tick us',round(r['tick']['us_per_launch'],1),'frac',round(r['tick']['frac_of_measured_hbm'],3),'tp us',round(r['tp_fill']['us_per_launch'],1),'both Menv/s',round(r['tick_plus_tp']['env_steps_per_s']/1e6,1))
"
timeout 600 python bench.py > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err; tail -3 gpurun_out/bench_r1i.err; cut -c1-400 gpurun_out/bench_r1i.json
