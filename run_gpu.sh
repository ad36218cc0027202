mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -6
python tools/sweep.py 4096 16384 1048576 > gpurun_out/sweep_r1f.jsonl 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err; cut -c1-420 gpurun_out/sweep_r1f.jsonl
python bench.py > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; tail -3 gpurun_out/bench_r1h.err; cut -c1-1700 gpurun_out/bench_r1h.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 64 --warmup 4 > gpurun_out/ncu_bench.log 2>&1; grep -c hs_ gpurun_out/launches_r1h.csv
