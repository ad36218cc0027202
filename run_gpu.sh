mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -6
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; tail -3 gpurun_out/bench_r1g.err; cat gpurun_out/bench_r1g.json
python bench.py --impl reference > gpurun_out/bench_ref_r1g.json 2>/dev/null; cat gpurun_out/bench_ref_r1g.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 64 --warmup 4 > gpurun_out/ncu_bench.log 2>&1
