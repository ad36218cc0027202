mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 256 --warmup 16 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -3 gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json
python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/bench_ref_r1.json 2>/dev/null; cat gpurun_out/bench_ref_r1.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 64 --warmup 4 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hs_tick_kernel -s 40 -c 2 -o gpurun_out/tick_r1 python bench.py --steps 64 --warmup 4 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
