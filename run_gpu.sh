mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 512 --warmup 16 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -3 gpurun_out/bench_r1c.err; cat gpurun_out/bench_r1c.json
