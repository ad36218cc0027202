mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python tools/sweep.py > gpurun_out/sweep_r1b.jsonl 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err; cut -c1-600 gpurun_out/sweep_r1b.jsonl
python bench.py --steps 512 --warmup 16 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; tail -3 gpurun_out/bench_r1d.err; cat gpurun_out/bench_r1d.json
