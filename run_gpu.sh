mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python tools/sweep.py 4096 16384 65536 262144 1048576 > gpurun_out/sweep_r1e.jsonl 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err; cut -c1-420 gpurun_out/sweep_r1e.jsonl
python bench.py --steps 512 --warmup 16 > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -3 gpurun_out/bench_r1f.err; cat gpurun_out/bench_r1f.json
