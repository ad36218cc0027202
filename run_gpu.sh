mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
timeout 300 python tools/reset_bench.py 4096 65536 > gpurun_out/reset_bench_r1.jsonl 2> gpurun_out/reset_bench.err; tail -3 gpurun_out/reset_bench.err; cat gpurun_out/reset_bench_r1.jsonl
timeout 300 python tools/e2e_profile.py > gpurun_out/e2e_profile.txt 2>&1; grep -A30 "tottime" gpurun_out/e2e_profile.txt | cut -c1-150 | head -34
HS_BENCH_TIMED_ONLY=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1m.csv python bench.py --steps 64 --warmup 4 > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300; wc -l gpurun_out/launches_r1m.csv
