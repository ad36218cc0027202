mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -6
HS_TP_VARIANT=1 python tools/sweep.py 4096 16384 1048576 > gpurun_out/sweep_r1g_mma.jsonl 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err; cut -c1-420 gpurun_out/sweep_r1g_mma.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 512 --warmup 16 > gpurun_out/bench_n2_r1.json 2> gpurun_out/bench_n2_r1.err; grep -E "Error|error" gpurun_out/bench_n2_r1.err | head -3; cut -c1-700 gpurun_out/bench_n2_r1.json
