mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 512 --warmup 16 > gpurun_out/bench_n2_r1v.json 2> gpurun_out/bench_n2_r1v.err; tail -3 gpurun_out/bench_n2_r1v.err | cut -c1-300; cut -c1-700 gpurun_out/bench_n2_r1v.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 6 --warmup 1 2>&1 | tail -2 | cut -c1-400
