mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tcgen05" -x 2>&1 | tail -12
HS_TP_VARIANT=2 timeout 200 python tools/sweep.py 4096 1048576 > gpurun_out/sweep_r1h_tc.jsonl 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err; cut -c1-420 gpurun_out/sweep_r1h_tc.jsonl
