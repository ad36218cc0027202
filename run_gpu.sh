mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_envgen_device.py -q -m gpu 2>&1 | grep -E "Error|error|passed|failed|FAILED|off \(|assert|differs" | head -20
