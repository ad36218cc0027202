mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_envgen_device.py tests/test_gpu_env.py -q -m gpu -x 2>&1 | tail -12
timeout 300 python - <<'PY' 2>&1 | tail -5
import time, torch, numpy as np, sys
sys.path.insert(0, '.')
import mupe_b200
from mupe_b200.envs.hideandseek_envgen import GenBufferDevice, GenBuffer, farthest_point_sampling
gb = GenBufferDevice(3, 5, device="cuda:0")
pts = torch.rand(70000, 27, device="cuda")
for k in (500, 5000):
    gb.fps(pts, 16); torch.cuda.synchronize(); t = time.perf_counter(); idx = gb.fps(pts, k); torch.cuda.synchronize()
    print("hs_fps n=70000 dim=27 k=%d: %.2f ms (%.2f us/point)" % (k, 1e3 * (time.perf_counter() - t), 1e6 * (time.perf_counter() - t) / k))
t = time.perf_counter(); farthest_point_sampling(pts.cpu(), 200); dt = time.perf_counter() - t
print("torch-CPU FPS (the previous path) k=200: %.1f ms -> k=5000 would take ~%.1f s" % (1e3 * dt, dt * 25))
gb._history_buffer = torch.rand(5000, 27, device="cuda") * 0.4
gb.samplenearby(45000, True, 0.1); torch.cuda.synchronize(); t = time.perf_counter(); gb.samplenearby(45000, True, 0.1); torch.cuda.synchronize()
print("hs_gen_sample_nearby 45875 tasks (65536 envs x 0.7): %.3f ms" % (1e3 * (time.perf_counter() - t)))
hb = GenBuffer(3, 5); hb._history_buffer = gb._history_buffer.cpu().numpy()
t = time.perf_counter(); hb.samplenearby(500, True, 0.1); dt = time.perf_counter() - t
print("host loop (reference style) 500 tasks: %.1f ms -> 45875 tasks ~%.1f s" % (1e3 * dt, dt * 45875 / 500))
PY
