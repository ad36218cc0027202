"""bench.py --config envgen: BASELINE.json configs[3] - HideAndSeek_envgen (adaptive environment generator), 3v1,
8192 envs per GPU, env-sharded over the ranks.

One step = one whole (short) EPISODE of every env through the public API: reset with tasks drawn from the replicated
archive (hs_gen_sample_nearby) + `episode_len` control ticks (CUDA-graph replays) + the episode-end bookkeeping; every
`eval_iter` episodes the ranks all_gather the accepted tasks (ragged rows, parallel.gather_rows) and cap the archive with
farthest point sampling (hs_fps, one cooperative launch) - the collectives and generator kernels are INSIDE the timed
region.  Device time by CUDA events, barrier + synchronize on both sides, max over ranks.
"""
import json
import os
import time


def run(args, out):
    import torch
    import torch.distributed as dist
    import mupe_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"), timeout=datetime.timedelta(seconds=120))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    E = int(os.environ.get("HS_ENVGEN_ENVS", "8192"))
    L = int(os.environ.get("HS_ENVGEN_EPISODE", "64"))          # short episodes: the generator runs often
    cfg = mupe_b200.compose("HideAndSeek_envgen", "mappo", overrides={
        "task.env.num_envs": E, "task.env.max_episode_length": L, "task.sim.device": str(dev), "algo.use_TP_net": 1,
        "task.env.env_offset": rank * E, "task.env.global_num_envs": world * E, "seed": 0})
    base = mupe_b200.IsaacEnv.REGISTRY[cfg.task.name](cfg, headless=True)
    env = mupe_b200.TransformedEnv(base, mupe_b200.Compose(mupe_b200.InitTracker(), mupe_b200.PIDRateController()))
    A = base.num_agents
    act = torch.randn(E, A, 4, device=dev) * 0.3
    act[..., 3] += 0.35

    def episode():
        td = env.reset()
        for _ in range(L):
            td.set(("agents", "action"), act)
            td = mupe_b200.step_mdp(env.step(td))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(args.warmup, base.eval_iter)                          # at least one archive update before timing
    for _ in range(W):
        episode()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hist0 = int(base.gen_buffer._history_buffer.shape[0])
    w0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        episode()
    ev1.record()
    barrier()
    wall = time.perf_counter() - w0
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    hist1 = int(base.gen_buffer._history_buffer.shape[0])
    if rank == 0:
        line = {"metric": "env-steps/sec (3v1 HideAndSeek_envgen)", "value": world * E * L * args.steps / (ms * 1e-3), "unit": "env-steps/s",
                "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"HideAndSeek_envgen 3 pursuers + 1 evader, C=5 (min 4), {E} envs per GPU, use_TP_net=1, "
                                       f"episode length {L}", "envs_per_gpu": E, "ticks_per_step": L,
                           "step": "one episode: reset from the generator + the ticks + episode-end archive bookkeeping",
                           "parallelism": f"env-sharded x{world}"},
                "collective": f"per episode: all_reduce of the success counts; every {base.eval_iter} episodes: ragged all_gather of "
                              "the accepted tasks (parallel.gather_rows) + hs_fps on every rank (replicated archive)",
                "archive_rows_before_after": [hist0, hist1], "wall_ms_per_step": 1e3 * wall / args.steps,
                "api": "public API: TransformedEnv(HideAndSeek_envgen).reset()/step() in a Python loop (graph replay per tick)",
                "gpu_launches": int(base.engine.launches)}
        print(json.dumps(line), file=out, flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()
