"""CPU, gloo, world_size 2: the host-side logic of the env-sharded multi-GPU path
(multi-uav-pursuit-evasion_b200/parallel.py).  The data path itself has no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mupe_b200.parallel as par
        G = 11                                   # uneven split: 6 + 5
        sh = par.current_shard(G)
        lo, hi = sh.bounds
        local = torch.arange(lo, hi, dtype=torch.float32) * 10
        full = par.gather_env_vector(local, sh)
        assert torch.equal(full, torch.arange(G, dtype=torch.float32) * 10)
        # success mean over ALL envs: rank 0 all ones (6 envs), rank 1 all zeros (5 envs) -> 6/11
        succ = torch.ones(hi - lo) if rank == 0 else torch.zeros(hi - lo)
        m = par.global_mean(succ)
        assert abs(float(m) - 6 / 11) < 1e-6
        # curriculum: only rank 1 sees a done env; global success 1.0 -> everybody speeds up once
        v = torch.tensor([1.0])
        done = torch.zeros(hi - lo, dtype=torch.bool)
        if rank == 1:
            done[0] = True
        par.curriculum_step(v, done, torch.ones(hi - lo))
        assert abs(float(v) - 1.05) < 1e-6
        par.curriculum_step(v, torch.zeros(hi - lo, dtype=torch.bool), torch.ones(hi - lo))
        assert abs(float(v) - 1.05) < 1e-6        # nobody done -> unchanged
        par.curriculum_step(v, done, succ)          # success 6/11 < 0.98 -> unchanged
        assert abs(float(v) - 1.05) < 1e-6
        for _ in range(10):
            par.curriculum_step(v, done, torch.ones(hi - lo))
        assert abs(float(v) - 1.3) < 1e-6         # capped
        assert sh.seed(7) != par.Shard(1 - rank, 2, G).seed(7)
        # replicated envgen archive: ragged [n_r, d] blocks -> the same [sum n, d] on every rank, rank order
        rows = torch.full((3 if rank == 0 else 0, 4), float(rank))
        allr = par.gather_rows(rows)
        assert allr.shape == (3, 4) and torch.equal(allr, torch.zeros(3, 4))
        rows = torch.arange((2 + 3 * rank) * 5, dtype=torch.float32).reshape(-1, 5) + 100 * rank
        allr = par.gather_rows(rows)
        want = torch.cat([torch.arange(10, dtype=torch.float32).reshape(2, 5),
                          torch.arange(25, dtype=torch.float32).reshape(5, 5) + 100])
        assert torch.equal(allr, want)
        assert par.gather_rows(torch.zeros(0, 4)).shape == (0, 4)
        # envgen: success over the GLOBAL uniform prefix / archive suffix from per-rank [sum, count] pairs
        # (hideandseek_envgen.py:1241-1246): 8 global envs, prefix of 5 -> rank 0 holds 4 uniform envs, rank 1 one
        s_local = torch.tensor([1., 0., 1., 1.]) if rank == 0 else torch.tensor([0., 1., 1., 1.])
        nu = 4 if rank == 0 else 1
        acc = par.global_sum(torch.stack([s_local[:nu].sum(), torch.tensor(float(nu)), s_local[nu:].sum(), torch.tensor(float(4 - nu))]))
        assert abs(float(acc[0] / acc[1]) - 3 / 5) < 1e-6 and abs(float(acc[2] / acc[3]) - 3 / 3) < 1e-6
        q.put((rank, "ok"))
    except Exception as e:                        # surface the failure in the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_the_index_space():
    import mupe_b200.parallel as par
    for G in (1, 7, 8, 65536, 65537):
        for W in (1, 2, 3, 8):
            b = [par.Shard(r, W, G).bounds for r in range(W)]
            assert b[0][0] == 0 and b[-1][1] == G
            assert all(b[i][1] == b[i + 1][0] for i in range(W - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_gloo_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
