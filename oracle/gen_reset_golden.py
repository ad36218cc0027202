"""TEST INFRASTRUCTURE ONLY - generates tests/golden/reset_grid.npz by running the reference's
OWN source (AST-extracted by oracle/ref_harness.py from /root/reference, build container only):

* `rejection_sampling_random_cylinder` (hideandseek.py:576-607) is executed unmodified on the
  pursuer/evader positions drawn by oracle/reset_sampler.py; the occupancy `grid_map` it hands to
  `select_unoccupied_positions` is captured, together with the cylinder xy it returns (torch's
  randperm stream - used only for a distribution check) and the active-cylinder counts it draws.
* `grid_to_continuous` (hideandseek.py:120-142) is evaluated on every cell -> the cell->metres table.
* `euler_to_quaternion` (omni_drones/utils/torch.py) on the oracle's rpy draws.

Run:  python -m oracle.gen_reset_golden
"""
import os
import types

import numpy as np
import torch

from oracle import reset_sampler as RS
from oracle.ref_harness import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reset_grid.npz")


def run_case(R, name, E, **kw):
    d = RS.ResetDist.for_task(**kw)
    o = RS.sample_reset(d, E, epoch=1)
    ns = R["ns"]
    captured = {}
    orig = ns["select_unoccupied_positions"]

    def spy(occupancy_matrix, num_objects):
        captured["grid_map"] = occupancy_matrix.clone()
        return orig(occupancy_matrix, num_objects)
    ns["select_unoccupied_positions"] = spy
    try:
        fake = types.SimpleNamespace(
            cylinder_size=kw.get("cylinder_size", 0.1), arena_size=kw.get("arena_size", 0.9), device="cpu",
            num_agents=d.num_agents, use_fixed_num=d.fixed_num >= 0, fixed_num=d.fixed_num,
            min_cylinders=d.min_cylinders, num_cylinders=d.num_cylinders, boundary=d.boundary)
        torch.manual_seed(7)
        dpos = torch.from_numpy(o["drone_pos"])
        tpos = torch.from_numpy(o["target_pos"]).unsqueeze(1)
        ref_xy = R["env"]["rejection_sampling_random_cylinder"](fake, torch.arange(E), dpos, tpos)
    finally:
        ns["select_unoccupied_positions"] = orig
    ng = d.num_grid
    cells = torch.stack(torch.meshgrid(torch.arange(ng), torch.arange(ng), indexing="ij"), -1).reshape(1, ng * ng, 2)
    table = ns["grid_to_continuous"](cells, d.boundary, d.grid_size, torch.zeros(1, 1, 2),
                                     torch.ones(1, 1, 2, dtype=torch.int) * int(ng / 2))
    rot = R["ut"].euler_to_quaternion(torch.from_numpy(o["rpy"]))
    return {
        f"{name}/E": np.int64(E), f"{name}/kw": np.array(repr(sorted(kw.items()))),
        f"{name}/drone_pos": o["drone_pos"], f"{name}/target_pos": o["target_pos"], f"{name}/rpy": o["rpy"],
        f"{name}/ref_grid_map": captured["grid_map"].numpy().astype(np.int8),
        f"{name}/ref_cyl_xy": ref_xy.numpy().astype(np.float32),
        f"{name}/ref_active": fake.active_cylinders.numpy().astype(np.float32),
        f"{name}/ref_cell_table": table.reshape(ng * ng, 2).numpy().astype(np.float32),
        f"{name}/ref_rot": rot.numpy().astype(np.float32),
    }


def main():
    R = load_reference()
    data = {}
    data.update(run_case(R, "c5", 3000, num_cylinders=5, min_cylinders=0, seed=11))
    data.update(run_case(R, "c8", 3000, num_cylinders=8, min_cylinders=2, seed=12))
    data.update(run_case(R, "c5_big", 1000, num_cylinders=5, min_cylinders=0, seed=13, arena_size=1.1))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **data)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
