"""TEST INFRASTRUCTURE ONLY - generates tests/golden/envgen_sanity.npz by running the reference's
OWN `sanity_check`, `continuous_to_grid` and `set_outside_circle_to_one`
(omni_drones/envs/hide_and_seek/hideandseek_envgen.py:140-207, AST-extracted from /root/reference,
build container only) on candidate tasks produced by oracle/envgen_oracle.py, so that the
acceptance rule of the device sampler is pinned to the reference's verdicts.

Run:  python -m oracle.gen_envgen_golden
"""
import ast
import os

import numpy as np
import torch

from oracle import envgen_oracle as G
from oracle import reset_sampler as RS
from oracle.ref_harness import REF

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "envgen_sanity.npz")


def extract(names):
    src = (REF / "omni_drones/envs/hide_and_seek/hideandseek_envgen.py").read_text()
    ns = dict(torch=torch, np=np)
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), f"<ref:envgen:{node.name}>", "exec"), ns)
    assert all(n in ns for n in names)
    return ns


def main():
    ns = extract(["sanity_check", "continuous_to_grid", "set_outside_circle_to_one"])
    A, C, arena, cs, mh = 3, 5, 0.9, 0.1, 1.2
    ng = int(arena * 2 / (2 * cs))
    grid_map = ns["set_outside_circle_to_one"](np.zeros((1, ng, ng), dtype=int))[0]
    center_pos, center_grid = torch.zeros(1, 2), torch.ones(1, 2, dtype=torch.int) * int(ng / 2)
    d = RS.ResetDist.for_task(num_cylinders=C, max_height=mh, cylinder_height=mh, seed=41)
    o = RS.sample_reset(d, 600, 1)
    hist = np.concatenate([o["drone_pos"].reshape(600, -1), o["target_pos"], o["cyl_pos"].reshape(600, -1)], -1)
    rng = np.random.default_rng(4)
    cands = []
    for expand_cyl, step in ((True, 0.1), (False, 0.3), (True, 0.5)):
        r = G.sample_nearby(hist, 700, A, C, arena, cs, mh, expand_cyl, step, seed=9, epoch=2)
        cands.append(r["tasks"])
        # plus raw (unfiltered) perturbations so that rejected candidates are well represented
        raw = hist[rng.integers(0, 600, 700)].copy()
        raw[:, :12] += rng.uniform(-1, 1, (700, 12)).astype(np.float32) * np.float32(step)
        raw[:, 12:].reshape(700, C, 3)[:, :, :2] += rng.choice([-1, 0, 1], (700, C, 2)).astype(np.float32) * np.float32(2 * cs)
        b = G.task_bounds(A, C, arena, 2 * cs, mh)
        cands.append(np.clip(raw, b[:, 0], b[:, 1]).astype(np.float32))
    cands = np.concatenate(cands)
    verdict = np.zeros(len(cands), np.uint8)
    cells = np.zeros((len(cands), A + 1 + C, 2), np.int64)
    for i, t in enumerate(cands):
        dp, tp, cp = t[:3 * A].reshape(-1, 3), t[3 * A:3 * A + 3].reshape(-1, 3), t[3 * A + 3:].reshape(-1, 3)
        g = [ns["continuous_to_grid"](torch.from_numpy(x[:, :2].copy()), ng, 2 * cs, center_pos, center_grid).numpy()
             for x in (dp, tp, cp)]
        cells[i] = np.concatenate(g)
        verdict[i] = int(ns["sanity_check"](grid_map, g[0], g[1], g[2]))
    np.savez_compressed(OUT, tasks=cands, ref_verdict=verdict, ref_cells=cells, ref_grid_map=grid_map.astype(np.int8),
                        A=A, C=C, arena=arena, cylinder_size=cs, max_height=mh)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; accepted fraction", verdict.mean())


if __name__ == "__main__":
    main()
