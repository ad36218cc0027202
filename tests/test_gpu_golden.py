"""GPU: the CUDA kernels against fixtures produced by the REFERENCE'S OWN SOURCE
(tests/golden/*.npz, written by oracle/gen_golden.py in the build container).  Each recorded
tick is replayed from its stored pre-tick state; bar = 1e-4 relative fp32 (north_star)."""
import pytest
import torch

from tests.golden_util import Golden, golden_files
from tests.util import assert_close, hs_config_from_params

pytestmark = pytest.mark.gpu
FLIP = 5e-3          # see tests/test_gpu_parity.py
# 'wall' is mirror-symmetric about y=0 with the evader and one pursuer ON the axis: the y
# component of the evader's potential-field force is an exact cancellation (true value 0), so
# its sign-normalised velocity v*f/(|f|+1e-5) is pure rounding noise in ANY implementation
# (the reference's own value there is -8e-3 m/s from a force of -6e-8).  Everything derived
# from the evader's y gets a per-tensor budget for those elements in that fixture.
SYMMETRIC_FLIP = {"wall_tp": 0.02}


def load_engine_state(eng, st, v_prey=1.3):
    from mupe_b200 import _lib as L
    for f, k in ((L.FIELD_DRONE_POS, "pos"), (L.FIELD_DRONE_ROT, "quat"), (L.FIELD_DRONE_LINVEL, "linvel"),
                 (L.FIELD_DRONE_ANGVEL, "angvel"), (L.FIELD_THROTTLE, "throttle"), (L.FIELD_PID_INTEG, "integ"),
                 (L.FIELD_PID_LAST_RATE, "last_rate"), (L.FIELD_TARGET_POS, "tpos"), (L.FIELD_TARGET_VEL, "tvel"),
                 (L.FIELD_PROGRESS, "progress")):
        eng.set_state(f, st[k])
    if eng.C > 0:
        eng.set_state(L.FIELD_CYL_POS, st["cyl"])
    eng.prev_action.copy_(st["prev_action"])
    eng.stats.copy_(st["stats"].t().contiguous())
    eng.v_prey.fill_(v_prey)
    if "tp_hist" in st:
        eng.out["tp_input"].copy_(st["tp_hist"])


NAMES = {"cmds": "rotor_cmds", "others": "state_others", "cylinders": "obs_cylinders"}


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("hs_")[-1][:-4])
def test_kernels_replay_reference_ticks(path):
    import mupe_b200
    G = Golden(path)
    P, E = G.P, G.E
    dev = torch.device("cuda:0")
    eng = mupe_b200.HsEngine(hs_config_from_params(P, E), dev)
    init = G.group("init/")
    got = eng.reset(None, init["drone_pos"], init["drone_rot"], init["target_pos"], init["cyl_pos"])
    want = G.group("reset/out/")
    if P.use_tp_net:
        eng.step_post(want["tp_pred"].to(dev))
    for k, v in want.items():
        if k != "tp_pred":
            assert_close(f"{G.name}/reset/{k}", got[NAMES.get(k, k)], v)
    for t in range(G.ticks):
        load_engine_state(eng, G.group(f"t{t}/pre/"))
        io = G.group(f"t{t}/")
        got = eng.step_pre(io["action"].to(dev), raw=True, reset_pid=io["done_prev"].bool().to(dev))
        want = G.group(f"t{t}/out/")
        if P.use_tp_net:
            eng.step_post(want["tp_pred"].to(dev))
        for k, v in want.items():
            if k == "tp_pred":
                continue
            if k == "stats":
                g = eng.stats.t()
            elif k == "prev_action":
                g = eng.prev_action
            else:
                g = got[NAMES.get(k, k)].float()
            flip = SYMMETRIC_FLIP.get(G.name, FLIP)
            # ctbr carries the raw PID output whose D term amplifies 1-ulp body-rate differences by
            # 1/dt * kd * 180/pi ~ 1.4e4 -> compare it relative to the tensor's scale
            atol = 1e-4 if k == "stats" else (1e-4 * float(v.abs().max()) if k == "ctbr" else 1e-5)
            if G.name in SYMMETRIC_FLIP and k in ("state_self", "state_drones", "tp_input", "tp_groundtruth"):
                atol = 5e-3         # these carry the evader's y position / velocity (noise, see above)
            assert_close(f"{G.name}/t{t}/{k}", g, v, rtol=1e-4, atol=atol, max_bad_frac=flip)
        post = G.group(f"t{t}/post/")
        from mupe_b200 import _lib as L
        for f, k in ((L.FIELD_DRONE_POS, "pos"), (L.FIELD_DRONE_ROT, "quat"), (L.FIELD_DRONE_LINVEL, "linvel"),
                     (L.FIELD_DRONE_ANGVEL, "angvel"), (L.FIELD_THROTTLE, "throttle"), (L.FIELD_PID_INTEG, "integ"),
                     (L.FIELD_TARGET_POS, "tpos"), (L.FIELD_TARGET_VEL, "tvel"), (L.FIELD_PROGRESS, "progress")):
            atol = 5e-3 if (G.name in SYMMETRIC_FLIP and k in ("tpos", "tvel")) else 1e-5
            assert_close(f"{G.name}/t{t}/post/{k}", eng.get_state(f), post[k], atol=atol,
                         max_bad_frac=SYMMETRIC_FLIP.get(G.name, FLIP))
    eng.close()
