"""omni_drones/controllers of the reference.  `PIDRateController(dt, g, params)` is constructed by scripts/train.py:166-169
and handed to the PIDrate transform; the body-rate PID itself (lee_position_controller.py:476-550) runs inside the tick
kernel, so this object only carries the parameters."""
import torch


class PIDRateController(torch.nn.Module):
    def __init__(self, dt, g, uav_params):
        super().__init__()
        self.dt, self.g, self.uav_params = dt, g, uav_params

    def forward(self, *a, **k):
        raise RuntimeError("the rate PID is fused into libhs_b200.so (hs_step_pre); it is not evaluated on the host")


def _out_of_scope(name):
    class _C:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name}: not part of the HideAndSeek hot path (tasks use action_transform: PIDrate)")
    _C.__name__ = name
    return _C


LeePositionController = _out_of_scope("LeePositionController")
AttitudeController = _out_of_scope("AttitudeController")
RateController = _out_of_scope("RateController")
