"""Loading of the golden fixtures written by oracle/gen_golden.py (reference's own code)."""
import ast
import glob
import os

import numpy as np
import torch

from oracle import hs_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "hs_*.npz")))


class Golden:
    def __init__(self, path):
        self.name = os.path.basename(path)[3:-4]
        z = np.load(path, allow_pickle=False)
        self.z = {k: z[k] for k in z.files}
        self.P = O.HSParams(**ast.literal_eval(str(self.z["meta/params"])))
        self.E, self.ticks = int(self.z["meta/E"]), int(self.z["meta/ticks"])

    def group(self, prefix):
        n = len(prefix)
        return {k[n:]: torch.from_numpy(v.copy()) for k, v in self.z.items() if k.startswith(prefix)}

    def tp_fn(self):
        if not self.P.use_tp_net:
            return None
        w = self.group("tp_weights/")
        lstm = torch.nn.LSTM(self.P.tp_frame_dim, 64, 1, batch_first=True)
        fc = torch.nn.Linear(64, 3 * self.P.future_step)
        lstm.load_state_dict({k[5:]: v for k, v in w.items() if k.startswith("lstm.")})
        fc.load_state_dict({k[3:]: v for k, v in w.items() if k.startswith("fc.")})

        def fn(x):
            with torch.no_grad():
                out, _ = lstm(x)
                return torch.tanh(fc(out[:, -1, :]))
        return fn


def load_oracle_state(orc, st):
    for k in ("pos", "quat", "linvel", "angvel", "tpos", "tvel", "cyl", "progress"):
        orc.st[k] = st[k].clone()
    orc.throttle, orc.integ, orc.last_rate = st["throttle"].clone(), st["integ"].clone(), st["last_rate"].clone()
    orc.prev_action, orc.stats = st["prev_action"].clone(), st["stats"].clone()
    orc.tp_hist = st["tp_hist"].clone() if "tp_hist" in st else None
