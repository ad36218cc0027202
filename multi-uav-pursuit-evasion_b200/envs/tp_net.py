"""Trajectory predictor that the environment calls between the two halves of a tick.

Same architecture and parameter names (``lstm``, ``fc``) as the reference's TP_net
(omni_drones/learning/mappo.py:572-589) so that checkpoints interchange: one-layer LSTM,
hidden 64, last hidden state -> Linear -> tanh.  It is a plain torch module (the policy
side trains it); the env only runs its forward."""
import torch
import torch.nn as nn


class TP_net(nn.Module):
    def __init__(self, input_dim: int, output_dim: int, future_predcition_step: int, window_step: int = 1):
        super().__init__()
        self.hidden_dim, self.num_layers = 64, 1
        self.future_predcition_step, self.window_step = future_predcition_step, window_step
        self.lstm = nn.LSTM(input_dim, self.hidden_dim, self.num_layers, batch_first=True)
        self.fc = nn.Linear(self.hidden_dim, output_dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out, _ = self.lstm(x)                    # zero initial (h, c), like the reference
        return torch.tanh(self.fc(out[:, -1, :]))
