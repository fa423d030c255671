"""Pixel encoders (reference nets/cnns.py:37-103).  User-pluggable ``nn.Module``s: they stay PyTorch / cuDNN
and feed ``s_rep`` to the CUDA update path (a native conv encoder is SURVEY 8f row N3)."""
import torch
import torch.nn.functional as F
from torch import nn

from . import weight_init


def _conv_out(size, kernel, stride):
    return (size - (kernel - 1) - 1) // stride + 1


class BigPixelEncoder(nn.Module):
    """DrQ encoder: conv3x3(32) s2,1,1,1 -> FC -> LayerNorm -> tanh on obs/255 - 0.5."""

    def __init__(self, obs_shape, out_dim=50):
        super().__init__()
        c, h, w = obs_shape
        self.conv1 = nn.Conv2d(c, 32, kernel_size=3, stride=2)
        self.conv2 = nn.Conv2d(32, 32, kernel_size=3, stride=1)
        self.conv3 = nn.Conv2d(32, 32, kernel_size=3, stride=1)
        self.conv4 = nn.Conv2d(32, 32, kernel_size=3, stride=1)
        h, w = _conv_out(h, 3, 2), _conv_out(w, 3, 2)
        for _ in range(3):
            h, w = _conv_out(h, 3, 1), _conv_out(w, 3, 1)
        self.fc = nn.Linear(h * w * 32, out_dim)
        self.ln = nn.LayerNorm(out_dim)
        self.apply(weight_init)
        self.embedding_dim = out_dim

    def forward(self, obs):
        x = (obs / 255.0) - 0.5
        x = F.relu(self.conv1(x))
        x = F.relu(self.conv2(x))
        x = F.relu(self.conv3(x))
        x = F.relu(self.conv4(x))
        x = self.fc(x.view(x.size(0), -1))
        return torch.tanh(self.ln(x))
