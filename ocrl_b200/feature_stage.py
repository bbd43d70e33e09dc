"""Per-image feature stage in front of slot attention (reference: ocrs/common/models.py:96-107,
ocrs/common/utils.py:10-33): 4x conv5x5 (ReLU on the first three) and the linear-ramp position
embedding.  Same module/parameter names as the reference.

The convolutions currently call cuDNN through torch (library code); the
position-table add and the NCHW -> token-major transpose are fused into the token-stage kernel.
"""
import torch
import torch.nn as nn

from .networks import Conv2dBlock, conv2d


class SlotAttnCNNEncoder(nn.Module):
    def __init__(self, obs_size, obs_channels, hidden_size):
        super().__init__()
        self._encoder = nn.Sequential(
            Conv2dBlock(obs_channels, hidden_size, 5, 1, 2),
            Conv2dBlock(hidden_size, hidden_size, 5, 1, 2),
            Conv2dBlock(hidden_size, hidden_size, 5, 1, 2),
            conv2d(hidden_size, hidden_size, 5, 1, 2),
        )

    def forward(self, obs):
        return self._encoder(obs)


class PositionalEmbedding(nn.Module):
    def __init__(self, obs_size: int, obs_channels: int):
        super().__init__()
        ramp = torch.linspace(0, 1, obs_size)
        rev = torch.linspace(1, 0, obs_size)
        east = ramp.view(1, -1).expand(obs_size, obs_size)
        west = rev.view(1, -1).expand(obs_size, obs_size)
        south = ramp.view(-1, 1).expand(obs_size, obs_size)
        north = rev.view(-1, 1).expand(obs_size, obs_size)
        grid = torch.stack([north, south, west, east], dim=0).unsqueeze(0).contiguous()
        self.channels_map = nn.Conv2d(4, obs_channels, kernel_size=1)
        self.register_buffer("linear_position_embedding", grid)

    def table(self):
        """The input-independent [C, S*S] table the reference recomputes per batch element."""
        w = self.channels_map.weight.flatten(1)
        g = self.linear_position_embedding[0].flatten(1)
        return w @ g + self.channels_map.bias.unsqueeze(1)

    def forward(self, x):
        B, _, H, W = x.shape
        return x + self.table().view(1, -1, H, W)
