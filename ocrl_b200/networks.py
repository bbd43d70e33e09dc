"""Layer factories with the reference's initialisation families
(ocrs/common/networks.py:6-74): xavier-uniform by default, kaiming-uniform(relu) on request,
zero biases; GRUCell with xavier input weights and orthogonal recurrent weights.

The construction order of random draws (torch's default init first, then the re-init) is kept,
so a model built under ``torch.manual_seed(s)`` has the same weights as the reference built
under the same seed -- parity runs use seeded random-init weights.
"""
import torch.nn as nn
import torch.nn.functional as F


def _reinit(weight, weight_init, gain=1.0):
    if weight_init == "kaiming":
        nn.init.kaiming_uniform_(weight, nonlinearity="relu")
    else:
        nn.init.xavier_uniform_(weight, gain)


def linear(in_features, out_features, bias=True, weight_init="xavier", gain=1.0):
    layer = nn.Linear(in_features, out_features, bias)
    _reinit(layer.weight, weight_init, gain)
    if bias:
        nn.init.zeros_(layer.bias)
    return layer


def conv2d(in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
           padding_mode="zeros", weight_init="xavier"):
    layer = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode)
    _reinit(layer.weight, weight_init)
    if bias:
        nn.init.zeros_(layer.bias)
    return layer


class Conv2dBlock(nn.Module):
    """conv (kaiming init) + ReLU; the conv is registered as ``m`` (state_dict key ``*.m.weight``)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0):
        super().__init__()
        self.m = conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=True, weight_init="kaiming")

    def forward(self, x):
        return F.relu(self.m(x))


def gru_cell(input_size, hidden_size, bias=True):
    cell = nn.GRUCell(input_size, hidden_size, bias)
    nn.init.xavier_uniform_(cell.weight_ih)
    nn.init.orthogonal_(cell.weight_hh)
    if bias:
        nn.init.zeros_(cell.bias_ih)
        nn.init.zeros_(cell.bias_hh)
    return cell
