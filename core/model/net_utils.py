"""Overlay of the reference's core/model/net_utils.py (FC, MLP, LayerNorm) on B200 kernels.

Same class names, constructor arguments, sub-module / parameter names and therefore the same
state_dict keys as /root/reference/core/model/net_utils.py:11-60; the arithmetic runs in
libmcan_b200.so (tcgen05 GEMM with fused bias/ReLU/dropout epilogue, LayerNorm kernel).
`core` has no __init__.py on purpose: it is a namespace package, so putting this repo before
the reference on sys.path makes the reference's own core/exec.py import these classes.
"""
import torch
import torch.nn as nn

from mcan_vqa_b200 import autograd as _ag
from mcan_vqa_b200.blocks import LinearParams


class TCLinear(nn.Linear):
    """nn.Linear whose forward/backward run on the tcgen05 GEMM (same parameters, same keys)."""

    def lp(self):
        if getattr(self, "_lp", None) is None:
            self._lp = LinearParams([(self.weight, self.bias)])
        return self._lp

    def forward(self, x):
        return _ag.linear(self, x)


class FC(nn.Module):
    """Linear -> ReLU -> Dropout  (reference net_utils.py:11-34)."""

    def __init__(self, in_size, out_size, dropout_rate=0., use_relu=True):
        super(FC, self).__init__()
        self.in_size, self.out_size = in_size, out_size
        self.dropout_r = dropout_rate
        self.dropout_rate = dropout_rate
        self.use_relu = use_relu
        self.linear = nn.Linear(in_size, out_size)
        if use_relu:
            self.relu = nn.ReLU(inplace=True)
        if dropout_rate > 0:
            self.dropout = nn.Dropout(dropout_rate)
        self._lp = None

    def lp(self):
        if self._lp is None:
            self._lp = LinearParams([(self.linear.weight, self.linear.bias)])
        return self._lp

    def forward(self, x):
        return _ag.fc(self, x)


class MLP(nn.Module):
    """FC -> Linear  (reference net_utils.py:37-45); one GEMM with ReLU/dropout epilogue + one GEMM."""

    def __init__(self, in_size, mid_size, out_size, dropout_rate=0., use_relu=True):
        super(MLP, self).__init__()
        self.in_size, self.mid_size, self.out_size = in_size, mid_size, out_size
        self.dropout_rate = dropout_rate
        self.fc = FC(in_size, mid_size, dropout_rate=dropout_rate, use_relu=use_relu)
        self.linear = nn.Linear(mid_size, out_size)
        self._lp_out = None

    def lp_fc(self):
        return self.fc.lp()

    def lp_out(self):
        if self._lp_out is None:
            self._lp_out = LinearParams([(self.linear.weight, self.linear.bias)])
        return self._lp_out

    def all_lps(self):
        return [self.lp_fc(), self.lp_out()]

    def forward(self, x):
        return _ag.mlp(self, x)


class LayerNorm(nn.Module):
    """MCAN LayerNorm: a_2 * (x - mean) / (std_unbiased + eps) + b_2  (reference net_utils.py:48-60)."""

    def __init__(self, size, eps=1e-6):
        super(LayerNorm, self).__init__()
        self.size = size
        self.eps = eps
        self.a_2 = nn.Parameter(torch.ones(size))
        self.b_2 = nn.Parameter(torch.zeros(size))

    def forward(self, x):
        return _ag.layernorm(self, x)
