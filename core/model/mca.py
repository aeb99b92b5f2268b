"""Overlay of the reference's core/model/mca.py (MHAtt, FFN, SA, SGA, MCA_ED, MCAClassifier).

Module API, constructor signature (`opt` attribute bag), sub-module names and state_dict keys
follow /root/reference/core/model/mca.py:18-207; every forward/backward is a chain of
hand-written sm_100a kernels (mcan-vqa_b200/blocks.py).  There is no PyTorch math fallback:
on a machine without a B200 the forward raises.
"""
import torch.nn as nn

from core.model.net_utils import FC, MLP, LayerNorm  # noqa: F401  (same import surface as the reference)
from mcan_vqa_b200 import autograd as _ag
from mcan_vqa_b200.autograd import cfg_get
from mcan_vqa_b200.blocks import LinearParams


class MHAtt(nn.Module):
    """Multi-head attention, forward(v, k, q, mask) -- note the argument order (reference mca.py:30)."""

    def __init__(self, opt):
        super(MHAtt, self).__init__()
        self.opt = opt
        self.hidden_size = cfg_get(opt, "hidden_size")
        self.multi_head = cfg_get(opt, "multi_head")
        self.head_dim = int(cfg_get(opt, "hidden_size_head", self.hidden_size // self.multi_head))
        self.dropout_rate = cfg_get(opt, "dropout_rate")
        H = self.hidden_size
        self.linear_v = nn.Linear(H, H)
        self.linear_k = nn.Linear(H, H)
        self.linear_q = nn.Linear(H, H)
        self.linear_merge = nn.Linear(H, H)
        self.dropout = nn.Dropout(self.dropout_rate)
        self._lp_qkv = None
        self._lp_merge = None

    def lp_qkv(self):
        """bf16 operand copy of [Wq; Wk; Wv] ([3H, H]) -- one GEMM projects Q, K and V."""
        if self._lp_qkv is None:
            self._lp_qkv = LinearParams([(l.weight, l.bias) for l in (self.linear_q, self.linear_k, self.linear_v)])
        return self._lp_qkv

    def lp_merge(self):
        if self._lp_merge is None:
            self._lp_merge = LinearParams([(self.linear_merge.weight, self.linear_merge.bias)])
        return self._lp_merge

    def all_lps(self):
        return [self.lp_qkv(), self.lp_merge()]

    def forward(self, v, k, q, mask):
        return _ag.mhatt(self, v, k, q, mask)


class FFN(nn.Module):
    """Position-wise feed forward (reference mca.py:85-98)."""

    def __init__(self, opt):
        super(FFN, self).__init__()
        H = cfg_get(opt, "hidden_size")
        self.mlp = MLP(in_size=H, mid_size=cfg_get(opt, "ff_size", 4 * H), out_size=H,
                       dropout_rate=cfg_get(opt, "dropout_rate"), use_relu=True)

    def forward(self, x):
        return self.mlp(x)


class SA(nn.Module):
    """Self-attention layer: x = norm1(x + drop(mhatt(x))); x = norm2(x + drop(ffn(x)))  (mca.py:105-127)."""

    def __init__(self, opt):
        super(SA, self).__init__()
        self.hidden_size = cfg_get(opt, "hidden_size")
        self.dropout_rate = cfg_get(opt, "dropout_rate")
        self.mhatt = MHAtt(opt)
        self.ffn = FFN(opt)
        self.dropout1 = nn.Dropout(self.dropout_rate)
        self.norm1 = LayerNorm(self.hidden_size)
        self.activation = nn.GELU()      # present (unused) in the reference, mca.py:114
        self.dropout2 = nn.Dropout(self.dropout_rate)
        self.norm2 = LayerNorm(self.hidden_size)

    def all_lps(self):
        return self.mhatt.all_lps() + self.ffn.mlp.all_lps()

    def forward(self, x, x_mask):
        return _ag.sa(self, x, x_mask)


class SGA(nn.Module):
    """Self-attention + attention guided by y + FFN, each with post-LN residual  (mca.py:134-164)."""

    def __init__(self, opt):
        super(SGA, self).__init__()
        self.hidden_size = cfg_get(opt, "hidden_size")
        self.dropout_rate = cfg_get(opt, "dropout_rate")
        self.mhatt1 = MHAtt(opt)
        self.mhatt2 = MHAtt(opt)
        self.ffn = FFN(opt)
        self.dropout1 = nn.Dropout(self.dropout_rate)
        self.norm1 = LayerNorm(self.hidden_size)
        self.dropout2 = nn.Dropout(self.dropout_rate)
        self.norm2 = LayerNorm(self.hidden_size)
        self.dropout3 = nn.Dropout(self.dropout_rate)
        self.norm3 = LayerNorm(self.hidden_size)

    def all_lps(self):
        return self.mhatt1.all_lps() + self.mhatt2.all_lps() + self.ffn.mlp.all_lps()

    def forward(self, x, y, x_mask, y_mask):
        return _ag.sga(self, x, y, x_mask, y_mask)


class MCA_ED(nn.Module):
    """Encoder-decoder cascade (mca.py:171-186): L x SA on x, then L x SGA on y guided by the final x."""

    # its backward hands the gradients of each finished layer to dp.layer_hook() (overlapped all-reduce); every
    # other backbone is reduced through per-parameter hooks (mcan_vqa_b200/dp.py GradSync)
    reports_layers = True

    def __init__(self, opt):
        super(MCA_ED, self).__init__()
        self.hidden_size = cfg_get(opt, "hidden_size")
        self.dropout_rate = cfg_get(opt, "dropout_rate")
        layers = cfg_get(opt, "layer")
        self.enc_list = nn.ModuleList([SA(opt) for _ in range(layers)])
        self.dec_list = nn.ModuleList([SGA(opt) for _ in range(layers)])
        self._lp_kv_all = None

    def lp_kv_all(self):
        """[Wk_0; Wv_0; Wk_1; Wv_1; ...] of every decoder's guided attention: all decoder layers read
        the SAME final encoder output (mca.py:183-184), so their K/V projections are one GEMM."""
        if self._lp_kv_all is None:
            pairs = []
            for dec in self.dec_list:
                pairs.append((dec.mhatt2.linear_k.weight, dec.mhatt2.linear_k.bias))
                pairs.append((dec.mhatt2.linear_v.weight, dec.mhatt2.linear_v.bias))
            self._lp_kv_all = LinearParams(pairs)
        return self._lp_kv_all

    def all_lps(self):
        out = [self.lp_kv_all()] if len(self.dec_list) else []
        for layer in list(self.enc_list) + list(self.dec_list):
            out += layer.all_lps()
        return out

    def forward(self, x, y, x_mask, y_mask):
        return _ag.mca_ed(self, x, y, x_mask, y_mask)


class MCAClassifier(nn.Module):
    """SA-only stack over image features for multi-label classification (mca.py:189-207)."""

    def __init__(self, opt):
        super(MCAClassifier, self).__init__()
        self.hidden_size = cfg_get(opt, "hidden_size")
        self.dropout_rate = cfg_get(opt, "dropout_rate")
        self.enc_list = nn.ModuleList([SA(opt) for _ in range(cfg_get(opt, "layer"))])

    def all_lps(self):
        out = []
        for layer in self.enc_list:
            out += layer.all_lps()
        return out

    def forward(self, y, y_mask):
        return _ag.sa_stack(self, y, y_mask)
