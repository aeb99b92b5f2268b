"""Overlay of the reference's core/model/net.py: AttFlat, Net, Net2, ClassifierNet (+ FCNet).

Class names, constructor signatures, forward arity (Net: 8 outputs, Net2 / ClassifierNet: 5) and
state_dict keys follow /root/reference/core/model/net.py:20-381, so core/exec.py trains,
evaluates, checkpoints and visualises unchanged.  The co-attention backbone (MCA_ED), both
AttFlat poolings, proj_norm and the two big projections (img_feat_linear, proj) run on the
hand-written sm_100a kernels; so do the zero-row mask of the image features (computed in the pass that
casts them to the GEMM operand) and the output head proj_norm -> proj -> sigmoid (-> BCE(sum) through
`forward_with_loss`, what core/exec.py:178 computes with loss_fn), and the question encoder (embedding + LSTM +
token mask: csrc/lstm.cu; cuDNN only for hidden sizes the kernel does not cover and for fp32-grade inference).
"""
import torch
import torch.nn as nn
from torch.nn.utils.weight_norm import weight_norm

from core.model.mca import MCA_ED, MCAClassifier
from core.model.net_utils import FC, MLP, LayerNorm, TCLinear  # noqa: F401
from mcan_vqa_b200 import autograd as _ag
from mcan_vqa_b200.autograd import cfg_get
from mcan_vqa_b200 import blocks as _blocks
from mcan_vqa_b200.blocks import LinearParams, refresh_scope


class AttFlat(nn.Module):
    """Attention pooling over the sequence (reference net.py:20-55): returns (x_atted, att_w)."""

    def __init__(self, opt):
        super(AttFlat, self).__init__()
        self.opt = opt
        self.hidden_size = cfg_get(opt, "hidden_size")
        self.flat_mlp_size = cfg_get(opt, "flat_mlp_size")
        self.flat_glimpses = cfg_get(opt, "flat_glimpses")
        self.flat_out_size = cfg_get(opt, "flat_out_size")
        self.dropout_rate = cfg_get(opt, "dropout_rate")
        self.mlp = MLP(in_size=self.hidden_size, mid_size=self.flat_mlp_size, out_size=self.flat_glimpses,
                       dropout_rate=self.dropout_rate, use_relu=True)
        self.linear_merge = nn.Linear(self.hidden_size * self.flat_glimpses, self.flat_out_size)
        self._lp_merge = None

    def lp_merge(self):
        if self._lp_merge is None:
            self._lp_merge = LinearParams([(self.linear_merge.weight, self.linear_merge.bias)])
        return self._lp_merge

    def all_lps(self):
        return [self.mlp.lp_fc(), self.lp_merge()]

    def forward(self, x, x_mask):
        return _ag.attflat(self, x, x_mask)


def _make_mask(feature):
    """True where a feature row is entirely zero (reference net.py:135-137)."""
    return (feature.abs().sum(dim=-1) == 0).unsqueeze(1).unsqueeze(2)


class _VQABase(nn.Module):
    """Question encoder + image projection + MCA_ED + 2 x AttFlat + classifier head, shared by Net and Net2."""

    def _build(self, opt, pretrained_emb, token_size, answer_size, lstm_dropout):
        H = cfg_get(opt, "hidden_size")
        E = cfg_get(opt, "word_embed_size")
        self.embedding = nn.Embedding(num_embeddings=token_size, embedding_dim=E)
        if cfg_get(opt, "use_glove", False):
            self.embedding.weight.data.copy_(torch.from_numpy(pretrained_emb))
        self.lstm = nn.LSTM(input_size=E, hidden_size=H, num_layers=1, batch_first=True, **lstm_dropout)
        self.img_feat_linear = TCLinear(cfg_get(opt, "img_feat_size"), H)
        self.backbone = MCA_ED(opt)
        self.attflat_img = AttFlat(opt)
        self.attflat_lang = AttFlat(opt)
        self.proj_norm = LayerNorm(cfg_get(opt, "flat_out_size"))
        self.proj = TCLinear(cfg_get(opt, "flat_out_size"), answer_size)

    def lp_lstm_ih(self):
        if getattr(self, "_lp_ih", None) is None:
            self._lp_ih = LinearParams([(self.lstm.weight_ih_l0, self.lstm.bias_ih_l0)], pad=64)
        return self._lp_ih

    def lp_lstm_hh(self):
        if getattr(self, "_lp_hh", None) is None:
            self._lp_hh = LinearParams([(self.lstm.weight_hh_l0, self.lstm.bias_hh_l0)])
        return self._lp_hh

    def all_lps(self):
        return (self.backbone.all_lps() + self.attflat_img.all_lps() + self.attflat_lang.all_lps() +
                [self.img_feat_linear.lp(), self.proj.lp(), self.lp_lstm_ih(), self.lp_lstm_hh()])

    def _features_impl(self, v, ques_ix):
        split = _blocks.PRECISION == "fp32" and not torch.is_grad_enabled()
        if (ques_ix.is_cuda or _blocks.DRY_RUN) and _blocks.lstm_supported(self.lstm, split):
            # embedding + LSTM + make_mask(ques_ix) on the library's persistent LSTM kernels
            q, q_mask = _ag.question_encoder(self, ques_ix)
        else:
            q_mask = _make_mask(ques_ix.unsqueeze(2))
            if split:
                # fp32-grade inference: cuDNN's RNN GEMMs default to TF32 (1e-3); switch that off
                with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                    q, _ = self.lstm(self.embedding(ques_ix))
            else:
                q, _ = self.lstm(self.embedding(ques_ix))
        v, v_mask = _ag.linear_mask(self.img_feat_linear, v)      # img_feat_linear + make_mask(v) in one pass
        q, v = self.backbone(q, v, q_mask, v_mask)
        lang, q_w = self.attflat_lang(q, q_mask)
        img, v_w = self.attflat_img(v, v_mask)
        return q, v, q_mask, v_mask, q_w, v_w, lang, img

    def _features(self, v, ques_ix, target=None):
        with refresh_scope(self.all_lps(), self.training or torch.is_grad_enabled()):
            q, v, q_mask, v_mask, q_w, v_w, lang, img = self._features_impl(v, ques_ix)
            head = _ag.head(self.proj_norm, self.proj, lang, img, target)
        return (q, v, q_mask, v_mask, q_w, v_w) + tuple(head)

    def forward_with_loss(self, v, ques_ix, ans):
        """(BCELoss(reduction='sum')(probs, ans), probs): the loss core/exec.py:178 computes from forward()'s first
        output, fused with the sigmoid (forward) and with the proj gradient operand (backward)."""
        out = self._features(v, ques_ix, ans)
        return out[8], out[7]

    def make_mask(self, feature):
        return _make_mask(feature)


class Net(_VQABase):
    """reference net.py:62-137; forward -> (probs, v, v_mask, v_w, q, q_mask, q_w, a)."""

    def __init__(self, opt, pretrained_emb, token_size, answer_size):
        super(Net, self).__init__()
        self._build(opt, pretrained_emb, token_size, answer_size, {})

    def forward(self, v, ques_ix):
        q, v, q_mask, v_mask, q_w, v_w, a, probs = self._features(v, ques_ix)
        return probs, v, v_mask, v_w, q, q_mask, q_w, a


class Net2(_VQABase):
    """reference net.py:295-381 (same parameters and probabilities as Net); forward -> 5-tuple."""

    def __init__(self, opt, pretrained_emb, token_size, answer_size):
        super(Net2, self).__init__()
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")   # 1-layer LSTM + dropout only warns (net.py:310-316)
            self._build(opt, pretrained_emb, token_size, answer_size, {"dropout": cfg_get(opt, "dropout_rate")})

    def forward(self, v, ques_ix):
        q, v, q_mask, v_mask, _, _, a, probs = self._features(v, ques_ix)
        return probs, v, v_mask, q, q_mask


class ClassifierNet(nn.Module):
    """reference net.py:140-196: image-only SA stack -> AttFlat -> head; forward(v) -> 5-tuple."""

    def __init__(self, opt, answer_size):
        super(ClassifierNet, self).__init__()
        self.img_feat_linear = TCLinear(cfg_get(opt, "img_feat_size"), cfg_get(opt, "hidden_size"))
        self.backbone = MCAClassifier(opt)
        self.attflat_img = AttFlat(opt)
        self.attflat_lang = AttFlat(opt)     # unused but part of the reference state_dict (net.py:150)
        self.proj_norm = LayerNorm(cfg_get(opt, "flat_out_size"))
        self.proj = TCLinear(cfg_get(opt, "flat_out_size"), answer_size)

    def forward(self, v):
        lps = self.backbone.all_lps() + self.attflat_img.all_lps() + [self.img_feat_linear.lp(), self.proj.lp()]
        with refresh_scope(lps, self.training or torch.is_grad_enabled()):
            v, v_mask = _ag.linear_mask(self.img_feat_linear, v)
            v = self.backbone(v, v_mask)
            img, v_w = self.attflat_img(v, v_mask)
            a, probs = _ag.head(self.proj_norm, self.proj, img)
        return probs, v, v_mask, v_w, a

    def make_mask(self, feature):
        return _make_mask(feature)


class FCNet(nn.Module):
    """Weight-normed MLP helper kept for import compatibility (reference net.py:199-225; unused there)."""

    def __init__(self, dims, act='ReLU', dropout=0, bias=True):
        super(FCNet, self).__init__()
        layers = []
        for i, (d_in, d_out) in enumerate(zip(dims[:-1], dims[1:])):
            if dropout > 0:
                layers.append(nn.Dropout(dropout))
            layers.append(weight_norm(nn.Linear(d_in, d_out, bias=bias), dim=None))
            if act:
                layers.append(getattr(nn, act)())
        self.main = nn.Sequential(*layers)

    def forward(self, x):
        return self.main(x)
