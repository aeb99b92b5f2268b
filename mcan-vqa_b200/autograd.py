"""torch.autograd glue: one Function per module-level call, backed by the kernel chains in blocks.py.

Each call builds a small *runner* that knows the module, the shapes and how to run forward and
backward; `_Chain` hands the module parameters to autograd so `.grad`, `named_parameters()`,
`clip_grad_norm_`, AdamW and `state_dict()` of the reference's training loop (core/exec.py)
keep working on ordinary fp32 nn.Parameters.
"""
import torch
from torch.autograd.function import once_differentiable

from . import blocks, dp, ops


def cfg_get(opt, name, default=None):
    """Reads a config attribute the way this fork spells it (lower case, cfgs/base_cfgs.py:122-141)
    and falls back to upstream MCAN's upper-case spelling."""
    for key in (name, name.upper()):
        if hasattr(opt, key):
            return getattr(opt, key)
    if default is not None:
        return default
    raise AttributeError("config has no attribute %r / %r" % (name, name.upper()))


def _as_f32_2d(x, feat):
    if not x.is_cuda and not blocks.DRY_RUN:
        raise ops.capi.McanError("MCAN hot-path modules need CUDA tensors (no CPU fallback); got %s" % x.device)
    x2 = x.detach()
    if x2.dtype != torch.float32:
        x2 = x2.float()
    return x2.contiguous().view(-1, feat)


def _mask_u8(mask, B, S):
    """bool [B,1,1,S] (True = masked, net.py:135-137) -> contiguous uint8 [B,S]; None passes through.
    A mask produced by the library (linear_mask / the question encoder) carries its uint8 storage along."""
    if mask is None:
        return None
    u8 = getattr(mask, "_mcan_u8", None)
    if u8 is not None and u8.numel() == B * S:
        return u8.view(B, S)
    return mask.reshape(B, S).to(torch.uint8).contiguous()


def mask_from_u8(u8, B, S):
    """uint8 [B*S] written by a kernel -> the reference's bool mask [B,1,1,S] (a view of the same bytes)."""
    m = u8.view(torch.bool).view(B, 1, 1, S)
    m._mcan_u8 = u8
    return m


class _Chain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, n_act, *tensors):
        ctx.runner = runner
        ctx.n_act = n_act
        ctx.act_needs = [t is not None and t.requires_grad for t in tensors[:n_act]]
        if runner.sparse_grads:
            ctx.set_materialize_grads(False)    # unused outputs arrive as None instead of zero tensors
        if not blocks.DRY_RUN:
            for t in tensors[:n_act]:
                if t is not None:
                    ops.check_device(t, type(runner.module).__name__ + " input")
        with ops.trusted():
            outs = runner.forward(tensors[:n_act])
        for t in runner.non_differentiable:
            ctx.mark_non_differentiable(t)
        return outs

    @staticmethod
    @once_differentiable
    def backward(ctx, *gouts):
        runner = ctx.runner
        with ops.trusted():
            act_grads, grads = runner.backward(gouts, ctx.act_needs)
        # a gradient in another dtype than its parameter (bf16 weight gradients produced directly in the all-reduce
        # buffer, blocks.WGRAD_BF16) is not handed to autograd: the data-parallel optimiser gets it from dp.GradSync
        pg = tuple(g if (g is None or g.dtype == p.dtype) else None for p, g in ((p, grads.get(p)) for p in runner.params))
        return (None, None) + tuple(act_grads) + pg


class _Runner(object):
    non_differentiable = ()
    sparse_grads = False

    def __init__(self, module, params):
        self.module = module
        self.params = params
        self.rt = blocks.Runtime(module.training, getattr(module, "dropout_rate", 0.0))


def _run(runner, acts):
    return _Chain.apply(runner, len(acts), *(tuple(acts) + tuple(runner.params)))


# ------------------------------------------------------------------------------------------
class _LayerNormRunner(_Runner):
    def forward(self, acts):
        (x,) = acts
        m = self.module
        self.shape = x.shape
        s = _as_f32_2d(x, m.size)
        out, self.mean, self.sigma = blocks.ln_fwd(m, s, want_bf=False)
        self.s = s
        return out.f32.view(self.shape)

    def backward(self, gouts, needs):
        m = self.module
        dy = _as_f32_2d(gouts[0], m.size)
        dx, _, da2, db2 = blocks.ln_bwd(self.rt, m, dy, self.s, self.mean, self.sigma, want_bf=False)
        return [dx.view(self.shape)], {m.a_2: da2, m.b_2: db2}


def layernorm(module, x):
    return _run(_LayerNormRunner(module, [module.a_2, module.b_2]), [x])


# ------------------------------------------------------------------------------------------
class _MHAttRunner(_Runner):
    """MHAtt.forward(v, k, q, mask) -- mca.py:30-63; inputs that are the same tensor share one GEMM."""

    def __init__(self, module, same_vk, same_kq):
        _Runner.__init__(self, module, blocks.module_params(module))
        self.same_vk, self.same_kq = same_vk, same_kq

    def forward(self, acts):
        v, k, q, mask = acts
        m = self.module
        H = m.hidden_size
        B, Sq, Sk = q.shape[0], q.shape[1], k.shape[1]
        self.B, self.Sq, self.Sk = B, Sq, Sk
        qa = blocks.act_from_f32(_as_f32_2d(q, H), self.rt.split)
        mu8 = _mask_u8(mask, B, Sk)
        if self.same_vk and self.same_kq:
            out, self.c = blocks.att_fwd(self.rt, m, qa, B, Sq, key_mask=mu8)
        else:
            ka = blocks.act_from_f32(_as_f32_2d(k, H), self.rt.split)
            va = None if self.same_vk else blocks.act_from_f32(_as_f32_2d(v, H), self.rt.split)
            out, self.c = blocks.att_fwd(self.rt, m, qa, B, Sq, kv_src=ka, Sk=Sk, key_mask=mu8, v_src=va)
        return out.view(B, Sq, H)

    def backward(self, gouts, needs):
        m = self.module
        H = m.hidden_size
        dout = _as_f32_2d(gouts[0], H)
        dx, dk_src, dv_src, grads = blocks.att_bwd(self.rt, m, self.c, dout)
        B, Sq, Sk = self.B, self.Sq, self.Sk
        if self.same_vk and self.same_kq:
            # one tensor was passed three times: autograd sums the three slots
            return [None, None, dx.view(B, Sq, H), None], grads
        dq = dx.view(B, Sq, H)
        if self.same_vk:
            return [None, dk_src.view(B, Sk, H), dq, None], grads
        return [dv_src.view(B, Sk, H), dk_src.view(B, Sk, H), dq, None], grads


def mhatt(module, v, k, q, mask):
    return _run(_MHAttRunner(module, v is k, k is q), [v, k, q, mask])


# ------------------------------------------------------------------------------------------
class _MLPRunner(_Runner):
    """MLP / FFN: Linear -> ReLU -> dropout -> Linear (net_utils.py:37-45, mca.py:85-98)."""

    def forward(self, acts):
        (x,) = acts
        mlp = self.module
        self.shape = x.shape
        xa = blocks.act_from_f32(_as_f32_2d(x, mlp.in_size), self.rt.split)
        out, self.c = blocks.mlp_fwd(self.rt, mlp, xa)
        return out.contiguous().view(self.shape[:-1] + (mlp.out_size,))

    def backward(self, gouts, needs):
        mlp = self.module
        dout = gouts[0].reshape(-1, mlp.out_size)
        dx, grads = blocks.mlp_bwd(self.rt, mlp, self.c, dout)
        return [dx.view(self.shape)], grads


def mlp(module, x):
    return _run(_MLPRunner(module, blocks.module_params(module)), [x])


class _FCRunner(_Runner):
    """FC: Linear -> ReLU -> dropout (net_utils.py:11-34), fp32 in / fp32 out."""

    def forward(self, acts):
        (x,) = acts
        fc = self.module
        self.shape = x.shape
        xa = blocks.act_from_f32(_as_f32_2d(x, fc.in_size))
        lp = fc.lp().get()
        M = xa.bf.shape[0]
        p = self.rt.p if fc.dropout_r > 0 else 0.0
        self.p, self.seed = p, self.rt.seed()
        out32 = torch.empty((M, lp.n), dtype=torch.float32, device=x.device)
        outbf = torch.empty((M, lp.n), dtype=torch.bfloat16, device=x.device)
        ops.gemm(xa.bf, lp.w, bias=lp.b, relu=fc.use_relu, dropout_p=p, seed=self.seed, out_f32=out32, out_bf16=outbf)
        self.xbf, self.lp, self.outbf = xa.bf, lp, outbf
        return out32.view(self.shape[:-1] + (lp.n,))

    def backward(self, gouts, needs):
        fc = self.module
        lp = self.lp
        dev = gouts[0].device
        M = self.xbf.shape[0]
        dout = gouts[0].reshape(M, lp.n).contiguous()
        # gate the incoming gradient through ReLU/dropout with the saved activation
        gated = torch.empty((M, lp.n), dtype=torch.bfloat16, device=dev)
        if fc.use_relu:
            ops.gate_bf16(dout, self.outbf, 1.0 / (1.0 - self.p) if self.p > 0 else 1.0, gated)
        elif self.p > 0:
            raise ops.capi.McanError("FC with dropout but no ReLU is not supported")
        else:
            ops.cast_bf16(dout, gated)
        g = blocks.GradBuf(self.rt, lp)
        ops.colsum(gated, g.b)
        ops.gemm(gated, self.xbf, a_layout=1, b_layout=1, out_f32=g.w, accumulate=True)
        dx = torch.empty((M, lp.k), dtype=torch.float32, device=dev)
        ops.gemm(gated, lp.w, b_layout=1, out_f32=dx)
        (gw, gb), = g.per_param()
        return [dx.view(self.shape)], {fc.linear.weight: gw, fc.linear.bias: gb}


def fc(module, x):
    return _run(_FCRunner(module, blocks.module_params(module)), [x])


# ------------------------------------------------------------------------------------------
class _SARunner(_Runner):
    def forward(self, acts):
        x, mask = acts
        m = self.module
        H = m.hidden_size
        self.B, self.S = x.shape[0], x.shape[1]
        xa = blocks.act_from_f32(_as_f32_2d(x, H), self.rt.split)
        out, self.c = blocks.sa_fwd(self.rt, m, xa, self.B, self.S, _mask_u8(mask, self.B, self.S))
        return out.f32.view(self.B, self.S, H)

    def backward(self, gouts, needs):
        m = self.module
        dz = _as_f32_2d(gouts[0], m.hidden_size)
        dx, grads = blocks.sa_bwd(self.rt, m, self.c, dz)
        return [dx.view(self.B, self.S, m.hidden_size), None], grads


def sa(module, x, x_mask):
    return _run(_SARunner(module, blocks.module_params(module)), [x, x_mask])


class _SGARunner(_Runner):
    def forward(self, acts):
        x, y, x_mask, y_mask = acts
        m = self.module
        H = m.hidden_size
        self.B, self.Sx, self.Sy = x.shape[0], x.shape[1], y.shape[1]
        xa = blocks.act_from_f32(_as_f32_2d(x, H), self.rt.split)
        ya = blocks.act_from_f32(_as_f32_2d(y, H), self.rt.split)
        out, self.c = blocks.sga_fwd(self.rt, m, xa, ya, self.B, self.Sx, self.Sy,
                                     _mask_u8(x_mask, self.B, self.Sx), _mask_u8(y_mask, self.B, self.Sy))
        return out.f32.view(self.B, self.Sx, H)

    def backward(self, gouts, needs):
        m = self.module
        H = m.hidden_size
        dz = _as_f32_2d(gouts[0], H)
        dx, dy, grads = blocks.sga_bwd(self.rt, m, self.c, dz)
        return [dx.view(self.B, self.Sx, H), dy.view(self.B, self.Sy, H), None, None], grads


def sga(module, x, y, x_mask, y_mask):
    return _run(_SGARunner(module, blocks.module_params(module)), [x, y, x_mask, y_mask])


# ------------------------------------------------------------------------------------------
class _MCAEDRunner(_Runner):
    """MCA_ED.forward(x, y, x_mask, y_mask) -- mca.py:178-186 -- as one kernel chain."""

    def forward(self, acts):
        x, y, x_mask, y_mask = acts
        m = self.module
        H = m.hidden_size
        self.B, self.Sx, self.Sy = x.shape[0], x.shape[1], y.shape[1]
        xo, yo, self.c = blocks.mca_ed_fwd(self.rt, m, _as_f32_2d(x, H), _as_f32_2d(y, H), self.B, self.Sx, self.Sy,
                                           _mask_u8(x_mask, self.B, self.Sx), _mask_u8(y_mask, self.B, self.Sy))
        return xo.view(self.B, self.Sx, H), yo.view(self.B, self.Sy, H)

    def backward(self, gouts, needs):
        m = self.module
        H = m.hidden_size
        dev = m.enc_list[0].norm1.a_2.device if len(m.enc_list) else gouts[0].device
        gx, gy = gouts
        dxo = _as_f32_2d(gx, H) if gx is not None else torch.zeros((self.B * self.Sx, H), device=dev)
        dyo = _as_f32_2d(gy, H) if gy is not None else torch.zeros((self.B * self.Sy, H), device=dev)
        hook = dp.layer_hook()
        if hook is None:
            from . import optim as _optim
            hook = _optim.early_hook()
        dx, dy, grads = blocks.mca_ed_bwd(self.rt, m, self.c, dxo, dyo, after_layer=hook)
        return [dx.view(self.B, self.Sx, H), dy.view(self.B, self.Sy, H), None, None], grads


def mca_ed(module, x, y, x_mask, y_mask):
    return _run(_MCAEDRunner(module, blocks.module_params(module)), [x, y, x_mask, y_mask])


class _SAStackRunner(_Runner):
    """MCAClassifier: SA-only stack over the image features (mca.py:200-207)."""

    def forward(self, acts):
        y, mask = acts
        m = self.module
        H = m.hidden_size
        self.B, self.S = y.shape[0], y.shape[1]
        mu8 = _mask_u8(mask, self.B, self.S)
        a = blocks.act_from_f32(_as_f32_2d(y, H), self.rt.split)
        self.ctxs = []
        for enc in m.enc_list:
            a, c = blocks.sa_fwd(self.rt, enc, a, self.B, self.S, mu8)
            self.ctxs.append(c)
        return a.f32.view(self.B, self.S, H)

    def backward(self, gouts, needs):
        m = self.module
        d = _as_f32_2d(gouts[0], m.hidden_size)
        grads = {}
        for enc, c in zip(reversed(list(m.enc_list)), reversed(self.ctxs)):
            d, g = blocks.sa_bwd(self.rt, enc, c, d)
            grads.update(g)
        return [d.view(self.B, self.S, m.hidden_size), None], grads


def sa_stack(module, y, y_mask):
    return _run(_SAStackRunner(module, blocks.module_params(module)), [y, y_mask])


# ------------------------------------------------------------------------------------------
class _AttFlatRunner(_Runner):
    def forward(self, acts):
        x, mask = acts
        m = self.module
        H = m.hidden_size
        self.B, self.S = x.shape[0], x.shape[1]
        xa = blocks.act_from_f32(_as_f32_2d(x, H), self.rt.split)
        out, att_w, self.c = blocks.attflat_fwd(self.rt, m, xa, self.B, self.S, _mask_u8(mask, self.B, self.S))
        self.non_differentiable = (att_w,)
        return out, att_w

    def backward(self, gouts, needs):
        m = self.module
        dx, grads = blocks.attflat_bwd(self.rt, m, self.c, gouts[0])
        return [dx.view(self.B, self.S, m.hidden_size), None], grads


def attflat(module, x, x_mask):
    return _run(_AttFlatRunner(module, blocks.module_params(module)), [x, x_mask])


# ------------------------------------------------------------------------------------------
class _LinearRunner(_Runner):
    """nn.Linear on the tcgen05 GEMM (img_feat_linear net.py:107, proj net.py:129)."""

    def forward(self, acts):
        (x,) = acts
        lin = self.module
        self.shape = x.shape
        lp = lin.lp().get(blocks._force(self.rt.grad), self.rt.split)
        out, self.c = blocks.linear_fwd(lp, _as_f32_2d(x, lp.k), self.rt.split)
        return out.contiguous().view(self.shape[:-1] + (lp.n,)) if out.stride(0) != lp.n else out.view(self.shape[:-1] + (lp.n,))

    def backward(self, gouts, needs):
        lin = self.module
        dout = gouts[0].reshape(-1, self.c.lp.n)
        dx, gw, gb = blocks.linear_bwd(self.rt, self.c, dout, need_dx=needs[0])
        return [dx.view(self.shape) if dx is not None else None], {lin.weight: gw, lin.bias: gb}


def linear(module, x):
    return _run(_LinearRunner(module, [module.weight, module.bias]), [x])


class _LinearMaskRunner(_LinearRunner):
    """img_feat_linear together with make_mask of its input (net.py:100,107,135-137): the fp32 features are read once."""

    def forward(self, acts):
        (x,) = acts
        lin = self.module
        self.shape = x.shape
        lp = lin.lp().get(blocks._force(self.rt.grad), self.rt.split)
        x2 = _as_f32_2d(x, lp.k)
        mask = torch.empty(x2.shape[0], dtype=torch.uint8, device=x.device)
        out, self.c = blocks.linear_fwd(lp, x2, self.rt.split, mask_out=mask)
        self.non_differentiable = (mask,)
        out = out.contiguous().view(self.shape[:-1] + (lp.n,)) if out.stride(0) != lp.n else out.view(self.shape[:-1] + (lp.n,))
        return out, mask


def linear_mask(module, x):
    """-> (module(x), bool mask [B,1,1,S] of all-zero rows of x)."""
    out, u8 = _run(_LinearMaskRunner(module, [module.weight, module.bias]), [x])
    return out, mask_from_u8(u8, x.shape[0], x.shape[1])


# ------------------------------------------------------------------------------------------
class _HeadRunner(_Runner):
    """proj_norm(lang [+ img]) -> proj -> sigmoid [-> BCELoss(reduction='sum')] as one kernel chain
    (net.py:125-129 / 182-184; the loss of exec.py:67,178 when a target is given)."""
    sparse_grads = True

    def __init__(self, norm, proj):
        _Runner.__init__(self, proj, [norm.a_2, norm.b_2, proj.weight, proj.bias])
        self.norm = norm

    def forward(self, acts):
        x, x2, target = acts
        proj = self.module
        lp = proj.lp().get(blocks._force(self.rt.grad), self.rt.split)
        self.shape = x.shape
        xs = _as_f32_2d(x, lp.k)
        x2s = _as_f32_2d(x2, lp.k) if x2 is not None else None
        tg = None
        if target is not None:
            tg = _as_f32_2d(target, lp.n)
        a, probs, loss, self.c = blocks.head_fwd(self.rt, self.norm, lp, xs, x2s, tg)
        self.has_x2 = x2 is not None
        lead = self.shape[:-1]
        a, probs = a.view(lead + (lp.k,)), probs.view(lead + (lp.n,))
        if loss is None:
            return a, probs
        self.non_differentiable = (probs,)
        return a, probs, loss

    def backward(self, gouts, needs):
        proj = self.module
        lp = self.c.lp
        g_a = _as_f32_2d(gouts[0], lp.k) if gouts[0] is not None else None
        if self.c.target is not None:
            g_probs, g_loss = None, gouts[2]
            if g_loss is None:
                g_loss = torch.zeros((), dtype=torch.float32, device=self.c.probs.device)
            g_loss = g_loss.detach().float().contiguous()
        else:
            g_loss = None
            g_probs = gouts[1]
            if g_probs is None:
                g_probs = torch.zeros_like(self.c.probs)
            g_probs = _as_f32_2d(g_probs, lp.n)
        ds, gw, gb, da2, db2 = blocks.head_bwd(self.rt, self.norm, self.c, g_a, g_probs, g_loss)
        ds = ds.view(self.shape)
        return [ds, ds if self.has_x2 else None, None], {self.norm.a_2: da2, self.norm.b_2: db2, proj.weight: gw, proj.bias: gb}


def head(norm, proj, x, x2=None, target=None):
    """-> (proj_feat, probs) or, with a target, (proj_feat, probs, BCE-sum loss)."""
    return _run(_HeadRunner(norm, proj), [x, x2, target])


# ------------------------------------------------------------------------------------------
class _QuestionEncoderRunner(_Runner):
    """embedding -> LSTM (net.py:103-104) and make_mask(ques_ix) (net.py:99) on the library's kernels (csrc/lstm.cu)."""

    def __init__(self, net):
        lstm = net.lstm
        _Runner.__init__(self, net, [net.embedding.weight, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0,
                                     lstm.bias_hh_l0])

    def forward(self, acts):
        (tokens,) = acts
        net = self.module
        force = blocks._force(self.rt.grad)
        lp_ih, lp_hh = net.lp_lstm_ih().get(force, blocks.LSTM_SPLIT_INPUT), net.lp_lstm_hh().get(force)
        tok = tokens.detach().to(torch.int64).contiguous()
        q, mask, self.c = blocks.qenc_fwd(self.rt, net.embedding.weight.detach(), lp_ih, lp_hh, tok, self.rt.grad)
        self.non_differentiable = (mask,)
        B, T = tok.shape
        return q.view(B, T, lp_hh.k), mask

    def backward(self, gouts, needs):
        net = self.module
        H = self.c.lp_hh.k
        dq = _as_f32_2d(gouts[0], H)
        dt, dwi, dbi, dwh, dbh = blocks.qenc_bwd(self.rt, self.c, dq, net.embedding.weight.shape[0])
        lstm = net.lstm
        return [None], {net.embedding.weight: dt, lstm.weight_ih_l0: dwi, lstm.bias_ih_l0: dbi, lstm.weight_hh_l0: dwh,
                        lstm.bias_hh_l0: dbh}


def question_encoder(net, ques_ix):
    """-> (lstm(embedding(ques_ix)) fp32 [B, T, H], bool mask [B,1,1,T] of padding tokens)."""
    q, u8 = _run(_QuestionEncoderRunner(net), [ques_ix])
    return q, mask_from_u8(u8, ques_ix.shape[0], ques_ix.shape[1])
