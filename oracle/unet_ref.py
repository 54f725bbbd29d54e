"""ORACLE (test infrastructure only -- never on the product path).

CPU restatement, in plain functional torch (fp32 or fp64), of the reference denoiser
`ml_mdm.models.unet.UNet` / `ml_mdm.models.nested_unet.NestedUNet` forward.  It consumes the
reference's own ``state_dict`` (same keys, OIHW fp32 weights) and a duck-typed config object with the
reference's field names (UNetConfig, models/unet.py:62-156; NestedUNetConfig, nested_unet.py:21-51).
Backward is obtained with torch autograd, exactly as the reference obtains it (trainer.py:46,75).

Pinned against: the unmodified reference run on CPU in the authoring container (tests/test_oracle.py,
which imports /root/reference when present) and the golden fixtures under tests/golden/ that were
generated from the reference by tests/golden/make_golden.py.  The reference's own test-suite holds
no known-answer vectors for this path, so those two pins are the anchor.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.
"""
import math

import torch
import torch.nn.functional as F


def _get(cfg, name, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


def _int_list(v, n=None):
    if isinstance(v, str):
        v = [int(x) for x in v.split(",")] if len(v) else []
    v = list(v)
    if n is not None and len(v) == 1:
        v = v * n
    return v


class UNetPlan:
    """Static structure of one U-Net level, derived from the config the way UNet.__init__ does
    (models/unet.py:581-773)."""

    def __init__(self, cfg, in_ch=3, out_ch=3, lm_dim=None):
        rc = _get(cfg, "resnet_config")
        self.cfg = cfg
        self.in_ch, self.out_ch = in_ch, out_ch
        self.channels = _int_list(_get(cfg, "resolution_channels"))
        L = len(self.channels)
        self.nres = _int_list(_get(cfg, "num_resnets_per_resolution"), L)
        self.attn_levels = _int_list(_get(cfg, "attention_levels") or [])
        self.nattn = _int_list(_get(cfg, "num_attention_layers"), L)
        self.groups = _get(rc, "num_groups_norm", 32)
        self.use_ffn = bool(_get(rc, "use_attention_ffn", False))
        td = _get(cfg, "temporal_dim")
        self.temporal_dim = self.channels[0] * 4 if td is None else td  # unet.py:594-598
        cdim = _get(cfg, "conditioning_feature_dim", -1) if lm_dim is None else lm_dim
        pdim = _get(cfg, "conditioning_feature_proj_dim", -1)
        self.lm_dim = cdim
        self.has_lm_proj = cdim > 0 and pdim > 0 and not _get(cfg, "skip_cond_emb", False)
        self.cond_dim = pdim if (cdim > 0 and pdim > 0) else cdim  # unet.py:588-593
        self.has_cond_emb = self.cond_dim > 0 and not _get(cfg, "skip_cond_emb", False)
        self.masked_cross_attention = _get(cfg, "masked_cross_attention", 1)
        self.skip_mid = bool(_get(cfg, "skip_mid_blocks", False))
        self.nesting = bool(_get(cfg, "nesting", False))
        mc = _get(cfg, "micro_conditioning")
        self.conditions = None
        if mc is not None:
            self.conditions = {c.split(":")[0]: float(c.split(":")[1]) for c in mc.split(",")}
        # block structure
        ch = self.channels[0]
        skips = [ch]
        self.down = []
        for i in range(L):
            res = []
            for _ in range(self.nres[i]):
                res.append((ch, self.channels[i]))
                ch = self.channels[i]
                skips.append(ch)
            if i != L - 1:
                skips.append(ch)
            na = self.nattn[i] if i in self.attn_levels else 0
            self.down.append(dict(res=res, nattn=na, cond=(i in self.attn_levels), down=(i != L - 1),
                                  up=False))
        self.mid_ch = ch
        self.up = []
        for i in reversed(range(L)):
            res = []
            for _ in range(self.nres[i] + 1):
                res.append((ch + skips.pop(), self.channels[i]))
                ch = self.channels[i]
            na = self.nattn[i] if i in self.attn_levels else 0
            self.up.append(dict(res=res, nattn=na, cond=(i in self.attn_levels), down=False,
                                up=(i != 0), level=i))
        self.out_feat_ch = ch


def _gn(x, P, pre, groups):
    return F.group_norm(x, groups, P[pre + ".weight"], P[pre + ".bias"], eps=1e-5)


def resnet(P, pre, x, temb, cin, cout, groups):
    """ResNet.forward (unet.py:223-238)."""
    h = F.conv2d(F.silu(_gn(x, P, pre + ".norm1", groups)), P[pre + ".conv1.weight"],
                 P[pre + ".conv1.bias"], padding=1)
    t = F.linear(F.silu(temb), P[pre + ".time_layer.weight"], P[pre + ".time_layer.bias"])
    ta, tb = t[:, :cout, None, None], t[:, cout:, None, None]
    if h.shape[0] > ta.shape[0]:
        n = h.shape[0] // ta.shape[0]
        ta = ta.repeat_interleave(n, 0)
        tb = tb.repeat_interleave(n, 0)
    h = F.silu(_gn(h, P, pre + ".norm2", groups) * (1 + ta) + tb)
    h = F.conv2d(h, P[pre + ".conv2.weight"], P[pre + ".conv2.bias"], padding=1)
    if cin != cout:
        x = F.conv2d(x, P[pre + ".conv3.weight"], P[pre + ".conv3.bias"])
    return h + x


def _attend(q, k, v, heads, mask=None):
    """SelfAttention.attention (unet.py:276-294): q (B,C,T), k,v (B,C,S)."""
    b, c, t = q.shape
    d = c // heads
    scale = 1.0 / math.sqrt(math.sqrt(d))
    qh = (q * scale).reshape(b * heads, d, t)
    kh = (k * scale).reshape(b * heads, d, -1)
    w = torch.einsum("bct,bcs->bts", qh, kh)
    if mask is not None:
        m = mask.view(b, 1, 1, -1).expand(b, heads, 1, mask.shape[1]).reshape(b * heads, 1, -1)
        w = w.masked_fill(m == 0, float("-inf"))
    w = torch.softmax(w.float() if w.dtype != torch.float64 else w, dim=-1).to(w.dtype)
    a = torch.einsum("bts,bcs->bct", w, v.reshape(b * heads, d, -1))
    return a.reshape(b, c, t)


def attention_block(P, pre, x, cond, cond_mask, has_cond, use_ffn, heads=8):
    """SelfAttention.forward (unet.py:296-313)."""
    b, c, hh, ww = x.shape
    qkv = F.conv2d(_gn(x, P, pre + ".norm", 32), P[pre + ".qkv.weight"], P[pre + ".qkv.bias"])
    q, k, v = qkv.reshape(b, 3 * c, -1).chunk(3, dim=1)
    h = _attend(q, k, v, heads)
    if has_cond:
        cn = F.layer_norm(cond, (cond.shape[-1],), P[pre + ".norm_cond.weight"],
                          P[pre + ".norm_cond.bias"], eps=1e-5)
        kv = F.linear(cn, P[pre + ".kv_cond.weight"], P[pre + ".kv_cond.bias"]).transpose(-2, -1)
        kc, vc = kv.chunk(2, dim=1)
        h = h + _attend(q, kc, vc, heads, cond_mask)
    h = F.conv2d(h.reshape(b, c, hh, ww), P[pre + ".proj_out.weight"], P[pre + ".proj_out.bias"])
    x = x + h
    if use_ffn:
        f = F.conv2d(_gn(x, P, pre + ".ffn.0", 32), P[pre + ".ffn.1.weight"], P[pre + ".ffn.1.bias"])
        f = F.conv2d(F.gelu(f), P[pre + ".ffn.3.weight"], P[pre + ".ffn.3.bias"])
        x = f + x
    return x


def res_block(P, pre, blk, plan, x, temb, cond, cond_mask, skips=None, trace=None):
    """ResNetBlock.forward (unet.py:534-576). Returns (x, activations)."""
    acts = []
    use_cond = blk["cond"] and plan.cond_dim > 0
    for i, (cin, cout) in enumerate(blk["res"]):
        if skips is not None:
            x = torch.cat((x, skips.pop(0)), dim=1)
        x = resnet(P, f"{pre}.resnets.{i}", x, temb, cin, cout, plan.groups)
        for j in range(blk["nattn"]):
            x = attention_block(P, f"{pre}.attn.{i * blk['nattn'] + j}", x, cond, cond_mask, use_cond,
                                plan.use_ffn)
        if trace is not None:
            trace[f"{pre}.{i}"] = x
        acts.append(x)
    if blk["down"]:
        x = F.conv2d(x, P[pre + ".resample.weight"], P[pre + ".resample.bias"], stride=2, padding=1)
        acts.append(x)
    elif blk["up"]:
        x = F.interpolate(x, scale_factor=2)  # nearest
        x = F.conv2d(x, P[pre + ".resample.weight"], P[pre + ".resample.bias"], padding=1)
        acts.append(x)
    return x, acts


def time_embedding(P, plan, values, l1, l2):
    """UNet.create_temporal_embedding (unet.py:834-845) with the t_emb buffer of :600-603."""
    half = plan.temporal_dim // 8
    dt = P[l1 + ".weight"].dtype
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / half)).to(dt).to(values.device)
    e = values.view(-1, 1).to(dt) * freq.unsqueeze(0)
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    e = F.linear(e, P[l1 + ".weight"], P[l1 + ".bias"])
    return F.linear(F.silu(e), P[l2 + ".weight"], P[l2 + ".bias"])


def micro_embedding(P, pre, plan, times, micros):
    """UNet.forward_micro_conditioning (unet.py:920-933)."""
    out = 0
    for key, default in plan.conditions.items():
        dt = P[f"{pre}cond_layers.{key}.0.weight"].dtype
        m = micros.get(key, default * torch.ones_like(times)) if micros else default * torch.ones_like(times)
        m = m.to(dt)
        m = (m / default).clamp(max=1) * default if key == "scale" else m * 1000
        out = out + time_embedding(P, plan, m, f"{pre}cond_layers.{key}.0", f"{pre}cond_layers.{key}.1")
    return out


def forward_conditioning(P, pre, plan, lm, mask):
    """UNet.forward_conditioning (unet.py:847-865); num_lm_head_layers == 0 in all shipped configs."""
    cond = lm
    if plan.has_lm_proj:
        cond = F.linear(cond, P[pre + "lm_proj.weight"], P[pre + "lm_proj.bias"])
    if mask is None:
        y = cond.mean(dim=1)
    else:
        y = (mask.unsqueeze(-1) * cond).sum(dim=1) / mask.sum(dim=1, keepdim=True)
    if not plan.masked_cross_attention:
        mask = None
    emb = F.linear(y, P[pre + "cond_emb.weight"])
    return emb, cond, mask


def _temb(P, pre, plan, times, cond_emb, micros):
    temb = time_embedding(P, plan, times, pre + "temb_layer1", pre + "temb_layer2")
    if cond_emb is not None:
        temb = temb + cond_emb
    if plan.conditions is not None:
        temb = temb + micro_embedding(P, pre, plan, times, micros)
    return temb


def _down_up(P, pre, plan, x, temb, cond, cond_mask, mid_fn, trace):
    skips = [x]
    for i, blk in enumerate(plan.down):
        x, acts = res_block(P, f"{pre}down_blocks.{i}", blk, plan, x, temb, cond, cond_mask, trace=trace)
        skips.extend(acts)
    x = mid_fn(x)
    for i, blk in enumerate(plan.up):
        n = len(blk["res"])
        sk = skips[-n:][::-1]
        del skips[-n:]
        x, _ = res_block(P, f"{pre}up_blocks.{i}", blk, plan, x, temb, cond, cond_mask, skips=sk,
                         trace=trace)
    return x


def unet_denoise(P, pre, plan, x_t, times, cond_emb, cond, cond_mask, micros, trace=None):
    """UNet.forward_denoising (unet.py:935-969)."""
    temb = _temb(P, pre, plan, times, cond_emb, micros)
    x_feat = None
    if plan.nesting:
        x_t, x_feat = x_t
    if isinstance(x_t, (list, tuple)) and len(x_t) == 1:
        x_t = x_t[0]
    x = F.conv2d(x_t, P[pre + "conv_in.weight"], P[pre + "conv_in.bias"], padding=1)
    if plan.nesting:
        x = x + x_feat
    if trace is not None:
        trace[pre + "conv_in"] = x

    def mid(x):
        if plan.skip_mid:
            return x
        blk0 = dict(res=[(plan.mid_ch, plan.mid_ch)], nattn=1, cond=True, down=False, up=False)
        blk1 = dict(res=[(plan.mid_ch, plan.mid_ch)], nattn=0, cond=False, down=False, up=False)
        x, _ = res_block(P, pre + "mid_blocks.0", blk0, plan, x, temb, cond, cond_mask, trace=trace)
        x, _ = res_block(P, pre + "mid_blocks.1", blk1, plan, x, temb, cond, cond_mask, trace=trace)
        return x

    x = _down_up(P, pre, plan, x, temb, cond, cond_mask, mid, trace)
    out = F.conv2d(F.silu(_gn(x, P, pre + "norm_out", plan.groups)), P[pre + "conv_out.weight"],
                   P[pre + "conv_out.bias"], padding=1)
    if plan.nesting:
        return out, x
    return out


class OracleNet:
    """Whole (nested) denoiser: `forward` mirrors UNet.forward / NestedUNet.forward
    (unet.py:971-987, nested_unet.py:165-230)."""

    def __init__(self, cfg, lm_dim):
        self.levels = []  # outermost first; each (prefix, plan, cfg)
        pre, c = "", cfg
        while True:
            inner = _get(c, "inner_config")
            plan = UNetPlan(c, lm_dim=lm_dim)
            self.levels.append((pre, plan, c))
            if inner is None:
                break
            pre, c = pre + "inner_unet.", inner
        self.nested = len(self.levels) > 1

    def forward(self, P, x_t, times, lm, lm_mask, micros=None, trace=None):
        ipre, iplan, _ = self.levels[-1]
        cond_emb, cond, cmask = None, lm, lm_mask
        if iplan.cond_dim > 0:
            cond_emb, cond, cmask = forward_conditioning(P, ipre, iplan, lm, lm_mask)
        if not self.nested:
            return unet_denoise(P, "", iplan, x_t, times, cond_emb, cond, cmask, micros, trace)
        return self._nested(0, P, x_t, None, times, cond_emb, cond, cmask, micros, trace)

    def _nested(self, li, P, x_list, x_feat, times, cond_emb, cond, cmask, micros, trace):
        """NestedUNet.forward_denoising (nested_unet.py:168-230)."""
        pre, plan, cfg = self.levels[li]
        if li == len(self.levels) - 1:
            return unet_denoise(P, pre, plan, (x_list, x_feat), times, cond_emb, cond, cmask, micros, trace)
        temb = _temb(P, pre, plan, times, cond_emb, micros)
        bh, bl = x_list[0].shape[0], x_list[1].shape[0]
        x_hi, x_low = x_list[0], list(x_list[1:])
        if not _get(cfg, "skip_normalization", False):
            x_hi = x_hi / x_hi.std((1, 2, 3), keepdim=True)
        x = F.conv2d(x_hi, P[pre + "conv_in.weight"], P[pre + "conv_in.bias"], padding=1)
        if plan.nesting:
            x = x + x_feat
        cm = cmask[:bh] if cmask is not None else None
        hold = {}

        def mid(x):
            xi = F.conv2d(x, P[pre + "in_adapter.weight"], P[pre + "in_adapter.bias"], padding=1)
            if bh < bl:
                xi = torch.cat([xi, xi.new_zeros(bl - bh, *xi.shape[1:])], 0)
            low, feat = self._nested(li + 1, P, x_low, xi, times, cond_emb, cond, cmask, micros, trace)
            hold["low"] = low
            feat = F.conv2d(feat, P[pre + "out_adapter.weight"], P[pre + "out_adapter.bias"], padding=1)
            return x + (feat[:bh] if bh < bl else feat)

        x = _down_up(P, pre, plan, x, temb[:bh], cond[:bh], cm, mid, trace)
        out = F.conv2d(F.silu(_gn(x, P, pre + "norm_out", plan.groups)), P[pre + "conv_out.weight"],
                       P[pre + "conv_out.bias"], padding=1)
        low = hold["low"]
        outs = [out] + (list(low) if isinstance(low, (list, tuple)) else [low])
        if plan.nesting:
            return outs, x
        return outs

    @property
    def nest_ratio(self):
        """NestedUNet.nest_ratio (nested_unet.py:134-145)."""
        r = []
        for (_, plan, _) in reversed(self.levels[:-1]):
            k = int(2 ** (len(plan.channels) - 1))
            r = [k * r[0]] + r if r else [k]
        return r
