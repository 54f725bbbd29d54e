"""Bridge between the parameter-container modules and the native engine (mdm_net_* in
include/mdm_b200.h).  torch is used for device memory, streams and autograd bookkeeping only."""
import ctypes as C
import os

import torch

from .. import _lib

MAX_RES, MAX_LEVELS = 8, 4


class LevelCfg(C.Structure):
    _fields_ = [
        ("num_res", C.c_int32),
        ("channels", C.c_int32 * MAX_RES),
        ("num_resnets", C.c_int32 * MAX_RES),
        ("num_attn", C.c_int32 * MAX_RES),
        ("cond_level", C.c_int32 * MAX_RES),
        ("temporal_dim", C.c_int32),
        ("groups", C.c_int32),
        ("use_attention_ffn", C.c_int32),
        ("skip_mid_blocks", C.c_int32),
        ("nesting", C.c_int32),
        ("skip_normalization", C.c_int32),
        ("has_micro_scale", C.c_int32),
        ("micro_scale_default", C.c_float),
    ]


class NetCfg(C.Structure):
    _fields_ = [
        ("num_levels", C.c_int32),
        ("levels", LevelCfg * MAX_LEVELS),
        ("in_channels", C.c_int32),
        ("out_channels", C.c_int32),
        ("lm_dim", C.c_int32),
        ("cond_dim", C.c_int32),
        ("has_lm_proj", C.c_int32),
        ("has_cond_emb", C.c_int32),
        ("masked_cross_attention", C.c_int32),
        ("num_heads", C.c_int32),
    ]


class NetIO(C.Structure):
    _fields_ = [
        ("batch", C.c_int32),
        ("tokens", C.c_int32),
        ("res", C.c_int32 * MAX_LEVELS),
        ("x_t", C.c_void_p * MAX_LEVELS),
        ("times", C.c_void_p),
        ("lm", C.c_void_p),
        ("lm_mask", C.c_void_p),
        ("micro_scale", C.c_void_p),
        ("out", C.c_void_p * MAX_LEVELS),
        ("save_for_backward", C.c_int32),
        ("level_batch", C.c_int32 * MAX_LEVELS),
        ("apply_lm_mask", C.c_int32),
    ]


class NetGradIO(C.Structure):
    _fields_ = [("dout", C.c_void_p * MAX_LEVELS)]


def _ints(v, n=None):
    if v is None:
        return []
    if isinstance(v, str):
        v = [int(x) for x in v.split(",")] if v else []
    v = [int(x) for x in v]
    if n is not None and len(v) == 1:
        v = v * n
    return v


def build_net_cfg(module) -> NetCfg:
    """module: UNet / NestedUNet container. Translates its config objects into the C struct."""
    cfgs = module._level_configs()
    mods = module._levels()
    nc = NetCfg()
    nc.num_levels = len(cfgs)
    for li, (cfg, m) in enumerate(zip(cfgs, mods)):
        lc = nc.levels[li]
        ch = _ints(cfg.resolution_channels)
        L = len(ch)
        assert L <= MAX_RES
        nres = _ints(cfg.num_resnets_per_resolution, L)
        nattn = _ints(cfg.num_attention_layers, L)
        levels = _ints(cfg.attention_levels)
        lc.num_res = L
        for i in range(L):
            lc.channels[i] = ch[i]
            lc.num_resnets[i] = nres[i]
            lc.num_attn[i] = nattn[i] if i in levels else 0
            lc.cond_level[i] = 1 if i in levels else 0
        lc.temporal_dim = m.temporal_dim
        lc.groups = cfg.resnet_config.num_groups_norm
        lc.use_attention_ffn = int(bool(cfg.resnet_config.use_attention_ffn))
        lc.skip_mid_blocks = int(bool(cfg.skip_mid_blocks))
        lc.nesting = int(bool(cfg.nesting))
        lc.skip_normalization = int(bool(getattr(cfg, "skip_normalization", True)))
        lc.has_micro_scale = int(m.conditions is not None)
        lc.micro_scale_default = float(m.conditions["scale"]) if m.conditions is not None else 0.0
    inner = mods[-1]
    icfg = cfgs[-1]
    nc.in_channels = module.input_channels
    nc.out_channels = module.output_channels
    nc.lm_dim = max(int(inner.input_conditioning_feature_dim), 0)
    nc.cond_dim = max(int(icfg.conditioning_feature_dim), 0)
    nc.has_lm_proj = int(hasattr(inner, "lm_proj"))
    nc.has_cond_emb = int(inner.cond_emb is not None)
    nc.masked_cross_attention = int(icfg.masked_cross_attention)
    nc.num_heads = 8
    return nc


class _DenoiseFn(torch.autograd.Function):
    """One autograd node for the whole denoiser: forward = mdm_net_forward, backward = mdm_net_backward.
    Parameters are passed so autograd routes their gradients (DDP hooks, accumulation, clipping work
    on ordinary .grad tensors)."""

    @staticmethod
    def forward(ctx, native, nlev, need_grad, times, lm, mask, micro, *rest):
        xs = rest[:nlev]
        outs = native._forward(list(xs), times, lm, mask, micro, save=need_grad, apply_lm_mask=native.apply_lm_mask)
        ctx.native = native
        ctx.nlev = nlev
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        native = ctx.native
        grads = native._backward(list(gouts))
        return (None,) * 7 + (None,) * ctx.nlev + tuple(grads)


ARENA_ALIGN = 64  # elements: every gradient view starts on a 256-byte boundary (vector loads in the optimiser sweep)


def _pad(n):
    return (n + ARENA_ALIGN - 1) // ARENA_ALIGN * ARENA_ALIGN


def arena_views(arena, params, offsets):
    """Fresh per-parameter views of a flat gradient arena for autograd. They must not be referenced anywhere
    else: AccumulateGrad adopts an incoming gradient as `.grad` only when it holds the sole reference and clones
    it otherwise, and the flat-arena paths (one NCCL all-reduce, GradientOverlap, FusedAdam) rely on `.grad`
    aliasing the arena."""
    return [arena[off:off + p.numel()].view_as(p) for p, off in zip(params, offsets)]


class NativeNet:
    def __init__(self, module):
        self.module = module
        self.lib = _lib.lib()
        self.cfg = build_net_cfg(module)
        self.handle = C.c_void_p()
        _lib.check(self.lib.mdm_net_create(C.byref(self.cfg), C.byref(self.handle)), "mdm_net_create")
        self.lib.mdm_net_workspace_bytes.restype = C.c_uint64
        self.lib.mdm_net_workspace_high_water.restype = C.c_uint64
        self.lib.mdm_net_debug_fetch.restype = C.c_int64
        # parameter table of the engine
        self.names, self.shapes = [], {}
        n = self.lib.mdm_net_num_params(self.handle)
        for i in range(n):
            name = C.c_char_p()
            nd = C.c_int32()
            shape = (C.c_int64 * 4)()
            _lib.check(self.lib.mdm_net_param_info(self.handle, i, C.byref(name), C.byref(nd), shape), "param_info")
            nm = name.value.decode()
            self.names.append(nm)
            self.shapes[nm] = tuple(shape[j] for j in range(nd.value))
        self._check_tree()
        self.params = None
        self.sig = None
        self.grad_arena = None
        self.active_arena = None
        self.order = None
        self.arena_zeroed = False
        self.offsets = None
        self._ready_cb = None
        self._keep = None
        self.apply_lm_mask = False
        # CUDA-graph replay of forward / backward (mdm_net_set_graph_mode): on unless MDM_NO_GRAPH is set; switched
        # off for this net by gradient accumulation (a fresh arena per backward would re-record every step). With a
        # gradient-ready callback installed the backward is recorded as one graph per reported range.
        self.graphs = os.environ.get("MDM_NO_GRAPH") is None
        self.lib.mdm_net_set_graph_mode(self.handle, int(self.graphs))

    def set_graph_mode(self, on):
        self.graphs = bool(on)
        _lib.check(self.lib.mdm_net_set_graph_mode(self.handle, int(self.graphs)), "mdm_net_set_graph_mode")

    def __del__(self):
        try:
            if self.handle:
                self.lib.mdm_net_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _tensors(self):
        d = dict(self.module.named_parameters())
        for k, v in self.module.named_buffers():
            if k.endswith("t_emb"):
                d[k] = v
        return d

    def _check_tree(self):
        t = self._tensors()
        mine, theirs = set(t), set(self.names)
        if mine != theirs:
            raise _lib.MdmError(f"parameter tree mismatch: only in module {sorted(mine - theirs)[:5]}, "
                                f"only in engine {sorted(theirs - mine)[:5]}")
        for k in self.names:
            if tuple(t[k].shape) != self.shapes[k]:
                raise _lib.MdmError(f"shape mismatch for {k}: module {tuple(t[k].shape)} engine {self.shapes[k]}")

    # ---------------------------------------------------------------- binding
    def _bind(self):
        t = self._tensors()
        plist = [(k, t[k]) for k in self.names]
        for k, p in plist:
            if not p.is_cuda:
                raise _lib.MdmError("mdm_b200 runs on a CUDA (sm_100a) device only; move the model with .to('cuda'). "
                                    "There is no CPU path.")
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.MdmError(f"parameter {k} must be contiguous fp32")
        sig = tuple(p.data_ptr() for _, p in plist) + tuple(p.requires_grad for _, p in plist)
        if sig == self.sig:
            return
        dev = plist[0][1].device
        total = sum(_pad(p.numel()) for k, p in plist if not k.endswith("t_emb"))
        self.grad_arena = torch.zeros(total, device=dev, dtype=torch.float32)
        self.arena_zeroed = True
        off = 0
        self.params, self.param_names, self.offsets = [], [], []
        # Arena layout: registration order at first; optimize_arena_layout() re-sorts it by how late each
        # gradient becomes final so that mdm_net_set_grad_ready can report it from the top down.
        if self.order is not None:
            by_name = dict(plist)
            plist = [(k, by_name[k]) for k in self.order]
        for k, p in plist:
            if k.endswith("t_emb"):
                _lib.check(self.lib.mdm_net_bind_param(self.handle, k.encode(), C.c_void_p(p.data_ptr()), None), "bind")
                continue
            self.param_names.append(k)
            self.offsets.append(off)
            g = self.grad_arena[off:off + p.numel()].view_as(p)
            off += _pad(p.numel())
            # frozen parameters (requires_grad=False, e.g. freeze_inner_unet, nested_unet.py:147-150) get no
            # gradient pointer: the engine skips their weight gradients, their arena slot stays zero and the
            # global norm / optimiser sweep never see them (clip_grad_norm_ ignores them in the reference too)
            _lib.check(self.lib.mdm_net_bind_param(self.handle, k.encode(), C.c_void_p(p.data_ptr()),
                                                   C.c_void_p(g.data_ptr()) if p.requires_grad else None), "bind")
            self.params.append(p)
        self.sig = sig
        self.versions = None

    def _sync_weights(self):
        v = sum(p._version for p in self.params)
        if v != self.versions:
            self.lib.mdm_net_weights_changed(self.handle)
            self.versions = v

    # ---------------------------------------------------------------- public
    def run(self, xs, times, lm, mask, micros, apply_lm_mask=False):
        self._bind()
        self.apply_lm_mask = bool(apply_lm_mask)
        micro = None
        if micros:
            micro = micros.get("scale", None)
        # (Function.forward runs with grad mode off, so the decision is taken here)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.params)
        return _DenoiseFn.apply(self, len(xs), need_grad, times, lm, mask, micro, *xs, *self.params)

    def _forward(self, xs, times, lm, mask, micro, save, apply_lm_mask=False):
        self._sync_weights()
        if save and self.graphs and self.grad_arena is not None:
            lo, hi = self.grad_arena.data_ptr(), self.grad_arena.data_ptr() + self.grad_arena.numel() * 4
            if any(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in self.params):
                self.set_graph_mode(False)  # gradient accumulation: this backward needs a fresh arena (see _backward)
        io = NetIO()
        B = xs[-1].shape[0]  # the innermost level always runs the whole batch (nested_unet.py:180,200-204)
        io.batch = B
        keep = []

        def f32(t):
            t = t.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            keep.append(t)
            return t

        outs = []
        for i, x in enumerate(xs):
            if not x.is_cuda:
                raise _lib.MdmError("inputs must be CUDA tensors")
            x = f32(x)
            if x.shape[2] != x.shape[3]:
                raise _lib.MdmError("square images only")
            if not (1 <= x.shape[0] <= B) or (i > 0 and x.shape[0] < xs[i - 1].shape[0]):
                raise _lib.MdmError("mixed-resolution batches: each level runs a leading part of the batch and inner "
                                    f"levels at least as many samples as outer ones; got {[t.shape[0] for t in xs]}")
            io.level_batch[i] = x.shape[0]
            io.res[i] = x.shape[2]
            io.x_t[i] = x.data_ptr()
            o = torch.empty_like(x)
            outs.append(o)
            io.out[i] = o.data_ptr()
        t64 = times.detach().to(torch.int64).contiguous()
        keep.append(t64)
        io.times = t64.data_ptr()
        if lm is not None:
            lm = f32(lm)
            io.tokens = lm.shape[1]
            io.lm = lm.data_ptr()
            if mask is not None:
                mask = f32(mask)
                io.lm_mask = mask.data_ptr()
        if micro is not None:
            micro = f32(micro)
            io.micro_scale = micro.data_ptr()
        io.save_for_backward = int(save)
        io.apply_lm_mask = int(bool(apply_lm_mask))
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self.lib.mdm_net_forward(self.handle, C.byref(io), C.c_void_p(st)), "mdm_net_forward")
        self._keep = keep if save else None  # inputs must outlive the tape
        return outs

    def _backward(self, gouts):
        gio = NetGradIO()
        keep = []
        for i, g in enumerate(gouts):
            if g is None:
                continue
            g = g.detach()
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = g.float().contiguous()
            keep.append(g)
            gio.dout[i] = g.data_ptr()
        # fresh arena when existing .grad tensors alias the persistent one (gradient accumulation)
        arena = self.grad_arena
        lo, hi = arena.data_ptr(), arena.data_ptr() + arena.numel() * 4
        aliased = any(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in self.params)
        if aliased:
            arena = torch.zeros_like(self.grad_arena)
            views = self._views(arena)
            for k, p, g in zip(self.param_names, self.params, views):
                _lib.check(self.lib.mdm_net_bind_param(self.handle, k.encode(), C.c_void_p(p.data_ptr()),
                                                       C.c_void_p(g.data_ptr()) if p.requires_grad else None), "bind")
            self.sig = None  # rebinding to the persistent arena happens at the next forward
        else:
            if not self.arena_zeroed:  # the fused optimiser sweep (optim.FusedAdam) leaves it zeroed
                arena.zero_()
            self.arena_zeroed = False
            views = self._views(arena)
        st = torch.cuda.current_stream().cuda_stream
        self.active_arena = arena  # what a gradient-ready callback (parallel.GradientOverlap) indexes into
        _lib.check(self.lib.mdm_net_backward(self.handle, C.byref(gio), C.c_void_p(st)), "mdm_net_backward")
        self._keep = None
        return [g if p.requires_grad else None for p, g in zip(self.params, views)]

    def _views(self, arena):
        return arena_views(arena, self.params, self.offsets)

    def optimize_arena_layout(self):
        """Re-sort the gradient arena by the order gradients become final in backward (learned by the
        engine during a previous backward, mdm_net_grad_order), latest at the lowest address. Returns
        False when nothing has been learned yet. Takes effect at the next forward; gradients already
        handed out keep their (old) storage."""
        n = len(self.names)
        rank = (C.c_int32 * n)()
        if self.lib.mdm_net_grad_order(self.handle, rank, C.c_int32(n)) != 0:
            return False
        idx = sorted(range(n), key=lambda i: (rank[i], i))
        order = [self.names[i] for i in idx]
        if order != self.order:
            self.order = order
            self.sig = None  # rebind at the next forward
        return True

    def set_grad_ready(self, fn, min_bytes=0):
        """Install (or with fn=None remove) the engine's gradient-ready notification: fn(lo_ptr, hi_ptr) is
        called from inside mdm_net_backward whenever the gradient bytes at addresses [lo, hi) are final."""
        if fn is None:
            self._ready_cb = None
            _lib.check(self.lib.mdm_net_set_grad_ready(self.handle, None, None, C.c_uint64(0)), "set_grad_ready")
            return
        cb = _lib.GRAD_READY_FN(lambda user, lo, hi: fn(int(lo or 0), int(hi or 0)))
        self._ready_cb = cb  # keep the trampoline alive as long as the engine may call it
        _lib.check(self.lib.mdm_net_set_grad_ready(self.handle, cb, None, C.c_uint64(int(min_bytes))), "set_grad_ready")

    def workspace_bytes(self):
        return int(self.lib.mdm_net_workspace_bytes(self.handle)), int(self.lib.mdm_net_workspace_high_water(self.handle))

    def debug_fetch(self, name, shape):
        out = torch.empty(shape, device="cuda", dtype=torch.float32)
        st = torch.cuda.current_stream().cuda_stream
        n = self.lib.mdm_net_debug_fetch(self.handle, name.encode(), C.c_void_p(out.data_ptr()), C.c_int64(out.numel()),
                                         C.c_void_p(st))
        if n < 0:
            raise _lib.MdmError(f"no debug tensor {name} ({n})")
        return out
