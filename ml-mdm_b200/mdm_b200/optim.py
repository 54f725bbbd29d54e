"""Fused post-backward sweep: global-norm clip + Adam/AdamW + EMA + zero-grad in one
pass over the engine's flat gradient arena (csrc/optim.cu, `mdm_grad_norm` + `mdm_adam_ema_sweep`).

Replaces, with the same arithmetic (torch/optim/adam.py::_single_tensor_adam, torch/nn/utils/clip_grad.py,
ml_mdm/models/model_ema.py:25-34), the reference's four separate full-parameter passes
(ml_mdm/trainer.py:78-93; optimizer construction clis/train_parallel.py:122-134: Adam or AdamW with
weight_decay=0, eps=1e-8). There is no CPU path: parameters and gradients must live on the B200.
"""
import ctypes as C

import torch

from . import _lib
from .parallel import flat_grads

CHUNK = 32768  # elements per CTA of the sweep


class OptChunk(C.Structure):  # mirrors mdm_opt_chunk (include/mdm_b200.h)
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("ema", C.c_void_p),
                ("n", C.c_int64)]


class AdamCfg(C.Structure):  # mirrors mdm_adam_cfg
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("weight_decay", C.c_double), ("ema_decay", C.c_double), ("grad_scale", C.c_float),
                ("max_norm", C.c_float), ("adamw", C.c_int32), ("step", C.c_int32), ("zero_grad", C.c_int32)]


GRAD_NORM_SCRATCH = 1184


def chunk_rows(p_ptr, g_ptr, m_ptr, v_ptr, ema_ptr, numel, chunk=CHUNK):
    """Rows (p, g, m, v, ema, n) of the chunk table for one tensor: byte addresses advance together."""
    rows = []
    for start in range(0, numel, chunk):
        n = min(chunk, numel - start)
        b = 4 * start
        rows.append((p_ptr + b, g_ptr + b, m_ptr + b, v_ptr + b, (ema_ptr + b) if ema_ptr else 0, n))
    return rows


class FusedAdam(torch.optim.Optimizer):
    """`torch.optim.Adam(vision_model.parameters(), lr, eps=1e-8)` / `AdamW(..., weight_decay=0)` as the reference
    builds them (train_parallel.py:122-134), stepping through the fused sweep. `param_groups[0]["lr"]` is read at
    every step, so torch LR schedulers work; `state_dict()` has torch's Adam layout (step / exp_avg / exp_avg_sq).

        opt = FusedAdam(vision_model, lr=args.lr, adamw=args.use_adamw)
        ...loss.backward()
        opt.step(max_grad_norm=args.gradient_clip_norm, ema_model=ema)   # clip + Adam + EMA + zero the arena
        opt.zero_grad()                                                  # drops the (already zeroed) .grad views
    """

    def __init__(self, vision_model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, adamw=False):
        if not hasattr(vision_model, "native"):
            raise TypeError("FusedAdam drives an mdm_b200 UNet / NestedUNet (it needs the engine's gradient arena)")
        self.vision_model = vision_model
        self.adamw = bool(adamw)
        self.steps = 0
        self.last_grad_norm = None  # device scalar: the global norm before clipping (clip_grad_norm_'s return value)
        self._sig = None
        self._table = self._m = self._v = self._scratch = self._norm = None
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(list(vision_model.parameters()), defaults)

    # ------------------------------------------------------------------ chunk table
    def _build(self, native, arena, ema_model):
        ema_params = None
        if ema_model is not None:
            ema_named = dict(ema_model.module.named_parameters())
            ema_params = [ema_named[k] for k in native.param_names]
            for e, p in zip(ema_params, native.params):
                if e.shape != p.shape or e.device != p.device or e.dtype != torch.float32 or not e.is_contiguous():
                    raise _lib.MdmError("EMA copy must be a contiguous fp32 clone of the model on the same device")
        sig = (arena.data_ptr(), tuple(p.data_ptr() for p in native.params),
               tuple(e.data_ptr() for e in ema_params) if ema_params else None)
        if sig == self._sig:
            return
        dev = arena.device
        if self._m is None or self._m.numel() != arena.numel():
            # Adam state shares the arena's layout; a re-sorted arena (optimize_arena_layout) keeps per-tensor state
            old = {id(p): (self.state[p]["exp_avg"].clone(), self.state[p]["exp_avg_sq"].clone())
                   for p in native.params if p in self.state and "exp_avg" in self.state[p]}
            self._m = torch.zeros_like(arena)
            self._v = torch.zeros_like(arena)
        else:
            old = {id(p): (self.state[p]["exp_avg"].clone(), self.state[p]["exp_avg_sq"].clone())
                   for p in native.params if p in self.state and "exp_avg" in self.state[p]}
            self._m.zero_()
            self._v.zero_()
        rows = []
        a, m0, v0 = arena.data_ptr(), self._m.data_ptr(), self._v.data_ptr()
        for i, (p, off) in enumerate(zip(native.params, native.offsets)):
            mv = self._m[off:off + p.numel()].view_as(p)
            vv = self._v[off:off + p.numel()].view_as(p)
            if id(p) in old:
                mv.copy_(old[id(p)][0])
                vv.copy_(old[id(p)][1])
            st = self.state[p]
            st["exp_avg"], st["exp_avg_sq"] = mv, vv
            st.setdefault("step", torch.tensor(float(self.steps)))
            if not p.requires_grad:
                continue
            rows += chunk_rows(p.data_ptr(), a + 4 * off, m0 + 4 * off, v0 + 4 * off,
                               ema_params[i].data_ptr() if ema_params else 0, p.numel())
        self._table = torch.tensor(rows, dtype=torch.int64).reshape(-1, 6).to(dev)
        if self._scratch is None:
            self._scratch = torch.zeros(GRAD_NORM_SCRATCH, device=dev, dtype=torch.float64)
            self._norm = torch.zeros(1, device=dev, dtype=torch.float32)
        self._sig = sig

    def load_state_dict(self, state_dict):
        """torch's Adam layout in, fused state out: the step counter continues from the loaded `step` (bias
        corrections 1 - beta^t stay consistent with the loaded moments) and the chunk table is rebuilt so the
        kernel reads the loaded exp_avg / exp_avg_sq rather than its previous arena-shaped copies."""
        super().load_state_dict(state_dict)
        steps = [float(st["step"]) for st in self.state.values() if "step" in st]
        self.steps = int(max(steps)) if steps else 0
        self._sig = None

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None, max_grad_norm=None, ema_model=None, grad_scale=1.0, zero_grad=True):
        if closure is not None:
            raise NotImplementedError("closures are not used by the reference trainer")
        vm = self.vision_model
        native = vm.native()
        if all(p.grad is None for p in native.params or []):
            return None  # nothing to do (the reference's NaN-loss path steps with no gradients)
        arena = flat_grads(vm)
        if arena is None:
            raise _lib.MdmError("FusedAdam needs every gradient in the engine's arena (gradients produced by "
                                "mdm_net_backward); found foreign .grad tensors")
        self._build(native, arena, ema_model)
        g = self.param_groups[0]
        self.steps += 1
        lib = _lib.lib()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        clip = float(max_grad_norm) if max_grad_norm is not None and max_grad_norm > 0 else 0.0
        _lib.check(lib.mdm_grad_norm(C.c_void_p(arena.data_ptr()), C.c_int64(arena.numel()), C.c_float(grad_scale),
                                     C.c_void_p(self._scratch.data_ptr()), C.c_int32(GRAD_NORM_SCRATCH),
                                     C.c_void_p(self._norm.data_ptr()), st), "mdm_grad_norm")
        self.last_grad_norm = self._norm
        cfg = AdamCfg()
        cfg.lr, (cfg.beta1, cfg.beta2), cfg.eps = g["lr"], g["betas"], g["eps"]
        cfg.weight_decay, cfg.adamw, cfg.step = g["weight_decay"], int(self.adamw), self.steps
        cfg.grad_scale, cfg.max_norm, cfg.zero_grad = grad_scale, clip, int(bool(zero_grad))
        cfg.ema_decay = 0.0
        if ema_model is not None:
            # ModelEma.update (model_ema.py:25-27): decay = (counter >= warmup_steps) * decay; counter += 1
            cfg.ema_decay = float(ema_model.counter >= ema_model.warmup_steps) * float(ema_model.decay)
            ema_model.counter += 1
        _lib.check(lib.mdm_adam_ema_sweep(C.c_void_p(self._table.data_ptr()), C.c_int32(self._table.shape[0]),
                                          C.byref(cfg), C.c_void_p(self._norm.data_ptr()), st), "mdm_adam_ema_sweep")
        for p in native.params:
            self.state[p]["step"] = torch.tensor(float(self.steps))
        native.versions = None            # fp32 masters changed in place: fp16 operand copies are rebuilt next forward
        if ema_model is not None:
            # the sweep wrote the EMA parameters through raw pointers too (their torch ._version did not move):
            # an engine already built for the EMA module must repack its fp16 operands before its next forward
            ema_native = getattr(ema_model.module, "_native", None)
            if ema_native is not None:
                ema_native.versions = None
        if zero_grad:
            native.arena_zeroed = True    # the next backward skips its 1.8 GB memset
        return None
