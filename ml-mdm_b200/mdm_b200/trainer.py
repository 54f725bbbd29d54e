"""`train_batch` with the reference's signature and control flow (ml_mdm/trainer.py:13-97, fp32 branch), using the
fused clip + Adam + EMA + zero-grad sweep when the optimizer is `mdm_b200.optim.FusedAdam` and the reference's
separate calls otherwise."""
import numpy as np
import torch
import torch.nn as nn

from .optim import FusedAdam


def train_batch(model, sample, optimizer, scheduler, logger, args, grad_scaler=None, accumulate_gradient=False,
                num_grad_accumulations=1, ema_model=None, loss_factor=1.0):
    model.train()
    lr = scheduler.get_last_lr()[0]
    if getattr(args, "fp16", False):
        raise NotImplementedError("args.fp16 (bf16 autocast + GradScaler, trainer.py:29-61) is not built: the "
                                  "engine already multiplies fp16 operands with fp32 accumulation")
    losses, times, x_t, means, targets, weights = model.get_loss(sample)
    if weights is None:
        loss = losses.mean()
    else:
        loss = (losses * weights).sum() / weights.sum()
    loss_val = loss.item()
    if np.isnan(loss_val):  # trainer.py:69-74
        optimizer.zero_grad()
        optimizer.step()
        scheduler.step()
        return loss_val, losses, times, x_t, means, targets

    loss.backward()
    # (the reference divides `loss` by num_grad_accumulations only after backward, trainer.py:77-78: no effect)
    if not accumulate_gradient:
        vision = getattr(model.model, "module", model.model).vision_model
        if isinstance(optimizer, FusedAdam):
            optimizer.step(max_grad_norm=args.gradient_clip_norm, ema_model=ema_model)
        else:
            nn.utils.clip_grad_norm_(model.parameters(), args.gradient_clip_norm)
            optimizer.step()
            if ema_model is not None:
                ema_model.update(vision)

    if logger is not None and not accumulate_gradient:
        logger.add_scalar("train/Loss", loss_val)
        logger.add_scalar("lr", lr)

    if not accumulate_gradient:
        optimizer.zero_grad()
        scheduler.step()

    return loss_val, losses, times, x_t, means, targets
