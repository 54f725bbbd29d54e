"""B200: forward AND backward parity of the three shipped architectures at full width (cc12m_64x64: 461 M parameters,
cc12m_256x256: 2-level nest, cc12m_1024x1024: 3-level nest with per-sample std input normalisation on its middle
level, 1- and 3-channel GroupNorm groups at C = 32 / 96, K = 13824 weight gradients), explicit micro-conditioning
values on both sides of the clamp, ragged caption masks.

Arbiter: the oracle in float64 on the same GPU. Calibration: the oracle in float32 with TF32 enabled -- the arithmetic
the reference itself trains with (clis/train_parallel.py:18-19). The north star's "1e-3 relative" is what that path
scores on the OUTPUTS (0.6e-3 .. 1.5e-3 measured, profiles/r02_parity_fullwidth.txt) and it scores ~3e-3 (median) to
~7e-3 (max) on parameter gradients; the bounds below are stated against those measured reference errors, tensor by
tensor, not as free constants:
  outputs     ours <= max(1e-3, 1.5 x reference-TF32 error of the same output)   (measured ratios 1.06 .. 1.42; both
              sides move by ~10 % run to run with the order of their fp32 atomics)
  gradients   median and 90th percentile over all parameters within 1.25 x of the reference-TF32's, and every single
              tensor within 3 x max(its reference-TF32 error, the median reference-TF32 error)
Gradients that are mathematically zero (a conv bias in front of a GroupNorm whose groups are single channels, C = 32)
are round-off on every implementation (reference-TF32 relative error > 0.5): their magnitude is checked against the
magnitude the reference-TF32 path leaves in the same tensor (or 1e-3 of the typical gradient, whichever is larger)."""
import pytest

import fullwidth_cases as fc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,batch", [("cc12m_64x64", 2), ("cc12m_256x256", 2), ("cc12m_1024x1024", 1)])
def test_forward_backward_full_width_calibrated_against_reference_tf32(name, batch):
    rep = fc.run_case(name, B=batch, S=8, micro=True)
    assert not rep["missing"], rep["missing"]
    for i, (ours, tf32) in enumerate(rep["out"]):
        assert ours <= max(1e-3, 1.5 * tf32), (name, i, ours, tf32)
    defined = {k: v for k, v in rep["grads"].items() if v[1] <= 0.5}
    undefined = {k: v for k, v in rep["grads"].items() if v[1] > 0.5}
    assert len(undefined) <= 0.05 * len(rep["grads"]), sorted(undefined)[:10]
    for k in undefined:  # zero by construction: round-off of the size the reference's own TF32 path leaves there
        lim = 3.0 * max(rep["grad_absmax_tf32"][k], 1e-3 * rep["grad_typical"])
        assert rep["grad_absmax"][k] <= lim, (k, rep["grad_absmax"][k], rep["grad_absmax_tf32"][k], rep["grad_typical"])
    o = sorted(v[0] for v in defined.values())
    r = sorted(v[1] for v in defined.values())
    n = len(o)
    assert o[n // 2] <= 1.25 * r[n // 2], (o[n // 2], r[n // 2])
    assert o[int(0.9 * n)] <= 1.25 * r[int(0.9 * n)], (o[int(0.9 * n)], r[int(0.9 * n)])
    med = r[n // 2]
    bad = {k: v for k, v in defined.items() if v[0] > 3.0 * max(v[1], med)}
    assert not bad, bad
