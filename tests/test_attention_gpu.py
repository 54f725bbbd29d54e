"""Fused attention operator (tcgen05 flash-style forward + backward) vs an fp64 torch restatement of
SelfAttention.attention (reference models/unet.py:276-307) on the same fp16 operands. Tolerance 4e-3
(max|delta|/max|ref|): P and dS tiles and all outputs are fp16; measured ~4e-4."""
import pytest

import attn_cases as ac

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,fn", ac.CASES, ids=[c[0] for c in ac.CASES])
def test_attention_case(name, fn):
    errs = fn()
    for k, v in errs.items():
        assert v <= ac.TOL, (name, k, v)
