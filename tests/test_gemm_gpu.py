"""tcgen05 GEMM engine parity on B200 (through the C ABI, mdm_gemm_raw) vs torch fp32 on the same
fp16 operands: tolerance 2e-5 relative (fp32 outputs; only the summation order differs), 1.5e-3 for
fp16-rounded outputs."""
import pytest

import gemm_cases as gc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,fn", gc.CASES, ids=[c[0] for c in gc.CASES])
def test_gemm_case(name, fn):
    errs = fn()
    for k, v in errs.items():
        assert v <= gc.TOL[k], (name, k, v)
