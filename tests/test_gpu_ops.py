"""-m gpu: every C-ABI operator against a plain torch fp32 reference of the same op
(cases shared with tools/gpu_probe_ops.py)."""
import pytest

from tools import gpu_probe_ops as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(P.CASES))
def test_operator(name):
    r = P.CASES[name]()
    # fp16 storage: one output rounding (2^-11 relative) + fp32 accumulation noise
    assert r["max_abs"] < 2.5e-3 * max(1.0, r.get("ref_max", 1.0)), r
    if "rel_fro" in r:
        assert r["rel_fro"] < 1e-3, r
