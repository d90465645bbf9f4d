"""Property tests of the index / bounds / gating logic: the C oracle against the pure-Python
transcription (tests/pyref.py) on random small shapes, radii and coordinates, including NaN, +-inf, huge and
exactly-integer coordinates.  Exact: zero pattern (gated taps), NaN pattern, in-place offset mutation; values within 2e-6 relative."""
import math

import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

import pyref

pytestmark = pytest.mark.filterwarnings("ignore::DeprecationWarning")

SPECIAL = [float("nan"), float("inf"), -float("inf"), 1e30, -1e30, 2147483648.0, -2147483649.0, -0.0, 0.0, -1.0,
           0.99999994, 1e-30]

coord = st.one_of(st.floats(-12.0, 20.0, allow_nan=False, width=32), st.sampled_from(SPECIAL),
                  st.integers(-3, 12).map(float))


@st.composite
def cases(draw):
    H1, W1 = draw(st.integers(1, 3)), draw(st.integers(1, 4))
    H2, W2 = draw(st.integers(1, 7)), draw(st.integers(1, 9))
    r = draw(st.integers(0, 3))
    E = draw(st.integers(1, 2))
    n = E * 2 * H1 * W1
    cs = draw(st.lists(coord, min_size=n, max_size=n))
    rd = 2 * r + 1
    m = E * H1 * W1 * rd * rd * 2
    offs = draw(st.lists(st.one_of(st.floats(-5.0, 5.0, allow_nan=False, width=32), st.sampled_from([0.0, 4.0, -4.0, 1e9])),
                         min_size=m, max_size=m))
    seed = draw(st.integers(0, 2 ** 16))
    return E, H1, W1, H2, W2, r, cs, offs, seed


def _eq(a, b, tol=2e-6):
    """Same NaN pattern, same zero (gated-tap) pattern, values within the fp32 operation-order tolerance: pyref is an
    independent plain-arithmetic restatement, the oracle follows the reference's FMA order."""
    if not (torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(a == 0, b == 0)):
        return False
    a, b = torch.nan_to_num(a, posinf=0, neginf=0), torch.nan_to_num(b, posinf=0, neginf=0)
    return bool(((a - b).abs() <= tol * (1 + b.abs())).all())


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(cases())
def test_lookup_index_logic_matches_python_transcription(oracle, case):
    E, H1, W1, H2, W2, r, cs, offs, seed = case
    g = torch.Generator().manual_seed(seed)
    vol = torch.randn(E, H1, W1, H2, W2, generator=g)
    coords = torch.tensor(cs, dtype=torch.float32).view(E, 2, H1, W1)
    rd = 2 * r + 1
    offset = torch.tensor(offs, dtype=torch.float32).view(E, H1, W1, rd, rd, 2)
    got, = oracle.corr_index_forward(vol, coords, r)
    want = torch.from_numpy(pyref.lookup_forward(vol.numpy(), coords.numpy(), None, r, deform=False))
    assert _eq(got, want)
    o1, o2 = offset.clone(), offset.clone().numpy()
    got, = oracle.defCorr_index_forward(vol, coords, o1, r)
    want = torch.from_numpy(pyref.lookup_forward(vol.numpy(), coords.numpy(), o2, r, deform=True))
    assert _eq(got, want)
    assert _eq(o1, torch.from_numpy(o2)), "in-place centre-tap zeroing (Q5)"
    # backward: the same taps must be gated (zero gradient pattern), values within fp32 accumulation order
    grad = torch.randn(E, rd, rd, H1, W1, generator=g)
    gv, go = oracle.defCorr_index_backward(vol, coords, offset.clone(), grad, r)
    wv, wo = pyref.lookup_backward(vol.numpy(), coords.numpy(), offset.clone().numpy(), grad.numpy(), r, deform=True)
    wv, wo = torch.from_numpy(wv), torch.from_numpy(wo)
    assert torch.equal(torch.isnan(gv), torch.isnan(wv)) and torch.equal(torch.isnan(go), torch.isnan(wo))
    assert torch.allclose(torch.nan_to_num(gv), torch.nan_to_num(wv), atol=1e-5, rtol=1e-5)
    assert torch.allclose(torch.nan_to_num(go), torch.nan_to_num(wo), atol=1e-4, rtol=1e-4)
