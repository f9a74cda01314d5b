"""The C restatement (oracle/jxl_oracle.c) against the committed golden vectors that
tests/golden/make_golden.py produced from the unmodified reference build."""
import numpy as np
import pytest

from libjxl_b200 import abi
from tests import support
from tests.golden.make_golden import transform_inputs

pytestmark = pytest.mark.usefixtures("built")


@pytest.fixture(scope="module")
def golden_transforms():
    return np.load(support.GOLDEN / "transforms.npz")


@pytest.mark.parametrize("strategy", range(27))
def test_transform_to_pixels_bit_exact(strategy, golden_transforms):
    from oracle import cpu
    coeffs, _ = transform_inputs(strategy)
    got = cpu.transform_to_pixels(strategy, coeffs)
    want = golden_transforms[f"px_{strategy}"]
    assert got.shape == want.shape
    assert np.array_equal(got, want), (abi.STRATEGY_NAMES[strategy], float(np.abs(got - want).max()))


@pytest.mark.parametrize("strategy", range(27))
def test_llf_from_dc_bit_exact(strategy, golden_transforms):
    from oracle import cpu
    _, dc = transform_inputs(strategy)
    n = 64 * abi.COVERED_X[strategy] * abi.COVERED_Y[strategy]
    got = cpu.llf_from_dc(strategy, dc, np.zeros(n, np.float32))
    want = np.zeros(n, np.float32)
    want[golden_transforms[f"llf_idx_{strategy}"]] = golden_transforms[f"llf_val_{strategy}"]
    assert np.array_equal(got, want)


def test_idct_dc_only_is_flat():
    """ac_strategy_test.cc:96-154 property: a DC-only block decodes to a constant."""
    from oracle import cpu
    for s in (0, 4, 5, 6, 7, 18, 21, 24):
        n = 64 * abi.COVERED_X[s] * abi.COVERED_Y[s]
        co = np.zeros(n, np.float32)
        co[0] = 1.25
        px = cpu.transform_to_pixels(s, co)
        assert np.allclose(px, 1.25, atol=1e-6)


def test_dct_idct_roundtrip():
    """dct_test.cc:217-249: ComputeScaledDCT then ComputeScaledIDCT is the identity."""
    from oracle import cpu
    rng = np.random.default_rng(5)
    for s in (0, 4, 5, 6, 7, 8, 9, 10, 11, 18, 19, 20):
        r, c = abi.COVERED_Y[s] * 8, abi.COVERED_X[s] * 8
        px = rng.normal(0, 1, (r, c)).astype(np.float32)
        co = cpu.scaled_dct(px)
        back = cpu.transform_to_pixels(s, co)
        assert np.abs(back - px).max() < 2e-5 * max(r, c)


def test_adjust_quant_bias():
    """quantizer-inl.h:35-67: 0 -> 0, +-1 -> +-biases[c], else q - biases[3]/q."""
    from oracle import cpu
    b = np.array([0.9453, 0.9299, 0.95, 0.145], np.float32)
    assert cpu.adjust_quant_bias(0, 0, b) == 0.0
    assert cpu.adjust_quant_bias(1, 1, b) == b[1]
    assert cpu.adjust_quant_bias(2, -1, b) == -b[2]
    for q in (2, -3, 17, -1000, 32767):
        want = np.float32(q) - np.float32(b[3] / np.float32(q))
        assert abs(cpu.adjust_quant_bias(0, q, b, 0) - want) <= 1e-6 * abs(q)
        # the rcpss flavour (reference's ApproximateReciprocal) stays within 12-bit accuracy
        assert abs(cpu.adjust_quant_bias(0, q, b, 1) - want) <= 0.145 * 4e-4 / abs(q) + 1e-6 * abs(q)


@pytest.mark.parametrize("tap", list(support.TAP_MASKS))
def test_frame_taps_against_reference(tap):
    """Whole small frame, stage by stage, against the reference's own DecodeGroupForRoundtrip +
    Gaborish/EPF/XYB stages (strict build). Exact-reciprocal mode: the only difference is the
    reference's 12-bit rcpps in AdjustQuantBias => tolerance, stated in absolute pixel units
    (XYB ~ [-1,1], linear RGB ~ [0,1])."""
    from oracle import cpu
    desc, coeffs, g = support.golden_desc(out_format=abi.OUT_PLANAR_F32)
    desc.stage_mask = abi.STAGE_EXPLICIT | support.TAP_MASKS[tap]
    got = cpu.render_frame(desc, coeffs, rcp_mode=0)
    want = g.taps[tap]
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 2e-5


def test_frame_full_decode_default_build():
    """Against what the public API of the default-flag reference build decodes (djxl
    --color_space=RGB_D65_SRG_Rel_Lin equivalent); conformance-style statistics."""
    from oracle import cpu
    desc, coeffs, g = support.golden_desc()
    got = cpu.render_frame(desc, coeffs, rcp_mode=0)
    want = g.decoded_default
    peak = float(np.abs(got - want).max())
    rmse = float(np.sqrt(np.mean((got - want) ** 2)))
    assert peak <= 5e-5 and rmse <= 5e-6, (peak, rmse)


def test_sigma_against_reference():
    from oracle import cpu
    desc, _, g = support.golden_desc()
    sg = cpu.compute_sigma(desc)[2:-2, 2:-2]
    assert np.array_equal(sg, g.sigma_interior)
