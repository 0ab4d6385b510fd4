"""Coefficient-ROW (modulated) filter forms: host coefficient design and the two CPU checkers.  CPU only.

Reference: Lopass::makeCoeffsVec / operator()(vx, omega, k) (MLDSPFilters.h:97-115,136-152),
interpolateCoeffsLinear (:32-44), LoShelf / HiShelf::vcoeffs and operator()(vx, vc) (:283-286,304-319,364-367,385-400).
"""
import numpy as np
import pytest

from madronalib_b200 import api, workloads as wl
from tests.common import assert_same_bits, assert_state_equal


def _sweeps(T, seed=3):
    rng = np.random.default_rng(seed)
    n = np.arange(T * 64, dtype=np.float32).reshape(T, 64)
    omega = (0.02 + 0.3 * (0.5 + 0.5 * np.sin(n * 0.013))).astype(np.float32)   # crosses the 0.5 clamp? no: <= 0.32
    omega[0, :4] = [0.6, 0.5, np.nan, -0.1]                                         # min(omega, 0.5) operand order
    k = (0.005 + 1.5 * rng.random((T, 64))).astype(np.float32)                     # some below the 0.01 clamp
    k[0, 4:6] = [np.nan, 0.0]
    return omega, k


def test_host_makecoeffsvec_equals_reference(ref, port):
    omega, k = _sweeps(5)
    for t in range(5):
        want = ref.coeffs_lopass_vec(omega[t], k[t])
        assert_same_bits(api.coeffs_lopass_vec(omega[t], k[t]), want, "mlb_coeffs_lopass_vec")
        assert_same_bits(port.coeffs_lopass_vec(omega[t], k[t]), want, "mlport_coeffs_lopass_vec")


@pytest.mark.parametrize("kind", ["loshelf", "hishelf"])
def test_host_vcoeffs_equals_reference(ref, kind):
    p0, p1 = (0.05, 0.7, api.db_to_gain(-6.0)), (0.21, 1.3, api.db_to_gain(9.0))
    want = ref.shelf_vcoeffs(kind, p0, p1)
    got = api.interpolate_coeffs_linear(api.coeffs(kind, *p0), api.coeffs(kind, *p1))
    assert_same_bits(got, want, kind + " vcoeffs")


def test_lopass_v_node_equals_reference_operator(ref, port):
    """LOPASS_V fed with mlb_coeffs_lopass_vec rows == the reference's own operator()(vx, omega, k)."""
    T = 6
    omega, k = _sweeps(T)
    x = (np.random.default_rng(1).standard_normal((T, 64)) * 0.5).astype(np.float32)
    want = ref.lopass_mod(x, omega, k)
    w = wl.swept_filter_case("lopass_v", 1, T, x=x[:, None], omega=omega[:, None], k=k[:, None])
    for O in (ref, port):
        out, _, _ = O.run(w.spec, 1, T, w.inputs(T), w.state, w.coef)
        assert_same_bits(out[:, 0, 0], want, "LOPASS_V vs Lopass::operator()(vx, omega, k)")
    w2 = wl.swept_filter_case("lopass_mod", 1, T, x=x[:, None], omega=omega[:, None], k=k[:, None])
    for O in (ref, port):
        out, _, _ = O.run(w2.spec, 1, T, w2.inputs(T), w2.state, w2.coef)
        assert_same_bits(out[:, 0, 0], want, "LOPASS_MOD (checker) vs the reference operator")


@pytest.mark.parametrize("name", wl.SWEPT_CASES)
def test_port_equals_reference_on_swept_filters(ref, port, name):
    V, T = 37, 7
    w = wl.swept_filter_case(name, V, T)
    inp = w.inputs(T)
    ro, _, rs = ref.run(w.spec, V, T, inp, w.state, w.coef, splits=(3, 4))
    po, _, ps = port.run(w.spec, V, T, inp, w.state, w.coef)
    assert_same_bits(po, ro, name)
    assert_state_equal(ps, rs, name)
    assert np.isfinite(ro).all() and np.abs(ro).max() > 1e-3
