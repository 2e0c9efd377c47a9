"""GPU parity of the classical arc fit: the resampling /
scrunching kernels of csrc/normsspec.cu through the C ABI against the numpy
restatement of the reference loop and against outputs of the reference itself
(tests/golden/norm_sspec_64x96.npz, fit_arc_128x160.npz); then the whole chain
scale_dyn -> calc_sspec(lamsteps) -> norm_sspec -> fit_arc -> prep_thetatheta
without curvature bounds (the reference's default call path, dynspec.py:1458-1473)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from test_arcfit_cpu import numpy_norm_rows, _bare_dynspec   # noqa: E402


@pytest.fixture(scope="module")
def sb():
    import scintools_b200
    from scintools_b200 import _device
    _device.device()
    return scintools_b200


def test_norm_sspec_kernels_vs_reference(sb, golden_dir):
    """masks bit-exact, samples to fp32 rounding, on the reference's own sspec"""
    g = np.load(os.path.join(golden_dir, "norm_sspec_64x96.npz"))
    ds = _bare_dynspec(g)
    ds.sspec, ds.fdop, ds.tdel = g["sspec"], g["fdop"], g["tdel"]
    ds.freq = float(g["freq"])
    ds.norm_sspec(eta=float(g["eta"]), lamsteps=False, cutmid=int(g["cutmid"]),
                  startbin=int(g["startbin"]))
    assert np.array_equal(np.ma.getmaskarray(ds.normsspec), g["mask"])
    a, b = np.ma.filled(ds.normsspec, 0.0), np.nan_to_num(g["normsspec"])
    assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max()
    assert np.allclose(np.ma.filled(ds.normsspecavg, np.nan), g["normsspecavg"], rtol=1e-5,
                       atol=1e-5, equal_nan=True)
    assert np.allclose(np.ma.filled(ds.powerspectrum, np.nan), g["powerspectrum"], rtol=1e-5,
                       equal_nan=True)


def test_norm_rows_random_geometry(sb):
    """random curvatures / crops / NaN columns: device rows vs np.interp rows"""
    from scintools_b200 import arcfit
    rng = np.random.default_rng(5)
    for trial in range(6):
        nr, nc = int(rng.integers(5, 70)), 2 * int(rng.integers(8, 90))
        sspec = rng.normal(size=(nr, nc)) * 10
        if trial % 2:
            sspec[:, nc // 2 - 1:nc // 2 + 2] = np.nan
        fdop = (np.arange(nc) - nc // 2) * 0.37
        tdel = (1 + np.arange(nr)) * 0.21
        eta = float(10 ** rng.uniform(-2.5, 0.5))
        mx = float(rng.choice([1.0, 2.5, 5.0]))
        nq = 2 * int(rng.integers(5, 200))
        fdopnew = np.linspace(-mx, mx, nq)
        wf = lambda p: np.linspace(1.0, 2.0, nr)          # noqa: E731
        n1, p1, a1 = arcfit.norm_rows_device(sspec, fdop, tdel, eta, mx, fdopnew, wf)
        n0, p0, a0 = numpy_norm_rows(sspec.astype(np.float32).astype(np.float64), fdop, tdel, eta,
                                     mx, fdopnew, wf)
        n0 = np.atleast_2d(n0)
        assert np.array_equal(np.isnan(n1), np.isnan(n0)), trial
        assert np.allclose(n1, n0, rtol=1e-6, atol=1e-5, equal_nan=True)
        assert np.allclose(p1, p0, rtol=1e-6, equal_nan=True)
        assert np.allclose(a1, a0, rtol=1e-6, atol=1e-6, equal_nan=True)


def test_fit_arc_end_to_end(sb, golden_dir):
    """dyn -> scale_dyn -> calc_sspec(lamsteps) -> fit_arc on the device vs the
    reference's fit (fp32 secondary spectrum: the profile agrees to 1e-4 dB, the
    curvature well inside its own error bar)."""
    g = np.load(os.path.join(golden_dir, "fit_arc_128x160.npz"))
    ds = _bare_dynspec(g)
    ds.fit_arc(lamsteps=True)
    assert ds.betaeta == pytest.approx(float(g["betaeta"]), rel=2e-3)
    assert abs(ds.betaeta - float(g["betaeta"])) < 0.05 * float(g["betaetaerr2"])
    assert ds.betaetaerr == pytest.approx(float(g["betaetaerr"]), rel=0.05)
    assert ds.noise == pytest.approx(float(g["noise"]), rel=1e-3)
    assert ds.eta_array.shape == g["eta_array"].shape
    prof, ref = np.ma.filled(ds.norm_sspec_avg, np.nan), g["norm_sspec_avg"]
    assert np.nanmax(np.abs(prof - ref)) < 2e-3
    ds.fit_arc(lamsteps=True, asymm=True, nsmooth=7, low_power_diff=-2.0, high_power_diff=-1.0)
    assert ds.betaeta_left == pytest.approx(float(g["betaeta_left"]), rel=2e-3)
    assert ds.betaeta_right == pytest.approx(float(g["betaeta_right"]), rel=2e-3)


def test_prep_thetatheta_without_bounds(sb, golden_dir):
    """the reference's default call: no eta_min / eta_max -> Hough prior from fit_arc
    (dynspec.py:1458-1473); eta range = prior +- 2 max(err) clipped to the grid limits"""
    g = np.load(os.path.join(golden_dir, "fit_arc_128x160.npz"))
    ds = _bare_dynspec(g)
    ds.prep_thetatheta(cwf=64, cwt=80, nedge=64)
    c = 299792458.0
    eta_h = c * ds.betaeta / ds.fref ** 2 * 1e-6
    err_h = c * 2 * max(ds.betaetaerr, ds.betaetaerr2) / ds.fref ** 2 * 1e-6
    assert hasattr(ds, "betaeta") and ds.eta_min < eta_h < ds.eta_max
    assert ds.eta_max <= eta_h + err_h + 1e-12 and ds.eta_min >= eta_h - err_h - 1e-12
    assert ds.neta >= 2 and ds.edges.shape[0] == 64
