"""GPU: randomised geometry / shape sweeps.  Index and mask arrays must be
bit-exact against the oracle for every draw (including gathers that wrap to
negative fd, crops, odd grids and non-symmetric edges); FFT paths must hold
1e-5 for arbitrary (odd, prime, tiny) shapes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import dynspec_oracle as DO   # noqa: E402
from oracle import thth_oracle as TO      # noqa: E402


@pytest.fixture(scope="module")
def sb():
    import scintools_b200
    from scintools_b200 import _device
    _device.device()
    return scintools_b200


def maxrel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def test_thth_indices_random_geometry(sb):
    thth = sb.ththmod
    rng = np.random.default_rng(2024)
    checked = 0
    for trial in range(24):
        ntau = int(rng.choice([16, 50, 64, 131, 256]))
        nfd = int(rng.choice([16, 37, 64, 200, 512]))
        dt = float(rng.uniform(5, 60))
        df = float(rng.uniform(0.01, 0.5))
        t = dt * np.arange(nfd)
        f = 1400 + df * np.arange(ntau)
        fd = TO.fft_axis(t, "mHz")
        tau = TO.fft_axis(f, "us")
        nedge = int(rng.choice([8, 22, 64, 130]))
        lim = float(rng.uniform(0.2, 1.6)) * fd.max()
        edges = np.linspace(-lim, lim, nedge)
        if trial % 5 == 4:      # uneven (still an odd number of centres, one at 0)
            edges = np.sort(np.concatenate((edges[:nedge // 2] * rng.uniform(0.8, 1.0),
                                            -edges[:nedge // 2][::-1] * rng.uniform(0.8, 1.0))))
        CS = rng.normal(size=(ntau, nfd)) + 1j * rng.normal(size=(ntau, nfd))
        eta = float(10 ** rng.uniform(-1.5, 1.5) * tau.max() / max(lim, 1e-9) ** 2)
        try:
            th, ti, fi, pn = TO.thth_indices(tau, fd, eta, edges)
        except Exception:
            continue
        try:
            ref = TO.thth_map(CS, tau, fd, eta, edges)
            ref_err = None
        except IndexError as e:
            ref, ref_err = None, e
        if ref_err is not None:
            with pytest.raises(IndexError):
                thth.thth_map(CS, tau, fd, eta, edges)
            continue
        m, gti, gfi, gpn = thth.thth_map(CS, tau, fd, eta, edges, return_indices=True)
        big = 2 ** 31 - 1
        assert np.array_equal(gti, np.clip(ti, -big - 1, big).astype(np.int32)), trial
        assert np.array_equal(gfi, np.clip(fi, -big - 1, big).astype(np.int32)), trial
        assert np.array_equal(gpn, pn), trial
        assert np.array_equal(thth.th_points(tau, fd, eta, edges),
                              TO.th_points(tau, fd, eta, edges)), trial
        if np.abs(ref).max() > 0:
            assert maxrel(m, ref) < 1e-6, trial
        assert np.array_equal(m == 0, ref == 0), trial
        checked += 1
    assert checked >= 12


@pytest.mark.parametrize("shape", [(3, 17), (37, 53), (5, 300), (127, 64), (64, 1000), (2, 16)])
def test_sspec_acf_random_shapes(sb, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    dyn = rng.exponential(1.0, shape)
    dt, df = 7.5, 0.3
    ds = sb.Dynspec(dyn=sb.BasicDyn(dyn, times=dt * np.arange(shape[1]),
                                    freqs=1200 + df * np.arange(shape[0]), dt=dt, df=df),
                    verbose=False)
    for kw in (dict(), dict(window="hamming", window_frac=0.3), dict(window=None, halve=False)):
        fdop, tdel, sec = ds.calc_sspec(return_sspec=True, **kw)
        rf, rt, ref = DO.calc_sspec(dyn, dt, df, **kw)
        assert sec.shape == ref.shape
        assert np.array_equal(fdop, rf) and np.array_equal(tdel, rt)
        assert maxrel(10 ** (sec / 10), 10 ** (ref / 10)) < 1e-5
    ds.calc_acf()
    assert ds.acf.shape == (2 * shape[0], 2 * shape[1])
    assert maxrel(ds.acf, DO.calc_acf(dyn)) < 1e-5
    # NaN in the input poisons the ACF like the reference (no nan_to_num)
    bad = dyn.copy()
    bad[0, 0] = np.nan
    ds2 = sb.Dynspec(dyn=sb.BasicDyn(bad, times=dt * np.arange(shape[1]),
                                     freqs=1200 + df * np.arange(shape[0]), dt=dt, df=df),
                     verbose=False)
    ds2.calc_acf()
    assert np.isnan(ds2.acf).all() and np.isnan(DO.calc_acf(bad)).all()


def test_eta_sweep_large_grid_direct_path(sb):
    """theta grids above 512 centres use the direct-load eigen kernel."""
    rng = np.random.default_rng(77)
    nf, nt, npad = 64, 256, 1
    t = np.arange(nt) * 10.0
    f = 1400 + 0.1 * np.arange(nf)
    fdk = rng.uniform(-20, 20, 20)
    ak = (rng.normal(size=20) + 1j * rng.normal(size=20)) * np.exp(-(fdk / 10) ** 2)
    E = sum(a * np.exp(2j * np.pi * (k * 1e-3 * t[None, :] - 0.004 * k ** 2 * (f[:, None] - f[0])))
            for a, k in zip(ak, fdk))
    d = np.abs(E) ** 2
    d -= d.mean()
    tau = TO.fft_axis(f, "us", npad)
    fd = TO.fft_axis(t, "mHz", npad)
    edges = np.linspace(-24, 24, 700)          # 699 centres
    etas = np.linspace(0.002, 0.008, 7)
    cs = sb.ththmod.conjugate_spectrum(d, npad, 0.0)
    got, info = sb.ththmod.eta_sweep(cs, tau, fd, etas, edges, return_info=True)
    ref = TO.eta_sweep(cs.numpy(), tau, fd, etas, edges)
    assert info["nred"].max() > 512
    assert (np.abs(got - ref) / ref).max() < 1e-5
