"""GPU: BASELINE.json full-size configurations checked through size-independent
properties (Parseval, symmetry, normalisation, known answer) plus exact
float64 spot checks of individual bins / lags computed on the CPU in O(N) each.

C2: 4096x8192 calc_sspec + calc_acf.   C3: CS 16384x32768 + 1024-eta sweep.
C4 (reduced count): one 8192^2 Simulation realisation, energy conservation.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import dynspec_oracle as DO   # noqa: E402
from oracle import sim_oracle as SO       # noqa: E402
from oracle import thth_oracle as TO      # noqa: E402

NF, NT = 4096, 8192


@pytest.fixture(scope="module")
def sb():
    import scintools_b200
    from scintools_b200 import _device
    _device.device()
    return scintools_b200


@pytest.fixture(scope="module")
def dyn_c2():
    rng = np.random.default_rng(2)
    return rng.exponential(1.0, (NF, NT)).astype(np.float32)


def _windowed(dyn, frac=0.1):
    x = dyn.astype(np.float64)
    x -= x.mean()
    cw, sw = DO.get_window(x.shape[1], x.shape[0], "hanning", frac)
    x *= cw[None, :]
    x *= sw[:, None]
    x -= x.mean()
    return x


def test_c2_sspec_parseval_and_bins(sb, dyn_c2):
    dt, df = 10.0, 0.03125
    ds = sb.Dynspec(dyn=sb.BasicDyn(dyn_c2, times=dt * np.arange(NT),
                                    freqs=1400 + df * np.arange(NF), dt=dt, df=df),
                    verbose=False)
    fdop, tdel, sec = ds.calc_sspec(halve=False, return_sspec=True, dtype=np.float32)
    assert sec.shape == (2 * NF, 2 * NT)
    x = _windowed(dyn_c2)
    lin = np.power(10.0, sec.astype(np.float64) / 10.0)
    # Parseval: sum |X|^2 = N * sum x^2
    assert abs(lin.sum() / (lin.size * (x ** 2).sum()) - 1) < 1e-5
    # point symmetry of the power spectrum of a real array
    assert np.allclose(lin[1:, 1:], lin[1:, 1:][::-1, ::-1], rtol=1e-4, atol=1e-6 * lin.max())
    # exact float64 DFT of a few bins
    nr, nc = 2 * NF, 2 * NT
    f = np.arange(NF)[:, None]
    t = np.arange(NT)[None, :]
    for (kr, kc) in ((0, 0), (1, 3), (17, 4001), (4095, 8191), (5000, 12000)):
        ph = np.exp(-2j * np.pi * (kr * f / nr)) * np.exp(-2j * np.pi * (kc * t / nc))
        X = (x * ph).sum()
        got = lin[(kr + nr // 2) % nr, (kc + nc // 2) % nc]
        assert abs(got - abs(X) ** 2) <= 1e-5 * lin.max()
    # halved output = rows nrfft/2.. of the full one
    ds.calc_sspec(dtype=np.float32)
    assert ds.sspec.shape == (NF, 2 * NT)
    assert np.array_equal(ds.sspec, sec[NF:])


def test_c2_acf_properties(sb, dyn_c2):
    ds = sb.Dynspec(dyn=sb.BasicDyn(dyn_c2, times=10.0 * np.arange(NT),
                                    freqs=1400 + 0.03125 * np.arange(NF), dt=10.0,
                                    df=0.03125), verbose=False)
    ds.calc_acf(dtype=np.float32)
    acf = ds.acf
    assert acf.shape == (2 * NF, 2 * NT)
    assert acf[NF, NT] == pytest.approx(1.0, abs=1e-6) and acf.max() <= 1.0 + 1e-6
    # autocovariance is even: acf[nf+a, nt+b] == acf[nf-a, nt-b]
    a = acf[1:, 1:]
    assert np.allclose(a, a[::-1, ::-1], atol=2e-6)
    # lag -nf / -nt rows are empty (no overlap)
    assert np.abs(acf[0]).max() < 1e-6 and np.abs(acf[:, 0]).max() < 1e-6
    # exact float64 lags
    x = dyn_c2.astype(np.float64)
    x -= x.mean()
    z = (x * x).sum()
    for (lf, lt) in ((0, 1), (3, 0), (-2, 5), (100, -300), (4000, 8000)):
        xa = x[max(0, lf):NF + min(0, lf), max(0, lt):NT + min(0, lt)]
        xb = x[max(0, -lf):NF + min(0, -lf), max(0, -lt):NT + min(0, -lt)]
        want = (xa * xb).sum() / z
        assert abs(acf[NF + lf, NT + lt] - want) < 1e-5


def test_c3_cs_and_sweep(sb):
    import bench
    thth = sb.ththmod
    dyn, freq, t = bench.make_dynspec()
    npad = bench.NPAD
    tau = TO.fft_axis(freq, "us", npad)
    fd = TO.fft_axis(t, "mHz", npad)
    cs = thth.conjugate_spectrum(dyn, npad, 0.0, half=True)
    ntau, nfd = cs.shape
    assert (ntau, nfd) == (16384, 32768)
    H = cs.t.cpu().numpy()                      # [ntau][pitch][2], unshifted cols
    h = nfd // 2
    P = (H[:, :h + 1, 0].astype(np.float64) ** 2 + H[:, :h + 1, 1].astype(np.float64) ** 2)
    wgt = np.full(h + 1, 2.0)
    wgt[0] = wgt[h] = 1.0
    x = dyn.astype(np.float64)
    assert abs((P * wgt).sum() / (ntau * nfd * (x ** 2).sum()) - 1) < 1e-5
    # exact float64 DFT of a few bins (shifted row r <-> k = r - ntau/2)
    f = np.arange(dyn.shape[0])[:, None]
    tt = np.arange(dyn.shape[1])[None, :]
    scale = np.sqrt(P.max())
    for (kr, kc) in ((0, 0), (5, 7), (-300, 1234), (8000, 16000), (-8192, 16384)):
        ph = np.exp(-2j * np.pi * (kr * f / ntau)) * np.exp(-2j * np.pi * (kc * tt / nfd))
        X = (x * ph).sum()
        g = H[(kr + ntau // 2) % ntau, kc]
        assert abs(complex(g[0], g[1]) - X) <= 1e-5 * scale
    # sweep: known answer eta_true and oracle eigenvalues at a few curvatures
    edges = np.linspace(-bench.EDGE_LIM, bench.EDGE_LIM, bench.NEDGE)
    etas = bench.eta_grid(bench.NETA)
    eigs, info = thth.eta_sweep(cs, tau, fd, etas, edges, return_info=True)
    assert (info["status"] == 0).all()
    assert abs(etas[np.argmax(eigs)] / bench.ETA_TRUE - 1) < 0.02
    # 64 curvatures across the grid (plus the four corners of the old check) against the
    # oracle (numpy gather + ARPACK) on the same conjugate spectrum, full 511 x 511 maps
    CS_host = cs.numpy().astype(np.complex64)
    # incl. the slowly converging curvatures that exposed the round-1 stopping bug
    pick = sorted(set(list(range(0, bench.NETA, 16)) + [400, 700, 1023, 95, 118, 119, 120, 121, 162,
                                                         174, 325, 784, 809, 810, 905]))
    worst = 0.0
    for i in pick:
        ref = TO.Eval_calc(CS_host, tau, fd, etas[i], edges)
        worst = max(worst, abs(eigs[i] - ref) / ref)
    print("C3 full size: max rel eigenvalue error over %d etas = %.2e" % (len(pick), worst))
    assert worst < 1e-5
    # bit-exact crop sizes
    nred = np.array([TO.th_points(tau, fd, e, edges).sum() for e in etas[::64]])
    assert np.array_equal(info["nred"][::64], nred)


def test_sim_2048_vs_oracle(sb):
    from scintools_b200.scint_sim import Simulation
    rng = np.random.default_rng(8)
    n, nf = 2048, 3
    n1, n2 = rng.normal(size=(n, n)), rng.normal(size=(n, n))
    kw = dict(mb2=2, ns=n, nf=nf, dlam=0.25)
    ref = SO.SimOracle(noise_re=n1, noise_im=n2, **kw)
    got = Simulation(noise=(n1, n2), **kw)
    mr = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))  # noqa: E731
    assert mr(got.xyp, ref.xyp) < 1e-9
    assert mr(got.spe, ref.spe) < 1e-5
    assert mr(got.xyi, ref.xyi) < 1e-5


def test_c4_sim_8192_energy(sb):
    """One C4 realisation (8192^2, device noise): the propagation is unitary,
    so mean(xyi) == 1 exactly; the dynamic spectrum has unit mean."""
    from scintools_b200.scint_sim import Simulation
    s = Simulation(mb2=2, ns=8192, nf=4, dlam=0.25, seed=0, device_rng=True)
    assert s.xyp.shape == (8192, 8192) and np.isfinite(s.xyp).all()
    assert abs(s.xyi.mean() - 1.0) < 1e-5
    assert s.dyn.shape == (4, 8192)
    assert abs(s.dyn.mean() - 1.0) < 0.25
    # Kolmogorov screen: rms phase difference grows with separation
    d1 = np.std(s.xyp[:, 1:] - s.xyp[:, :-1])
    d64 = np.std(s.xyp[:, 64:] - s.xyp[:, :-64])
    assert d64 > 5 * d1
