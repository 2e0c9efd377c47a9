"""GPU parity of scint_sim.Simulation against the fixtures generated from the
reference (legacy MT19937 noise drawn on the host exactly as the reference
does) and against the numpy oracle with explicit noise."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import sim_oracle as SO   # noqa: E402

RTOL = 1e-5


def maxrel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.fixture(scope="module")
def Sim():
    from scintools_b200 import _device
    from scintools_b200.scint_sim import Simulation
    _device.device()
    return Simulation


@pytest.mark.parametrize("tag", ["iso", "lam", "aniso2", "strong"])
def test_simulation_golden(Sim, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "sim_small.npz"))
    with open(os.path.join(golden_dir, "sim_small_cfg.json")) as f:
        kw = json.load(f)[tag]
    s = Sim(**kw)
    assert s.w.shape == g[tag + "_w"].shape
    assert maxrel(s.w, g[tag + "_w"]) < 1e-12
    assert maxrel(s.xyp, g[tag + "_xyp"]) < 1e-10
    assert s.spe.shape == g[tag + "_spe"].shape and s.spe.dtype == np.csingle
    assert maxrel(s.spe, g[tag + "_spe"]) < RTOL
    assert maxrel(s.spi, g[tag + "_spi"]) < RTOL
    assert maxrel(s.xyi, g[tag + "_xyi"]) < RTOL
    assert s.dyn.shape == g[tag + "_dyn"].shape
    assert maxrel(s.dyn, g[tag + "_dyn"]) < RTOL
    np.testing.assert_allclose(s.freqs, g[tag + "_freqs"], rtol=1e-13)
    np.testing.assert_allclose(s.times, g[tag + "_times"], rtol=1e-13)
    assert s.eta == pytest.approx(float(g[tag + "_eta"]), rel=1e-13)
    assert s.df == pytest.approx(float(g[tag + "_df"]), rel=1e-13)


def test_simulation_explicit_noise_vs_oracle(Sim):
    rng = np.random.default_rng(42)
    nx, ny, nf = 256, 512, 5
    n1, n2 = rng.normal(size=(nx, ny)), rng.normal(size=(nx, ny))
    kw = dict(mb2=8, ar=1.3, psi=15, nx=nx, ny=ny, nf=nf, dlam=0.2, inner=0.002)
    ref = SO.SimOracle(noise_re=n1, noise_im=n2, **kw)
    got = Sim(noise=(n1, n2), **kw)
    assert maxrel(got.w, ref.w) < 1e-12
    assert maxrel(got.xyp, ref.xyp) < 1e-10
    assert maxrel(got.spe, ref.spe) < RTOL
    assert maxrel(got.xyi, ref.xyi) < RTOL
    assert maxrel(got.dyn, ref.dyn) < RTOL


def test_simulation_device_rng_statistics(Sim):
    """Counter-based device noise: same screen statistics as host noise."""
    kw = dict(mb2=2, ns=256, nf=2, dlam=0.25)
    a = Sim(seed=1, device_rng=True, **kw)
    b = Sim(seed=2, device_rng=True, **kw)
    c = Sim(seed=1, device_rng=True, **kw)
    h = Sim(seed=1, **kw)
    assert np.array_equal(a.xyp, c.xyp) and not np.array_equal(a.xyp, b.xyp)
    assert abs(a.xyp.std() / h.xyp.std() - 1) < 0.35
    assert abs(np.mean(a.dyn) - 1) < 0.2


def test_simulation_feeds_dynspec(Sim):
    from scintools_b200 import Dynspec
    s = Sim(mb2=2, ns=128, nf=64, dlam=0.1, seed=4)
    ds = Dynspec(dyn=s, verbose=False)
    ds.calc_sspec()
    assert ds.sspec.shape == (64, 256) and np.isfinite(ds.sspec).any()
    assert ds.eta == s.eta


def test_simulation_lazy_matches_eager(Sim):
    """lazy=True keeps w / xyp / xyi on the device until they are asked for (batch
    production); every attribute equals the eager object's."""
    Simulation = Sim
    kw = dict(mb2=2, ns=128, nf=6, dlam=0.25, seed=7)
    a = Simulation(**kw)
    b = Simulation(lazy=True, **kw)
    assert "xyp" not in b.__dict__ and "w" not in b.__dict__ and "xyi" not in b.__dict__
    for name in ("dyn", "spe", "spi", "dm", "pulsewin"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    for name in ("w", "xyp", "xyi"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
        assert name in b.__dict__
    with pytest.raises(AttributeError):
        b.no_such_attribute
