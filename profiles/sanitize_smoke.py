"""Small invocation of every device path for compute-sanitizer (memcheck / racecheck):
  compute-sanitizer --tool memcheck  python profiles/sanitize_smoke.py
  compute-sanitizer --tool racecheck python profiles/sanitize_smoke.py
Sizes are tiny (the tools slow kernels down 10-100x): sweep (default fp16-iterate solver,
fp32 solver, thin), conjugate spectrum (radix + chirp-z), sspec / acf (TMA tile path),
scale_dyn, norm_sspec / fit_arc, phase retrieval (pow2 + chirp-z), Simulation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scintools_b200 import Dynspec, BasicDyn, ththmod as thth
from scintools_b200.scint_sim import Simulation

rng = np.random.default_rng(0)
nf, nt, npad = 64, 128, 3
dt, df = 10.0, 0.05
t = np.arange(nt) * dt
f = 1400.0 + np.arange(nf) * df
fdk = rng.uniform(-3, 3, 16)
ak = (rng.normal(size=16) + 1j * rng.normal(size=16)) * np.exp(-(fdk / 2) ** 2)
E = sum(a * np.exp(2j * np.pi * (fd_ * 1e-3 * t[None, :] - 40.0 * fd_ ** 2 * (f[:, None] - f[0])))
        for a, fd_ in zip(ak, fdk))
dyn = np.abs(E) ** 2
d0 = dyn - dyn.mean()
fd = thth.fft_axis(t, "mHz", npad)
tau = thth.fft_axis(f, "us", npad)
edges = np.linspace(-4, 4, 128)
etas = np.linspace(20, 60, 24)
cs = thth.conjugate_spectrum(d0, npad, 0.0)
e1 = thth.eta_sweep(cs, tau, fd, etas, edges)
os.environ["SB_EIG_FP32"] = "1"
e2 = thth.eta_sweep(cs, tau, fd, etas, edges)
del os.environ["SB_EIG_FP32"]
assert np.nanmax(np.abs(e1 - e2) / e2) < 1e-5
res = thth.single_search([d0, f, t, etas, edges, None, False, 0.1, npad, True, 0.0, False])
ds = Dynspec(dyn=BasicDyn(dyn, times=t, freqs=f, dt=dt, df=df), verbose=False)
ds.calc_sspec(); ds.calc_acf(); ds.calc_sspec(lamsteps=True)
ds.norm_sspec(eta=float(np.median(ds.beta[1:]) / 4.0), lamsteps=True, maxnormfac=2)
thth.single_chunk_retrieval((d0, edges, t, f, 40.0, 0, 0, npad, 0.0, False))
d1 = d0[:48, :100]          # non-power-of-two padded sizes: chirp-z forward + inverse
thth.single_chunk_retrieval((d1, edges, t[:100], f[:48], 40.0, 0, 0, npad, 0.0, False))
arc = edges[np.abs(edges) < 2.0]
thth.single_search_thin([d0, f, t, etas[:8], edges, None, False, 0.1, npad, True, False, arc, 0.0])
Simulation(mb2=2, ns=128, nf=4, dlam=0.25, seed=1)
print("sanitize smoke ok", float(np.asarray(res[0])))
