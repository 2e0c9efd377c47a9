"""Device / API time of the two widened rows, run on
the GPU box: single_chunk_retrieval (rev_map + top eigenpair + ifft2) and the
thin theta-theta sweep, each with the CPU oracle timed beside it on the same
input.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scintools_b200 import _device as D, _lib, ththmod as T
from oracle import thth_oracle as TO

D.device()
rng = np.random.default_rng(3)
nf, nt, npad = 256, 512, 3            # padded CS 1024 x 2048
t = np.arange(nt) * 10.0
f = 1400.0 + np.arange(nf) * 0.03125
fdk = rng.uniform(-20, 20, 48)
ak = (rng.normal(size=48) + 1j * rng.normal(size=48)) * np.exp(-(fdk / 10) ** 2)
eta_true = 0.3
E = sum(a * np.exp(2j * np.pi * (k * 1e-3 * t[None, :] - eta_true * k ** 2 * (f[:, None] - f[0])))
        for a, k in zip(ak, fdk))
dyn = np.abs(E) ** 2 + rng.normal(0, 0.02, (nf, nt))
d0 = dyn - dyn.mean()
fd = TO.fft_axis(t, "mHz", npad)
tau = TO.fft_axis(f, "us", npad)
edges = np.linspace(-24, 24, 512)

def wall(fn, reps):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r

def dev_ms(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = {"chunk": [nf, nt], "cs": [len(tau), len(fd)], "n_edges": len(edges)}
params = (d0, edges, t, f, eta_true, 0, 0, npad, 0.0, False)
w, res = wall(lambda: T.single_chunk_retrieval(params), 3)
t0 = time.perf_counter(); ref = TO.single_chunk_retrieval(d0, edges, t, f, eta_true, npad); tc = time.perf_counter() - t0
z = np.vdot(res[0], ref); al = res[0] * (z / abs(z))
out["single_chunk_retrieval"] = {"api_wall_ms": w * 1e3, "cpu_oracle_s": tc, "speedup": tc / w,
                                 "max_rel_err_abs": float(np.abs(np.abs(res[0]) - np.abs(ref)).max() / np.abs(ref).max()),
                                 "max_rel_err_phase_aligned": float(np.abs(al - ref).max() / np.abs(ref).max())}
# stages on the device
cs = T.conjugate_spectrum(d0, npad, None)
thth_red, edges_red = T.thth_redmap(cs, tau, fd, eta_true, edges)
n = thth_red.shape[0]
a = D.upload_f32(thth_red)
wd = D.empty((1,), torch.float64); V = D.empty((n, 2), torch.float32); info = D.zeros((2,), torch.int32)
L = _lib.lib
out["n_red"] = n
out["herm_eigvec_ms"] = dev_ms(lambda: L.sb_herm_eigvec(a.data_ptr(), n, n, 0.0, 0, wd.data_ptr(), V.data_ptr(), info.data_ptr(), D.stream_ptr()))
out["herm_eigvec_steps"] = int(info.cpu()[0])
th = D.upload(T.theta_centres(edges_red))
recov = D.empty((len(tau), len(fd), 2), torch.float32)
out["rev_map_ms"] = dev_ms(lambda: L.sb_rev_map(a.data_ptr(), n, th.data_ptr(), eta_true, float(tau[0]), float(tau[1] - tau[0]), len(tau), float(fd[0]), float(fd[1] - fd[0]), len(fd), 1, recov.data_ptr(), D.stream_ptr()))
outb = D.empty((nf, nt, 2), torch.float32)
out["ifft2_ms"] = dev_ms(lambda: L.sb_ifft2_c2c_f32(recov.data_ptr(), len(tau), len(fd), 1, nf, nt, 1.0, 0, outb.data_ptr(), D.stream_ptr()))
alg = 8 * len(tau) * len(fd) + 8 * nf * nt
out["ifft2_algorithmic_GB"] = alg / 1e9
out["ifft2_achieved_GBs"] = alg / out["ifft2_ms"] / 1e6
# thin sweep
etas = np.linspace(0.1, 0.6, 128)
arc = edges[np.abs(edges) < 12]
csf = T.conjugate_spectrum(d0, npad, None)
w, sv = wall(lambda: T.thin_sweep(csf, tau, fd, etas, edges, arc, 0.5), 3)
CS = csf.numpy()
t0 = time.perf_counter(); r2 = TO.thin_sweep(CS, tau, fd, etas[::32], edges, arc, 0.5); tc = (time.perf_counter() - t0) / 4
out["thin_sweep"] = {"etas": len(etas), "api_wall_ms": w * 1e3, "per_eta_ms": w * 1e3 / len(etas),
                     "cpu_oracle_s_per_eta": tc, "speedup": tc / (w / len(etas)),
                     "max_rel_err": float(np.abs(sv[::32] - r2).max() / r2.max())}
print(json.dumps(out))
