"""Where does the end-to-end (host API) time go?  Run on the GPU box."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from scintools_b200 import ththmod as thth, _device as D, _lib

dyn, freq, t = bench.make_dynspec()
h = torch.from_numpy(dyn).pin_memory().numpy()
etas = bench.eta_grid(1024); edges = np.linspace(-10, 10, 512)
fd = np.asarray(thth.fft_axis(t, "mHz", 3)); tau = np.asarray(thth.fft_axis(freq, "us", 3))
def T(label, fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize()
    print("%-32s %8.2f ms" % (label, (time.perf_counter() - t0) / n * 1e3), flush=True)
    return r
D.device()
T("host mean", lambda: float(h.mean()))
T("upload_f32 (134 MB pinned)", lambda: D.upload_f32(h))
cs = T("conjugate_spectrum(host)", lambda: thth.conjugate_spectrum(h, 3, None, tau, 0.0))
T("eta_sweep(DeviceCS)", lambda: thth.eta_sweep(cs, tau, fd, etas, edges))
eigs = thth.eta_sweep(cs, tau, fd, etas, edges)
T("peak_fit", lambda: thth.peak_fit(etas, eigs, 0.1))
T("fft_axis x2", lambda: (thth.fft_axis(t, "mHz", 3), thth.fft_axis(freq, "us", 3)))
T("single_search", lambda: thth.single_search([h, freq, t, etas, edges, None, False, 0.1, 3, True, 0.0, False]))
