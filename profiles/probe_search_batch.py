"""Round-2 first-call probe (not yet run): where does ththmod.search_batch lose
its time?  The one measurement of round 1 (5 x 134 MB pinned dynspecs, no warm-up
of the copy stream) gave 13.6 k eta-trials/s against 84-99 k for a plain loop
over single_search.  Times, on the bench workload: the plain loop, search_batch
cold and warm, and the bare upload on the copy stream with / without compute
running on the main stream.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from scintools_b200 import _device as D, ththmod as thth

D.device()
dyn, freq, t = bench.make_dynspec()
edges = np.linspace(-bench.EDGE_LIM, bench.EDGE_LIM, bench.NEDGE)
etas = bench.eta_grid(bench.NETA)
h_dyn = torch.from_numpy(dyn).pin_memory()
params = [h_dyn.numpy(), freq, t, etas, edges, None, False, bench.FW, bench.NPAD, True, 0.0, False]
out = {}

def wall(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3

thth.single_search(params); thth.single_search(params)
out["loop5_ms"] = wall(lambda: [thth.single_search(params) for _ in range(5)])
out["batch5_cold_ms"] = wall(lambda: thth.search_batch([params] * 5))
out["batch5_warm_ms"] = wall(lambda: thth.search_batch([params] * 5))
out["batch1_ms"] = wall(lambda: thth.search_batch([params]))
side = torch.cuda.Stream()
def up(n):
    with torch.cuda.stream(side):
        ts = [torch.from_numpy(h_dyn.numpy()).to(D.device(), non_blocking=True) for _ in range(n)]
    side.synchronize()
    return ts
up(1)
out["upload_side_stream_ms"] = wall(lambda: up(3)) / 3
out["upload_default_stream_ms"] = wall(lambda: [D.upload_f32(h_dyn.numpy()) for _ in range(3)]) / 3
out["is_pinned_view"] = bool(torch.from_numpy(h_dyn.numpy()).is_pinned())
cs = thth.conjugate_spectrum(h_dyn.numpy(), bench.NPAD, None)
fd = thth.U.value(thth.fft_axis(t, "mHz", bench.NPAD), "mHz")
tau = thth.U.value(thth.fft_axis(freq, "us", bench.NPAD), "us")
def overlapped():
    with torch.cuda.stream(side):
        tt = torch.from_numpy(h_dyn.numpy()).to(D.device(), non_blocking=True)
    thth.eta_sweep(cs, tau, fd, etas, edges, True)
    side.synchronize()
    return tt
overlapped()
out["sweep_alone_ms"] = wall(lambda: thth.eta_sweep(cs, tau, fd, etas, edges, True))
out["sweep_with_concurrent_upload_ms"] = wall(overlapped)
print(json.dumps(out))
