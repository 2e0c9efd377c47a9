"""Ground truth for the curvatures where the default and the fp32 solver disagree:
oracle (numpy gather + ARPACK eigsh, tol=0) on the GPU-built full-size CS."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from scintools_b200 import ththmod as thth
from oracle import thth_oracle as TO

dyn, freq, t = bench.make_dynspec()
fd = np.asarray(thth.fft_axis(t, "mHz", bench.NPAD)); tau = np.asarray(thth.fft_axis(freq, "us", bench.NPAD))
edges = np.linspace(-bench.EDGE_LIM, bench.EDGE_LIM, bench.NEDGE)
etas = bench.eta_grid(bench.NETA)
cs = thth.conjugate_spectrum(dyn, bench.NPAD, 0.0)
os.environ["SB_EIG_FP32"] = "1"
ref, iref = thth.eta_sweep(cs, tau, fd, etas, edges, return_info=True)
del os.environ["SB_EIG_FP32"]
got, info = thth.eta_sweep(cs, tau, fd, etas, edges, return_info=True)
os.environ["SB_EIG_ETOL_B"] = "1e-6"
got6, info6 = thth.eta_sweep(cs, tau, fd, etas, edges, return_info=True)
del os.environ["SB_EIG_ETOL_B"]
rel = np.abs(got - ref) / ref
rel6 = np.abs(got6 - got) / got
pick = sorted(set(np.argsort(rel)[-12:].tolist() + np.argsort(rel6)[-8:].tolist() + [121, 162, 174, 325, 809]))
CS_host = cs.numpy().astype(np.complex64)
rows = []
for i in pick:
    A, _ = TO.thth_redmap(CS_host.astype(np.complex128), tau, fd, etas[i], edges)
    ev = np.linalg.eigvalsh(A)
    arp = TO.Eval_calc(CS_host, tau, fd, etas[i], edges)
    rows.append(dict(i=int(i), default=float(got[i]), fp32=float(ref[i]), arpack=float(arp), top=float(ev[-1]),
                     err_etol1e6=float(abs(got6[i] - ev[-1]) / ev[-1]), it_etol1e6=int(info6["iters"][i]),
                     second=float(ev[-2]), it_default=int(info["iters"][i]), it_fp32=int(iref["iters"][i]),
                     err_default=float(abs(got[i] - ev[-1]) / ev[-1]), err_fp32=float(abs(ref[i] - ev[-1]) / ev[-1]),
                     gap_rel=float((ev[-1] - ev[-2]) / ev[-1])))
print(json.dumps(rows))
