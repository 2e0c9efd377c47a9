"""CPU numerics experiment (no GPU): can the Lanczos iteration of the
theta-theta eigenvalue run on a bf16 / fp16 copy of the matrix (half the bytes
per step) if the final value is the Rayleigh quotient of the Ritz vector with
the fp32 matrix?  Emulates the kernel's arithmetic: fp32 matrix and vectors,
fp64 reductions, fp16-stored basis.  Uses the bench recipe at reduced size
(1024 x 2048 dynspec, npad 3, 512-point theta grid).  Prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from oracle import thth_oracle as TO

nf, nt = 512, 1024
dyn, freq, t = bench.make_dynspec(nf=nf, nt=nt)
dyn = dyn.astype(np.float64)
fd = TO.fft_axis(t, "mHz", bench.NPAD)
tau = TO.fft_axis(freq, "us", bench.NPAD)
CS = TO.conjugate_spectrum(dyn, bench.NPAD, None).astype(np.complex64)
edges = np.linspace(-bench.EDGE_LIM, bench.EDGE_LIM, bench.NEDGE)


def bf16(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def quant(A, kind):
    if kind == "bf16":
        return (bf16(A.real) + 1j * bf16(A.imag)).astype(np.complex64)
    s = np.abs(A).max() / 60000.0
    return ((A.real / s).astype(np.float16).astype(np.float32) * s +
            1j * ((A.imag / s).astype(np.float16).astype(np.float32) * s)).astype(np.complex64)


def lanczos(A32, tol_e=2e-7, basis_dtype=None, maxit=64):
    """fp32 vectors, fp64 reductions; stop on res^2 <= tol_e * theta * gap."""
    n = A32.shape[0]
    v = A32[n // 2].astype(np.complex64)
    v = (v / np.sqrt(np.vdot(v, v).real)).astype(np.complex64)
    vp = np.zeros(n, np.complex64)
    al, be, V = [], [0.0], []
    b = np.float32(0)
    for it in range(maxit):
        V.append(v if basis_dtype is None else
                 (v.real.astype(basis_dtype).astype(np.float32) + 1j * v.imag.astype(basis_dtype).astype(np.float32)))
        w = (A32 @ v).astype(np.complex64)
        a = float(np.vdot(v.astype(np.complex128), w.astype(np.complex128)).real)
        w = (w - np.float32(a) * v - b * vp).astype(np.complex64)
        b2 = float(np.vdot(w.astype(np.complex128), w.astype(np.complex128)).real)
        al.append(a); be.append(np.sqrt(b2))
        m = it + 1
        T = np.diag(al) + np.diag(be[1:m], 1) + np.diag(be[1:m], -1)
        ev, S = np.linalg.eigh(T)
        res = be[m] * abs(S[-1, -1])
        gap = ev[-1] - ev[-2] if m > 1 else 0.0
        if m >= 3 and (res * res <= tol_e * abs(ev[-1]) * gap or res <= 2e-5 * abs(ev[-1])):
            break
        vp = v
        b = np.float32(be[m])
        v = (w / b).astype(np.complex64)
    y = (np.array(V[:m]).T.astype(np.complex128) @ S[:, -1]).astype(np.complex64)
    return ev[-1], y, m


out = {"n_edges": bench.NEDGE, "cs": list(CS.shape), "cases": []}
etas = bench.ETA_TRUE * np.array([0.5, 0.8, 0.97, 1.0, 1.03, 1.3, 2.0])
for eta in etas:
    A, _ = TO.thth_redmap(CS.astype(np.complex128), tau, fd, eta, edges)
    lam = np.linalg.eigvalsh(A)[-1]
    A32 = A.astype(np.complex64)
    th32, _, m32 = lanczos(A32)
    row = {"eta": float(eta), "n": int(A.shape[0]), "fp32_steps": m32,
           "fp32_ritz_err": abs(th32 - lam) / lam}
    for kind in ("bf16", "fp16"):
        thq, y, mq = lanczos(quant(A32, kind), basis_dtype=np.float16)
        Ay = (A32 @ y).astype(np.complex64)
        rq = float(np.vdot(y.astype(np.complex128), Ay.astype(np.complex128)).real /
                   np.vdot(y.astype(np.complex128), y.astype(np.complex128)).real)
        row[kind] = {"steps": mq, "ritz_err_quantised": abs(thq - lam) / lam,
                     "rayleigh_fp32_err": abs(rq - lam) / lam}
    out["cases"].append(row)
print(json.dumps(out))
