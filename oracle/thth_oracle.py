"""Unit-free CPU restatement of the theta-theta curvature sweep.

TEST INFRASTRUCTURE (see oracle/__init__.py).  float64 / complex128 numpy
throughout, same operation order as the reference so that index arrays are
bit-identical.

Follows (reference = /root/reference/scintools):
  fft_axis        ththmod.py:473-493
  thth_map        ththmod.py:56-116
  thth_redmap     ththmod.py:119-173
  Eval_calc       ththmod.py:371-401
  single_search   ththmod.py:715-895   (eta loop + parabola fit)
  min_edges       ththmod.py:1671-1705
  chi_par         ththmod.py:38-53
  eta grid        dynspec.py:1476-1478, 1583-1584

Units: tau [us], fd/edges/theta [mHz], eta [s^3], time [s], freq [MHz].
"""
import numpy as np
from scipy.optimize import curve_fit
from scipy.sparse.linalg import eigsh

# 1/s -> mHz and 1/MHz -> us conversion factors applied by
# ``Quantity.to_value`` in ththmod.py:488-490.
_SCALE = {"mHz": 1.0e3, "us": 1.0}


def fft_axis(x, unit, pad=0):
    """ththmod.py:473-493. ``x`` in s (unit='mHz') or MHz (unit='us')."""
    x = np.asarray(x, dtype=np.float64)
    f = np.fft.fftfreq((pad + 1) * x.shape[0], x[1] - x[0])
    scale = _SCALE[unit]
    if scale != 1.0:
        f = f * scale
    return np.fft.fftshift(f)


def theta_centres(edges):
    """ththmod.py:83-84 (identical lines at :151-152)."""
    edges = np.asarray(edges, dtype=np.float64)
    th = (edges[1:] + edges[:-1]) / 2
    th = th - th[np.abs(th) == np.abs(th).min()]  # needs an odd centre count
    return th


def thth_indices(tau, fd, eta, edges):
    """Index/mask part of thth_map (ththmod.py:83-100).

    Returns th_cents, tau_inv, fd_inv (int64 N x N) and pnts (bool N x N).
    """
    tau = np.asarray(tau, dtype=np.float64)
    fd = np.asarray(fd, dtype=np.float64)
    th = theta_centres(edges)
    n = th.shape[0]
    th1 = np.ones((n, n)) * th          # th1[i, j] = th[j]
    th2 = th1.T                         # th2[i, j] = th[i]
    dtau = np.diff(tau).mean()
    dfd = np.diff(fd).mean()
    tau_inv = (((eta * (th1 ** 2 - th2 ** 2)) - tau[0] + dtau / 2)
               // dtau).astype(int)
    fd_inv = (((th1 - th2) - fd[0] + dfd / 2) // dfd).astype(int)
    pnts = (tau_inv > 0) * (tau_inv < tau.shape[0]) * (fd_inv < fd.shape[0])
    return th, tau_inv, fd_inv, pnts


def thth_map(CS, tau, fd, eta, edges, hermetian=True):
    """ththmod.py:56-116."""
    th, tau_inv, fd_inv, pnts = thth_indices(tau, fd, eta, edges)
    n = th.shape[0]
    th1 = np.ones((n, n)) * th
    th2 = th1.T
    thth = np.zeros((n, n), dtype=complex)
    # negative fd_inv wraps (python indexing); < -nfd raises IndexError
    thth[pnts] = np.asarray(CS)[tau_inv[pnts], fd_inv[pnts]]
    thth *= np.sqrt(np.abs(2 * eta * (th2 - th1)))
    if hermetian:
        thth -= np.tril(thth)
        thth += np.conjugate(np.triu(thth).T)
        thth -= np.diag(np.diag(thth))
        thth -= np.diag(np.diag(thth[::-1, :]))[::-1, :]
        thth = np.nan_to_num(thth)
    return thth


def th_points(tau, fd, eta, edges):
    """Crop mask of thth_redmap (ththmod.py:151-156)."""
    tau = np.asarray(tau, dtype=np.float64)
    fd = np.asarray(fd, dtype=np.float64)
    th = theta_centres(edges)
    return ((th ** 2) * eta < np.abs(tau.max())) * \
        (np.abs(th) < np.abs(fd.max()) / 2)


def thth_redmap(CS, tau, fd, eta, edges, hermetian=True):
    """ththmod.py:119-173."""
    thth = thth_map(CS, tau, fd, eta, edges, hermetian)
    th = theta_centres(edges)
    sel = th_points(tau, fd, eta, edges)
    red = thth[sel, :][:, sel]
    er = th[sel]
    er = (er[:-1] + er[1:]) / 2
    step = np.diff(er).mean()
    edges_red = np.concatenate((np.array([er[0] - step]), er,
                                np.array([er[-1] + step])))
    return red, edges_red


def Eval_calc(CS, tau, fd, eta, edges, return_iters=False):
    """ththmod.py:371-401: |largest-algebraic eigenvalue| of thth_red."""
    red, _ = thth_redmap(CS, tau, fd, eta, edges)
    v0 = np.copy(red[red.shape[0] // 2, :])
    v0 /= np.sqrt((np.abs(v0) ** 2).sum())
    w, _ = eigsh(red, 1, v0=v0, which="LA")
    return np.abs(w[0])


def chi_par(x, A, x0, C):
    """ththmod.py:38-53."""
    return A * (x - x0) ** 2 + C


def conjugate_spectrum(dspec2, npad, pad_value=None, tau=None, tau_mask=0.0):
    """CS stage of single_search (ththmod.py:777-787).

    pad_value None -> dspec2.mean() (single_search); 0.0 reproduces
    Dynspec.thetatheta_single (dynspec.py:1575-1579).
    """
    dspec2 = np.asarray(dspec2, dtype=np.float64)
    if pad_value is None:
        pad_value = dspec2.mean()
    pad = np.pad(dspec2, ((0, npad * dspec2.shape[0]),
                          (0, npad * dspec2.shape[1])),
                 mode="constant", constant_values=pad_value)
    CS = np.fft.fftshift(np.fft.fft2(pad))
    if tau is not None:
        CS[np.abs(tau) < tau_mask] = 0
    return CS


def peak_fit(etas, eigs, fw):
    """Parabola fit of the eigenvalue peak (ththmod.py:813-859).

    Returns (eta_fit, eta_sig, popt); NaNs on failure as the reference does.
    """
    try:
        etas = np.asarray(etas, dtype=np.float64)
        eigs = np.asarray(eigs, dtype=np.float64)
        good = np.isfinite(eigs)
        etas = etas[good]
        eigs = eigs[good]
        pk = etas[eigs == eigs.max()]
        win = np.abs(etas - pk) < fw * pk
        ef = etas[win]
        gf = eigs[win]
        C = gf.max()
        x0 = ef[gf == C][0]
        if x0 == ef[0]:
            A = (gf[-1] - C) / ((ef[-1] - x0) ** 2)
        else:
            A = (gf[0] - C) / ((ef[0] - x0) ** 2)
        popt, _ = curve_fit(chi_par, ef, gf, p0=np.array([A, x0, C]))
        eta_fit = popt[1]
        eta_sig = np.sqrt((gf - chi_par(ef, *popt)).std() / np.abs(popt[0]))
        return eta_fit, eta_sig, popt
    except Exception:
        return np.nan, np.nan, None


def eta_sweep(CS, tau, fd, etas, edges):
    """Eta loop of single_search (ththmod.py:789-799): NaN on failure."""
    eigs = np.zeros(len(etas))
    for i, eta in enumerate(etas):
        try:
            eigs[i] = Eval_calc(CS, tau, fd, eta, edges)
        except Exception:
            eigs[i] = np.nan
    return eigs


def single_search(dspec2, freq, time, etas, edges, fw=0.1, npad=3,
                  coher=True, tau_mask=0.0):
    """ththmod.py:715-895 without plotting."""
    fd = fft_axis(time, "mHz", npad)
    tau = fft_axis(freq, "us", npad)
    CS = conjugate_spectrum(dspec2, npad, None, tau, tau_mask)
    src = CS if coher else np.abs(CS)
    eigs = eta_sweep(src, tau, fd, etas, edges)
    eta_fit, eta_sig, _ = peak_fit(etas, eigs, fw)
    return eta_fit, eta_sig, np.mean(freq), np.mean(time), eigs


def min_edges(fd_lim, fd, tau, eta, factor=2):
    """ththmod.py:1671-1705."""
    dtau_lim = (tau[1] - tau[0]) / factor
    dtau_lim /= 2 * eta * fd_lim
    dfd_lim = (fd[1] - fd[0]) / factor
    npoints = (2 * fd_lim) // (min(dfd_lim, dtau_lim))
    npoints += np.mod(npoints, 2)
    return np.linspace(-fd_lim, fd_lim, int(npoints))


def eta_grid(eta_min, eta_max, fw, fref, fmean):
    """dynspec.py:1476-1478 (neta) and :1583-1584 (log grid, nu^-2 scaling)."""
    l0 = np.log10(eta_min)
    l1 = np.log10(eta_max)
    neta = int(1 + (l1 - l0) / np.log10(1 + fw / 10))
    return np.logspace(np.log10(eta_min), np.log10(eta_max), neta) * \
        (fref / fmean) ** 2


# ---------------------------------------------------------------------------
# "thin" theta-theta: two-curvature map + top
# singular value.  Follows ththmod.py:1557-1636 (two_curve_map), :496-512
# (singularvalue_calc), :515-712 (single_search_thin).
# ---------------------------------------------------------------------------
def two_curve_map(CS, tau, fd, eta1, edges1, eta2, edges2):
    """ththmod.py:1557-1636.  Returns (thth_red, edges_red1, edges_red2)."""
    tau = np.asarray(tau, dtype=np.float64)
    fd = np.asarray(fd, dtype=np.float64)
    edges1 = np.asarray(edges1, dtype=np.float64)
    edges2 = np.asarray(edges2, dtype=np.float64)
    c1 = (edges1[1:] + edges1[:-1]) / 2
    c2 = (edges2[1:] + edges2[:-1]) / 2
    th1 = np.ones((c2.shape[0], c1.shape[0])) * c1
    th2 = np.ones((c2.shape[0], c1.shape[0])) * c2[:, np.newaxis]
    dtau = np.diff(tau).mean()
    dfd = np.diff(fd).mean()
    tau_inv = (((eta1 * th1 ** 2 - eta2 * th2 ** 2) - tau[1] + dtau / 2)
               // dtau).astype(int)
    fd_inv = (((th1 - th2) - fd[1] + dfd / 2) // dfd).astype(int)
    thth = np.zeros(tau_inv.shape, dtype=complex)
    pnts = (tau_inv > 0) * (tau_inv < tau.shape[0] - 1) * \
        (fd_inv < fd.shape[0] - 1)
    thth[pnts] = np.asarray(CS)[tau_inv[pnts], fd_inv[pnts]]
    thth *= np.sqrt(np.abs(2 * eta1 * th1 - 2 * eta2 * th2))
    th2_max = np.sqrt(tau.max() / eta2)
    th1_max = np.sqrt(tau.max() / eta1)
    p1 = np.abs(c1) < th1_max
    p2 = np.abs(c2) < th2_max
    er1 = np.zeros(p1.sum() + 1)
    er1[:-1] = edges1[:-1][p1]
    er1[-1] = edges1[1:][p1].max()
    er2 = np.zeros(p2.sum() + 1)
    er2[:-1] = edges2[:-1][p2]
    er2[-1] = edges2[1:][p2].max()
    return thth[p2, :][:, p1], er1, er2


def singularvalue_calc(CS, tau, fd, eta, edges, etaArclet, edgesArclet,
                       centerCut):
    """ththmod.py:496-512: largest singular value of the two-curvature map."""
    red, er1, _ = two_curve_map(CS, tau, fd, eta, edges, etaArclet, edgesArclet)
    cents1 = (er1[1:] + er1[:-1]) / 2
    red[:, np.abs(cents1) < centerCut] = 0
    return np.linalg.svd(red, compute_uv=False)[0]


def thin_sweep(CS, tau, fd, etas, edges, edgesArclet, centerCut):
    """Eta loop of single_search_thin (ththmod.py:589-627): NaN on failure."""
    out = np.zeros(len(etas))
    for i, eta in enumerate(etas):
        try:
            out[i] = singularvalue_calc(CS, tau, fd, eta, edges, eta,
                                        edgesArclet, centerCut)
        except Exception:
            out[i] = np.nan
    return out


def single_search_thin(dspec2, freq, time, etas, edges, edgesArclet, centerCut,
                       fw=0.1, npad=3, coher=True):
    """ththmod.py:515-712 without plotting (incoherent uses |CS|**2, :609)."""
    fd = fft_axis(time, "mHz", npad)
    tau = fft_axis(freq, "us", npad)
    CS = conjugate_spectrum(dspec2, npad, None)
    src = CS if coher else np.abs(CS) ** 2
    eigs = thin_sweep(src, tau, fd, etas, edges, edgesArclet, centerCut)
    eta_fit, eta_sig, _ = peak_fit(etas, eigs, fw)
    return eta_fit, eta_sig, np.mean(freq), np.mean(time), eigs


# ---------------------------------------------------------------------------
# phase retrieval: inverse map, rank-1 model, wavefield of one chunk
# ---------------------------------------------------------------------------
def rev_map(thth, tau, fd, eta, edges, hermetian=True):
    """ththmod.py:176-258: bin the theta-theta points back into the conjugate
    spectrum (np.histogram2d with explicit edges; 1/sqrt|2 eta dtheta| weights;
    bins with a zero-Jacobian (diagonal) point come out NaN -> 0)."""
    tau = np.asarray(tau, dtype=np.float64)
    fd = np.asarray(fd, dtype=np.float64)
    th = theta_centres(edges)
    fd_map = th[np.newaxis, :] - th[:, np.newaxis]
    tau_map = eta * (th[np.newaxis, :] ** 2 - th[:, np.newaxis] ** 2)
    fd_edges = (np.linspace(0, fd.shape[0], fd.shape[0] + 1) - .5) * (fd[1] - fd[0]) + fd[0]
    tau_edges = (np.linspace(0, tau.shape[0], tau.shape[0] + 1) - .5) * (tau[1] - tau[0]) + tau[0]
    bins = (fd_edges, tau_edges)
    with np.errstate(divide="ignore", invalid="ignore"):
        wts = np.ravel(thth / np.sqrt(np.abs(2 * eta * fd_map.T)))
        x, y = np.ravel(fd_map), np.ravel(tau_map)
        recov = (np.histogram2d(x, y, bins=bins, weights=wts.real)[0]
                 + np.histogram2d(x, y, bins=bins, weights=wts.imag)[0] * 1j)
        norm = np.histogram2d(x, y, bins=bins)[0]
        if hermetian:
            recov += (np.histogram2d(-x, -y, bins=bins, weights=wts.real)[0]
                      - np.histogram2d(-x, -y, bins=bins, weights=wts.imag)[0] * 1j)
            norm += np.histogram2d(-x, -y, bins=bins)[0]
        recov /= norm
        recov = np.nan_to_num(recov)
    return recov.T


def modeler(CS, tau, fd, eta, edges):
    """ththmod.py:261-327, hermetian=True branch (the other branch raises
    IndexError in the reference).  The eigenvector's phase
    is arbitrary (ARPACK start vector)."""
    thth_red, edges_red = thth_redmap(CS, tau, fd, eta, edges)
    w, V = eigsh(thth_red, 1, which="LA")
    w, V = w[0], V[:, 0]
    thth2_red = np.outer(V, np.conjugate(V)) * np.abs(w)
    recov = rev_map(thth2_red, tau, fd, eta, edges_red, hermetian=True)
    model = np.fft.ifft2(np.fft.ifftshift(recov)).real
    return thth_red, thth2_red, recov, model, edges_red, w, V


def single_chunk_retrieval(dspec2, edges, time, freq, eta, npad, tau_mask=0.0):
    """ththmod.py:1390-1476: wavefield of one chunk (zeros if anything fails)."""
    fd = fft_axis(time, "mHz", npad)
    tau = fft_axis(freq, "us", npad)
    CS = conjugate_spectrum(dspec2, npad, None, tau, tau_mask)
    try:
        thth_red, _, _, _, edges_red, w, V = modeler(CS, tau, fd, eta, edges)
        ththE = thth_red * 0
        ththE[ththE.shape[0] // 2, :] = np.conjugate(V) * np.sqrt(w)
        recov_E = rev_map(ththE, tau, fd, eta, edges_red, hermetian=False)
        model_E = np.fft.ifft2(np.fft.ifftshift(recov_E))[:dspec2.shape[0], :dspec2.shape[1]]
        model_E = model_E * (dspec2.shape[0] * dspec2.shape[1] / 4)
    except Exception:
        model_E = np.zeros(dspec2.shape, dtype=complex)
    return model_E


def mask_func(w):
    """ththmod.py:1478-1489."""
    x = np.linspace(0, w - 1, w)
    return np.sin((np.pi / 2) * x / w) ** 2


def mosaic(chunks):
    """ththmod.py:1492-1554: phase-align and stack half-overlapping chunks."""
    ncf, nct, cwf, cwt = chunks.shape
    E = np.zeros(((ncf - 1) * (cwf // 2) + cwf, (nct - 1) * (cwt // 2) + cwt), dtype=complex)
    for cf in range(ncf):
        for ct in range(nct):
            new = chunks[cf, ct]
            sl = (slice(cf * cwf // 2, cf * cwf // 2 + cwf),
                  slice(ct * cwt // 2, ct * cwt // 2 + cwt))
            mask = np.ones(new.shape)
            if cf > 0:
                mask[:cwf // 2, :] *= mask_func(cwf // 2)[:, np.newaxis]
            if cf < ncf - 1:
                mask[cwf // 2:, :] *= 1 - mask_func(cwf // 2)[:, np.newaxis]
            if ct > 0:
                mask[:, :cwt // 2] *= mask_func(cwt // 2)
            if ct < nct - 1:
                mask[:, cwt // 2:] *= 1 - mask_func(cwt // 2)
            rot = np.angle((E[sl] * np.conjugate(new) * mask).mean())
            E[sl] += new * mask * np.exp(1j * rot)
    return E


def gerchberg_saxton(wavefield, dyn, freqs, niter=1):
    """dynspec.py:1858-1896 after calc_wavefield: amplitude = sqrt(dyn) where
    dyn is finite and positive, causality (tau < 0 zeroed) in between."""
    W = np.array(wavefield, dtype=complex)
    n0, n1 = W.shape
    d = np.asarray(dyn)[:n0, :n1]
    pos = np.isfinite(d) * (d > 0)
    tau = fft_axis(np.asarray(freqs)[:n0], "us")
    W *= np.sqrt(d[pos].mean() / np.abs(W[pos] ** 2).mean())
    W[pos] = np.sqrt(d[pos]) * np.exp(1j * np.angle(W[pos]))
    for _ in range(niter):
        C = np.fft.fftshift(np.fft.fft2(W))
        C[tau < 0] = 0
        W = np.fft.ifft2(np.fft.ifftshift(C))
        W[pos] = np.sqrt(d[pos]) * np.exp(1j * np.angle(W[pos]))
    return W
