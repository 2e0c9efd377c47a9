"""CPU restatement of Dynspec.calc_sspec / calc_acf (float64 numpy).

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows (reference = /root/reference/scintools):
  get_window   scint_utils.py:810-832
  calc_sspec   dynspec.py:3584-3748  (prewhite / halve / window paths; the
               lamsteps / velocity / trap variants only swap the input array)
  calc_acf     dynspec.py:3750-3814  ('direct' and 'sspec' methods)
"""
import numpy as np
from scipy.signal import convolve2d

_WINDOWS = {"hanning": np.hanning, "hamming": np.hamming,
            "blackman": np.blackman, "bartlett": np.bartlett}


def get_window(nt, nf, window="hanning", frac=0.1):
    """scint_utils.py:810-832. Returns (chan_window[nt], subint_window[nf])."""
    fn = _WINDOWS[window.lower()]
    cw = fn(int(np.floor(frac * nt)))
    sw = fn(int(np.floor(frac * nf)))
    chan = np.insert(cw, int(np.ceil(len(cw) / 2)), np.ones([nt - len(cw)]))
    sub = np.insert(sw, int(np.ceil(len(sw) / 2)), np.ones([nf - len(sw)]))
    return chan, sub


def fft_lengths(nf, nt):
    """dynspec.py:3677-3678."""
    return (int(2 ** (np.ceil(np.log2(nf)) + 1)),
            int(2 ** (np.ceil(np.log2(nt)) + 1)))


def sspec_axes(nf, nt, dt, df, halve=True):
    """dynspec.py:3691-3699: fdop [mHz], tdel [us]."""
    nrfft, ncfft = fft_lengths(nf, nt)
    td = np.arange(0, nrfft // 2 if halve else nrfft)
    fd = np.arange(-ncfft // 2, ncfft // 2)
    fdop = np.multiply(fd, 1e3 / (ncfft * dt))
    tdel = np.divide(td, (nrfft * df))
    return fdop, tdel


def calc_sspec(dyn, dt, df, prewhite=False, halve=True, window="hanning",
               window_frac=0.1, db=True):
    """dynspec.py:3664-3721. Returns (fdop, tdel, sec)."""
    dyn = np.array(dyn, dtype=np.float64)
    nf, nt = dyn.shape
    dyn = dyn - np.mean(dyn)
    if window is not None:
        chan, sub = get_window(nt, nf, window=window, frac=window_frac)
        dyn = chan * dyn
        dyn = (sub * dyn.T).T
    nrfft, ncfft = fft_lengths(nf, nt)
    dyn = dyn - np.mean(dyn)
    if prewhite:
        simpw = convolve2d([[1, -1], [-1, 1]], dyn, mode="valid")
    else:
        simpw = dyn
    simf = np.fft.fft2(simpw, s=[nrfft, ncfft])
    sec = np.fft.fftshift(np.real(simf * np.conj(simf)))
    if halve:
        sec = sec[nrfft // 2:]
    fdop, tdel = sspec_axes(nf, nt, dt, df, halve)
    if prewhite:
        if not halve:
            raise RuntimeError("Cannot apply prewhite to full frame")
        fd = np.arange(-ncfft // 2, ncfft // 2)
        td = np.arange(0, nrfft // 2)
        v1 = np.sin(np.pi / ncfft * fd) ** 2
        v2 = np.sin(np.pi / nrfft * td) ** 2
        postdark = np.outer(v2, v1)
        postdark[:, ncfft // 2] = 1
        postdark[0, :] = 1
        sec = sec / postdark
    if db:
        with np.errstate(divide="ignore"):
            sec = 10 * np.log10(sec)
    return fdop, tdel, sec


def calc_acf(dyn, normalise=True, subtract_mean=True):
    """dynspec.py:3780-3797 (method='direct')."""
    arr = np.array(dyn, dtype=np.float64)
    nf, nt = arr.shape
    if subtract_mean:
        arr = arr - np.mean(arr[np.isfinite(arr)])
    arr = np.fft.fft2(arr, s=[2 * nf, 2 * nt])
    arr = np.abs(arr)
    arr **= 2
    arr = np.real(np.fft.fftshift(np.fft.ifft2(arr)))
    if normalise:
        arr /= np.max(arr)
    return arr


def calc_acf_sspec(dyn, dt, df, normalise=True, window_frac=0.1):
    """dynspec.py:3798-3807 (method='sspec')."""
    _, _, sspec = calc_sspec(dyn, dt, df, prewhite=False, halve=False,
                             window_frac=window_frac)
    sspec = np.fft.fftshift(sspec)
    arr = np.real(np.fft.fftshift(np.fft.fft2(10 ** (sspec / 10))))
    if normalise:
        arr /= np.max(arr)
    return arr


def scale_dyn_lambda(dyn, freqs, spacing="auto"):
    """Dynspec.scale_dyn(scale='lambda') (dynspec.py:3926-3957): resample every
    time column from equal frequency steps to equal wavelength steps with a
    not-a-knot cubic spline (scipy interp1d kind='cubic'), flipped so that
    wavelength increases.  Returns (lamdyn, lam, dlam).
    ORACLE ONLY in round 1: the CUDA row comes next."""
    from scipy.constants import c
    from scipy.interpolate import interp1d
    arin = np.array(dyn, dtype=np.float64)
    nf, nt = arin.shape
    freqs = np.array(freqs, dtype=np.float64)
    lams = np.divide(c, freqs * 10 ** 6)
    adl = np.abs(np.diff(lams))
    if spacing == "auto":
        dlam = (np.max(lams) - np.min(lams)) / len(freqs)
    else:
        dlam = {"max": np.max, "median": np.median, "mean": np.mean, "min": np.min}[spacing](adl)
    lam_eq = np.arange(np.min(lams) + 1e-10, np.max(lams) - 1e-10, dlam)
    feq = np.round(np.divide(c, lam_eq) / 10 ** 6, 6)
    if max(feq) > max(freqs):
        feq[np.argmax(feq)] = max(freqs)
    if min(feq) < min(freqs):
        feq[np.argmin(feq)] = min(freqs)
    # one spline per column; the knots are shared, so this is a single banded
    # solve with nt right-hand sides
    arout = interp1d(freqs, arin, kind="cubic", axis=0)(feq)
    return np.flipud(arout), np.flipud(lam_eq), dlam


def norm_sspec(sspec, fdop, tdel, eta, freq, delmax=None, startbin=1, maxnormfac=5, cutmid=0,
               ref_freq=1400, numsteps=None, weighted=True):
    """Dynspec.norm_sspec (dynspec.py:1920-2183) for lamsteps=False, an explicit
    eta, linear steps, no artefact subtraction / NaN interpolation / spectrum fit:
    every delay row of the secondary spectrum is resampled (np.interp) on the
    normalised Doppler axis fdop / sqrt(tdel / eta).  Returns (normsspec masked
    array, normsspecavg, fdopnew, tdel_cut, powerspectrum).
    ORACLE ONLY in round 1: groundwork for the arc-fit row."""
    sspec = np.array(sspec, dtype=np.float64)
    fdop = np.asarray(fdop, dtype=np.float64)
    yaxis = np.asarray(tdel, dtype=np.float64)
    delmax = np.max(yaxis) if delmax is None else delmax
    c = 299792458.0
    eta = eta / (freq / ref_freq) ** 2 * (c * 1e6 / ((ref_freq * 10 ** 6) ** 2))
    ind = np.argmin(abs(yaxis - delmax))
    sspec = sspec[startbin:ind, :]
    nr, nc = sspec.shape
    sspec[:, int(nc / 2 - np.floor(cutmid / 2)):int(nc / 2 + np.floor(cutmid / 2))] = np.nan
    td = yaxis[startbin:ind]
    maxfdop = maxnormfac * np.sqrt(td[-1] / eta)
    if maxfdop > max(fdop):
        maxfdop = max(fdop)
    nfdop = 2 * len(fdop[abs(fdop) <= maxfdop]) if numsteps is None else numsteps
    if nfdop % 2 != 0:
        nfdop += 1
    fdopnew = np.linspace(-maxnormfac, maxnormfac, nfdop)
    rows, mask = [], []
    for ii in range(len(td)):
        s = np.sqrt(td[ii] / eta)
        sel = abs(fdop) <= maxnormfac * s
        ifdop = fdop[sel] / s
        rows.append(np.interp(fdopnew, ifdop, sspec[ii, sel]))
        mask.append(np.abs(fdopnew) > np.max(np.abs(ifdop)))
    mask = np.array(mask).squeeze()
    norm = np.array(rows).squeeze()
    mask = mask + np.isnan(norm)
    norm = np.ma.array(norm, mask=mask)
    power = np.ma.mean(np.power(10, norm / 10), axis=1)
    xdata = np.sqrt(td)
    ydata = np.sqrt(td) * power
    xdata = xdata[~np.isnan(xdata)]
    ydata = ydata[~np.isnan(ydata)]
    alpha = -11 / 3
    index = np.argmin(np.abs(xdata - 10))
    amp = ydata[index] * xdata[index] ** -alpha
    arc = amp * xdata ** alpha
    weights = 10 * np.log10(arc) if weighted else np.ones(np.shape(arc))
    avg = np.ma.average(norm, axis=0, weights=np.squeeze(weights)).squeeze()
    return norm, avg, fdopnew, td, power
