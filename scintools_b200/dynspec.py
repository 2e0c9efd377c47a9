"""B200-native mirror of the FFT / theta-theta part of scintools.dynspec.

Keeps the ``Dynspec`` method names, signatures and attribute side effects of
the reference for the arc-measurement hot path:

  calc_sspec         dynspec.py:3584-3748  -> sb_sspec_f32
  calc_acf           dynspec.py:3750-3814  -> sb_acf_f32 (+ sspec route)
  prep_thetatheta    dynspec.py:1348-1537  (host bookkeeping, unit free)
  thetatheta_single  dynspec.py:1539-1655  -> sb_cs_f32 + sb_eta_sweep
  fit_thetatheta     dynspec.py:1657-1763  -> per chunk ththmod.single_search

Everything outside that path (file I/O, cleaning, plotting, arc fitting,
lmfit models) is deliberately not here: use the reference for those and hand
the arrays over with ``BasicDyn`` exactly as the reference's tutorials do.
Units: times in s, freqs in MHz, eta in s^3, edges in mHz, tau in us.

Also here: scale_dyn('lambda') (dynspec.py:3926-3957 -> sb_scale_dyn_lambda_f32),
thetatheta_chunks / calc_wavefield / gerchberg_saxton (:1765-1896), and, through
``arcfit.ArcFitMixin``, norm_sspec / fit_arc (:1920-2183, :970-1346).

Provenance: ``prep_thetatheta`` is the reference's dynspec.py:1348-1537 with the
astropy units stripped, line for line (host scalar bookkeeping that SURVEY.md a11
keeps in Python; identical attributes are the contract).  The chunk loops of
``fit_thetatheta`` / ``thetatheta_chunks`` follow :1680-1712 / :1790-1828 the same
way.  Restated reference glue, not new design.
"""
from copy import deepcopy as cp

import numpy as np

from . import _device as D
from . import _lib
from . import ththmod as thth
from . import units as U
from .arcfit import ArcFitMixin

_WINDOWS = {"hanning": np.hanning, "hamming": np.hamming,
            "blackman": np.blackman, "bartlett": np.bartlett}


def get_window(nt, nf, window="hanning", frac=0.1):
    """Edge taper (scint_utils.py:810-832): returns (chan_window[nt],
    subint_window[nf]) -- the first one runs along the TIME axis."""
    try:
        fn = _WINDOWS[window.lower()]
    except KeyError:
        raise ValueError("Window unknown.. Please add it!")
    cw = fn(int(np.floor(frac * nt)))
    sw = fn(int(np.floor(frac * nf)))
    chan_window = np.insert(cw, int(np.ceil(len(cw) / 2)),
                            np.ones([nt - len(cw)]))
    subint_window = np.insert(sw, int(np.ceil(len(sw) / 2)),
                              np.ones([nf - len(sw)]))
    return chan_window, subint_window


def is_valid(array):
    """scint_utils.py:87-91."""
    return np.isfinite(array) * (~np.isnan(array))


class BasicDyn:
    """Container with the attributes Dynspec.load_dyn_obj reads
    (dynspec.py:4146-4230)."""

    def __init__(self, dyn, name="BasicDyn", header=["BasicDyn"], times=[],
                 freqs=[], nchan=None, nsub=None, bw=None, df=None,
                 freq=None, tobs=None, dt=None, mjd=60000):
        times = np.asarray(times)
        freqs = np.asarray(freqs)
        if times.size == 0 or freqs.size == 0:
            raise ValueError("must input array of times and frequencies")
        self.name = name
        self.header = header
        self.times = times
        self.freqs = freqs
        self.nchan = nchan if nchan is not None else len(freqs)
        self.nsub = nsub if nsub is not None else len(times)
        self.bw = bw if bw is not None else abs(max(freqs)) - abs(min(freqs))
        self.df = df if df is not None else freqs[1] - freqs[2]
        self.freq = freq if freq is not None else np.mean(np.unique(freqs))
        self.tobs = tobs
        self.dt = dt if dt is not None else times[1] - times[0]
        self.mjd = mjd
        self.dyn = dyn


class Dynspec(ArcFitMixin):

    def __init__(self, filename=None, dyn=None, verbose=True, process=False,
                 lamsteps=False, remove_short_subs=True, subint_thresh=2.33,
                 mjd=None):
        if filename:
            raise NotImplementedError(
                "psrflux file I/O is outside the B200 hot path: load with "
                "scintools.Dynspec and pass the object as dyn=")
        elif dyn is not None:
            self.load_dyn_obj(dyn, verbose=verbose, process=process,
                              lamsteps=lamsteps)
        else:
            print("Error: No dynamic spectrum file or object")

    def load_dyn_obj(self, dyn, verbose=True, process=True, lamsteps=False):
        """dynspec.py:378-420."""
        self.name = dyn.name
        self.header = dyn.header
        self.times = dyn.times
        self.freqs = dyn.freqs
        self.nchan = dyn.nchan
        self.nsub = dyn.nsub
        self.bw = dyn.bw
        self.df = dyn.df
        self.freq = dyn.freq
        self.dt = dyn.dt
        self.tobs = dyn.tobs if getattr(dyn, "tobs", None) is not None else \
            np.ptp(self.times) + self.dt
        self.mjd = dyn.mjd if getattr(dyn, "mjd", None) is not None else 60000.0
        self.dyn = dyn.dyn
        self.lamsteps = lamsteps
        for extra in ("eta", "betaeta"):      # Simulation objects carry these
            if hasattr(dyn, extra):
                setattr(self, extra, getattr(dyn, extra))
        if process:
            self.calc_acf()
            self.calc_sspec(lamsteps=lamsteps)
        if verbose:
            print("LOADED DYNSPEC OBJECT {0}".format(self.name))

    # ------------------------------------------------------------------
    # secondary spectrum
    # ------------------------------------------------------------------
    def _pick_dyn(self, lamsteps, velocity, trap):
        """The array calc_sspec transforms (reference dynspec.py:3642-3663): the
        wavelength-rescaled copy is made on demand (scale_dyn -> self.lamdyn);
        velocity / trapezoid resampling is outside the B200 hot path and must have
        been set by the caller."""
        if lamsteps and not velocity and not hasattr(self, "lamdyn"):
            self.scale_dyn()
        for flag, attr in ((lamsteps and velocity, "vlamdyn"),
                           (lamsteps, "lamdyn"), (velocity, "vdyn"),
                           (trap, "trapdyn")):
            if flag:
                if not hasattr(self, attr):
                    raise NotImplementedError(
                        "velocity / trapezoid resampling is outside the B200 hot "
                        "path; set self.%s first" % attr)
                return cp(getattr(self, attr))
        return self.dyn

    def calc_sspec(self, prewhite=False, halve=True, plot=False,
                   lamsteps=False, input_dyn=None, input_x=None, input_y=None,
                   trap=False, window='hanning', window_frac=0.1,
                   return_sspec=False, velocity=False, dtype=np.float64):
        """Secondary spectrum (reference dynspec.py:3584-3748).

        Same arguments and side effects (sets self.sspec / fdop / tdel [/ beta],
        or returns (fdop, yaxis, sec)).  The array comes back as float64 like
        the reference's; pass dtype=np.float32 to skip the widening."""
        import torch
        if plot:
            raise NotImplementedError("plotting is outside the B200 hot path")
        dyn = self._pick_dyn(lamsteps, velocity, trap) if input_dyn is None \
            else input_dyn
        dyn = np.asarray(dyn)
        nf, nt = dyn.shape
        if prewhite and not halve:
            raise RuntimeError('Cannot apply prewhite to full frame')
        nrfft = int(2 ** (np.ceil(np.log2(nf)) + 1))
        ncfft = int(2 ** (np.ceil(np.log2(nt)) + 1))
        d = D.upload_f32(dyn)
        wt = wf = None
        swt = swf = 0.0
        if window is not None:
            chan_window, subint_window = get_window(nt, nf, window=window,
                                                    frac=window_frac)
            swt, swf = float(chan_window.sum()), float(subint_window.sum())
            wt = D.upload(chan_window.astype(np.float32))
            wf = D.upload(subint_window.astype(np.float32))
        if halve:
            td = np.array(list(range(0, int(nrfft / 2))))
        else:
            td = np.array(list(range(0, int(nrfft))))
        fd = np.array(list(range(int(-ncfft / 2), int(ncfft / 2))))
        fdop = np.reshape(np.multiply(fd, 1e3 / (ncfft * self.dt)), [len(fd)])
        tdel = np.reshape(np.divide(td, (nrfft * self.df)), [len(td)])
        pd1 = pd2 = None
        if prewhite:   # post-darken vectors, dynspec.py:3706-3711
            pd1 = D.upload(np.power(np.sin(np.multiply(np.pi / ncfft, fd)), 2)
                           .astype(np.float32))
            pd2 = D.upload(np.power(np.sin(np.multiply(np.pi / nrfft, td)), 2)
                           .astype(np.float32))
        sec = D.empty((len(td), ncfft), torch.float32)
        _lib.check(_lib.lib.sb_sspec_f32(
            d.data_ptr(), nf, nt, D.ptr(wt), D.ptr(wf), swt, swf,
            1 if prewhite else 0, 1 if halve else 0, 1, D.ptr(pd1), D.ptr(pd2),
            sec.data_ptr(), D.stream_ptr()))
        sec = D.download(sec, dtype)
        np.seterr(divide='ignore')      # reference side effect (:3720)
        beta = None
        if lamsteps:
            beta = np.divide(td, (nrfft * self.dlam))
        if input_dyn is None and not return_sspec:
            if lamsteps:
                if velocity:
                    self.vlamsspec = sec
                else:
                    self.lamsspec = sec
            elif velocity:
                self.vsspec = sec
            elif trap:
                self.trapsspec = sec
            else:
                self.sspec = sec
            self.fdop = fdop
            self.tdel = tdel
            if lamsteps:
                self.beta = beta
        else:
            return fdop, (beta if lamsteps else tdel), sec

    # ------------------------------------------------------------------
    # autocovariance
    # ------------------------------------------------------------------
    # ------------------------------------------------------------------
    # wavelength rescaling (csrc/scale_dyn.cu)
    # ------------------------------------------------------------------
    @staticmethod
    def _spline_tables(x, xq):
        """Column-independent tables of the not-a-knot cubic spline through the
        ascending knots ``x`` evaluated at ``xq`` (= scipy interp1d(kind='cubic'),
        checked on the CPU to 1e-16): Thomas factors of the second-derivative
        system and the four weights of (y_i, y_i+1, M_i, M_i+1) per query."""
        x = np.asarray(x, dtype=np.float64)
        xq = np.asarray(xq, dtype=np.float64)
        n = x.shape[0]
        h = np.diff(x)
        a = np.zeros(n)
        b = np.zeros(n)
        c = np.zeros(n)
        a[1:n - 1] = h[:-1]
        b[1:n - 1] = 2 * (h[:-1] + h[1:])
        c[1:n - 1] = h[1:]
        p0 = h[0] / h[1]
        pn = h[n - 2] / h[n - 3]
        b[1] += h[0] * (1 + p0)
        c[1] -= h[0] * p0
        a[1] = 0.0
        b[n - 2] += h[n - 2] * (1 + pn)
        a[n - 2] -= h[n - 2] * pn
        c[n - 2] = 0.0
        cpr = np.zeros(n)
        inv = np.zeros(n)
        for i in range(1, n - 1):
            inv[i] = 1.0 / (b[i] - a[i] * cpr[i - 1])
            cpr[i] = c[i] * inv[i]
        g = np.zeros(n)
        g[:n - 1] = 6.0 / h
        idx = np.clip(np.searchsorted(x, xq, side="right") - 1, 0, n - 2)
        hi = h[idx]
        dl = x[idx + 1] - xq
        dr = xq - x[idx]
        W = np.stack([dl / hi, dr / hi, (dl ** 3 / hi - hi * dl) / 6,
                      (dr ** 3 / hi - hi * dr) / 6], axis=1)
        return dict(a=a, cp=cpr, inv=inv, g=g, p0=p0, pn=pn, idx=idx, W=W)

    def scale_dyn(self, scale='lambda', spacing='auto', **kwargs):
        """Resample the dynamic spectrum to equal wavelength steps (reference
        dynspec.py:3872-3957, scale='lambda' only) -> self.lamdyn, self.lam,
        self.nlam, self.dlam."""
        import torch
        from scipy.constants import c as c_light
        if not (('lambda' in scale) or ('wavelength' in scale)):
            raise NotImplementedError("only scale='lambda' is on the B200 path")
        freqs = np.array(self.freqs, dtype=np.float64)
        nf, nt = self.dyn.shape
        lams = np.divide(c_light, freqs * 10 ** 6)
        adl = np.abs(np.diff(lams))
        if spacing == 'auto':
            dlam = (np.max(lams) - np.min(lams)) / len(freqs)
        else:
            dlam = {'max': np.max, 'median': np.median, 'mean': np.mean,
                    'min': np.min}[spacing](adl)
        lam_eq = np.arange(np.min(lams) + 1e-10, np.max(lams) - 1e-10, dlam)
        feq = np.round(np.divide(c_light, lam_eq) / 10 ** 6, 6)
        if max(feq) > max(freqs):
            feq[np.argmax(feq)] = max(freqs)
        if min(feq) < min(freqs):
            feq[np.argmin(feq)] = min(freqs)
        d = np.diff(freqs)
        if np.all(d > 0):
            flip, x = 0, freqs
        elif np.all(d < 0):
            flip, x = 1, freqs[::-1]
        else:
            raise ValueError("scale_dyn needs a monotonic frequency axis")
        T = self._spline_tables(x, feq)
        f32 = lambda v: D.upload(np.ascontiguousarray(v, dtype=np.float32))
        dd = D.upload_f32(np.asarray(self.dyn))
        nlam = feq.shape[0]
        out = D.empty((nlam, nt), torch.float32)
        a, cpr, inv, g = f32(T["a"]), f32(T["cp"]), f32(T["inv"]), f32(T["g"])
        idx = D.upload(np.ascontiguousarray(T["idx"], dtype=np.int32))
        W = f32(T["W"])
        _lib.check(_lib.lib.sb_scale_dyn_lambda_f32(
            dd.data_ptr(), nf, nt, flip, a.data_ptr(), cpr.data_ptr(), inv.data_ptr(),
            g.data_ptr(), float(T["p0"]), float(T["pn"]), idx.data_ptr(), W.data_ptr(), nlam,
            out.data_ptr(), D.stream_ptr()))
        self.dlam = dlam
        self.lamdyn = out.cpu().numpy().astype(np.float64)
        self.lam = np.flipud(lam_eq)
        self.nlam = len(self.lam)

    def calc_acf(self, method='direct', input_dyn=None, normalise=True,
                 window_frac=0.1, dtype=np.float64):
        """Autocovariance function (reference dynspec.py:3750-3814)."""
        import torch
        if method == 'direct':
            src = np.asarray(self.dyn if input_dyn is None else input_dyn)
            nf, nt = src.shape
            d = D.upload_f32(src)
            out = D.empty((2 * nf, 2 * nt), torch.float32)
            _lib.check(_lib.lib.sb_acf_f32(
                d.data_ptr(), nf, nt, 1 if input_dyn is None else 0,
                1 if normalise else 0, out.data_ptr(), D.stream_ptr()))
            arr = D.download(out, dtype)
        elif method == 'sspec':     # FFT of the secondary spectrum, :3798-3807
            src = np.asarray(self.dyn)
            nf, nt = src.shape
            nrfft = int(2 ** (np.ceil(np.log2(nf)) + 1))
            ncfft = int(2 ** (np.ceil(np.log2(nt)) + 1))
            cw, sw = get_window(nt, nf, window='hanning', frac=window_frac)
            d = D.upload_f32(src)
            wt = D.upload(cw.astype(np.float32))
            wf = D.upload(sw.astype(np.float32))
            out = D.empty((nrfft, ncfft), torch.float32)
            _lib.check(_lib.lib.sb_acf_sspec_f32(
                d.data_ptr(), nf, nt, wt.data_ptr(), wf.data_ptr(),
                float(cw.sum()), float(sw.sum()), 1 if normalise else 0,
                out.data_ptr(), D.stream_ptr()))
            arr = D.download(out, dtype)
        else:
            print('Method not understood. Choose "direct" or "sspec"')
            return
        if input_dyn is None:
            self.acf = arr
        else:
            return arr

    # ------------------------------------------------------------------
    # theta-theta
    # ------------------------------------------------------------------
    def prep_thetatheta(self, fw=.1, npad=3, verbose=False,
                        fitting_proc='standard', **kwargs):
        """Set up the theta-theta search (reference dynspec.py:1348-1537).

        Recognises cwf, cwt, fref, eta_min, eta_max, nedge, edges_lim, tau_lim,
        tau_mask and, for fitting_proc='thin', arclet_lim and center_cut."""
        fitting_procs = ['standard', 'thin', 'incoherent']
        assert fitting_proc in fitting_procs, \
            f'fitting_proc must be one of {fitting_procs}'
        self.thetatheta_proc = fitting_proc
        self.npad = npad
        self.fw = fw
        if 'cwf' in kwargs:
            self.cwf = 2 * (kwargs['cwf'] // 2)
            self.ncf_fit = self.dyn.shape[0] // self.cwf
            self.ncf_ret = (self.dyn.shape[0] // (self.cwf // 2)) - 1
        else:
            self.cwf = self.dyn.shape[0]
            self.ncf_fit = self.ncf_ret = 1
        if 'cwt' in kwargs:
            self.cwt = 2 * (kwargs['cwt'] // 2)
            self.nct_fit = self.dyn.shape[1] // self.cwt
            self.nct_ret = (self.dyn.shape[1] // (self.cwt // 2)) - 1
        else:
            self.cwt = self.dyn.shape[1]
            self.nct_fit = self.nct_ret = 1
        tau_lim = float(U.value(kwargs['tau_lim'], "us")) \
            if 'tau_lim' in kwargs else None
        self.fref = float(U.value(kwargs['fref'], "MHz")) \
            if 'fref' in kwargs else float(np.mean(self.freqs))

        fd = U.value(thth.fft_axis(self.times[:self.cwt], "mHz"), "mHz")
        tau = U.value(thth.fft_axis(self.freqs[:self.cwf], "us"), "us")
        eta_min = 4 * (tau[1] - tau[0]) / fd.max() ** 2
        eta_max = tau.max() / (fd[1] - fd[0]) ** 2
        eta_min *= (np.max(self.freqs) / self.fref) ** 2
        eta_max *= (np.min(self.freqs) / self.fref) ** 2
        if 'eta_min' in kwargs:
            eta_min = max((float(U.value(kwargs['eta_min'], "s3")), eta_min))
        if 'eta_max' in kwargs:
            eta_max = min((float(U.value(kwargs['eta_max'], "s3")), eta_max))
        if not ('eta_min' in kwargs and 'eta_max' in kwargs):
            c = 299792458.0
            if not hasattr(self, "betaeta"):
                # Hough prior (reference dynspec.py:1458-1466): eta [s^3] -> betaeta
                # [1/(m mHz^2)] = eta * fref^2 [MHz^2 = 1e12 s^-2] / c / (1e6 s^2/mHz^-2)
                to_beta = self.fref ** 2 * 1e12 / c / 1e6
                self.fit_arc(lamsteps=True, numsteps=1e4, etamin=eta_min * to_beta,
                             etamax=eta_max * to_beta, delmax=tau_lim, plot=False)
            eta_hough = c * self.betaeta / self.fref ** 2 * 1e-12 * 1e6
            err_hough = c * 2 * max((self.betaetaerr, self.betaetaerr2)) \
                / self.fref ** 2 * 1e-12 * 1e6
            if 'eta_min' not in kwargs:
                eta_min = max((eta_min, eta_hough - err_hough))
            if 'eta_max' not in kwargs:
                eta_max = min((eta_max, eta_hough + err_hough))
        self.eta_min, self.eta_max = float(eta_min), float(eta_max)
        l0 = np.log10(self.eta_min)
        l1 = np.log10(self.eta_max)
        self.neta = int(1 + (l1 - l0) / np.log10(1 + self.fw / 10))

        if self.thetatheta_proc == 'thin':
            fd_cut = fd.max() * (self.fref / np.max(self.freqs))
        else:
            fd_cut = (fd.max() / 2) * (self.fref / np.max(self.freqs))
        if 'edges_lim' in kwargs:
            edges_lim = min((float(U.value(kwargs['edges_lim'], "mHz")), fd_cut))
        else:
            edges_lim = fd_cut
        if tau_lim is not None:
            edges_lim = min((edges_lim, np.sqrt(tau_lim / self.eta_max)))
        if 'nedge' in kwargs:
            assert np.mod(kwargs['nedge'], 2) == 0, 'nedge must be even!'
            self.edges = np.linspace(-edges_lim, edges_lim, kwargs['nedge'])
        else:
            self.edges = U.value(thth.min_edges(
                edges_lim, fd, tau,
                self.eta_max * (self.fref / np.min(self.freqs)), 2), "mHz") \
                * (np.min(self.freqs) / self.fref)
        if self.thetatheta_proc == 'thin':
            self.arclet_lim = float(U.value(kwargs['arclet_lim'], "mHz")) \
                if 'arclet_lim' in kwargs else float(edges_lim)
            self.center_cut = float(U.value(kwargs['center_cut'], "mHz")) \
                if 'center_cut' in kwargs else 0.0
        self.thth_tau_mask = float(U.value(kwargs['tau_mask'], "us")) \
            if 'tau_mask' in kwargs else 0.0
        if verbose:
            print("\n\t THETA-THETA PROPERTIES\n")
            print(f'Channels per chunk: {self.cwf}')
            print(f'Time bins per chunk: {self.cwt}')
            print(f'Number of fitting chunks: {self.ncf_fit}x{self.nct_fit}')
            print(f'Reference Frequency: {self.fref}')
            print(f'Eta range: {self.eta_min} to {self.eta_max} '
                  f'with {self.neta} points')
            print(f'Edges has {self.edges.shape[0]} point out to '
                  f'{self.edges[-1]}')
            print(f'Zero paddings: {self.npad}')

    def _chunk_etas(self, fmean):
        return np.logspace(np.log10(self.eta_min), np.log10(self.eta_max),
                           self.neta) * (self.fref / fmean) ** 2

    def thetatheta_single(self, cf=0, ct=0, fname=None, verbose=False,
                          plot=False, arrays=True):
        """Theta-theta on one chunk (reference dynspec.py:1539-1655).
        Returns (etas, eigs, popt) as the reference does with arrays=True."""
        if not hasattr(self, 'cwf'):
            self.prep_thetatheta(verbose=verbose)
        if plot:
            raise NotImplementedError("plotting is outside the B200 hot path")
        cf = min(cf, self.ncf_fit - 1)
        ct = min(ct, self.nct_fit - 1)
        fs = slice(cf * self.cwf, (cf + 1) * self.cwf)
        ts = slice(ct * self.cwt, (ct + 1) * self.cwt)
        time2 = np.asarray(self.times[ts], dtype=np.float64)
        freq2 = np.asarray(self.freqs[fs], dtype=np.float64)
        tau = U.value(thth.fft_axis(freq2, "us", self.npad), "us")
        fd = U.value(thth.fft_axis(time2, "mHz", self.npad), "mHz")
        dspec2 = np.copy(self.dyn[fs, ts]).astype(np.float64)
        dspec2 -= np.nanmean(dspec2)
        etas = self._chunk_etas(freq2.mean())
        edges = self.edges * (freq2.mean() / self.fref)
        if self.thetatheta_proc == 'thin':
            # dynspec.py:1593-1600: one singularvalue_calc per curvature
            cs = thth.conjugate_spectrum(np.nan_to_num(dspec2), self.npad, 0.0,
                                         tau, self.thth_tau_mask)
            eigs = thth.thin_sweep(cs, tau, fd, etas, edges,
                                   edges[np.abs(edges) < self.arclet_lim],
                                   self.center_cut)
        else:
            cs = thth.conjugate_spectrum(
                np.nan_to_num(dspec2), self.npad, 0.0, tau, self.thth_tau_mask,
                ncols_keep=thth.needed_fd_columns(fd, edges))
            eigs = thth.eta_sweep(cs, tau, fd, etas, edges,
                                  self.thetatheta_proc == 'standard')
        if not np.all(np.isfinite(eigs)) and verbose:
            print("some curvatures failed (NaN)")
        eta_fit, eta_sig, popt = thth.peak_fit(etas, eigs, self.fw)
        self.last_eta_fit, self.last_eta_sig = eta_fit, eta_sig
        if arrays:
            good = np.isfinite(eigs)
            return etas[good], eigs[good], popt

    def fit_thetatheta(self, verbose=False, plot=False, pool=None,
                       time_avg=False):
        """Loop theta-theta over all fitting chunks and fit eta ~ nu^-2
        (reference dynspec.py:1657-1763).  ``pool`` must be None: a CUDA
        context does not survive fork; the chunks run back to back on the GPU
        (each one already fills it)."""
        if pool is not None:
            raise ValueError("fit_thetatheta on the B200 path takes pool=None "
                             "(CUDA is not fork-safe); chunks are batched on "
                             "the device instead")
        if plot:
            raise NotImplementedError("plotting is outside the B200 hot path")
        if not hasattr(self, 'cwf'):
            self.prep_thetatheta(verbose=verbose)
        self.eta_evo = np.zeros((self.ncf_fit, self.nct_fit))
        self.eta_evo_err = np.zeros((self.ncf_fit, self.nct_fit))
        self.f0s = np.zeros(self.ncf_fit)
        self.t0s = np.zeros(self.nct_fit)
        coher = (self.thetatheta_proc != 'incoherent')
        pars, where = [], []
        for cf in range(self.ncf_fit):
            fs = slice(cf * self.cwf, (cf + 1) * self.cwf)
            freq2 = np.copy(self.freqs[fs]).astype(np.float64)
            self.f0s[cf] = freq2.mean()
            etas = self._chunk_etas(freq2.mean())
            for ct in range(self.nct_fit):
                ts = slice(ct * self.cwt, (ct + 1) * self.cwt)
                time2 = np.copy(self.times[ts]).astype(np.float64)
                dspec2 = np.copy(self.dyn[fs, ts]).astype(np.float64)
                dspec2 -= np.nanmean(dspec2)
                dspec2 = np.nan_to_num(dspec2)
                scale = freq2.mean() / self.fref
                params = [dspec2, freq2, time2, etas, self.edges * scale, None,
                          False, self.fw, self.npad, coher]
                if self.thetatheta_proc == 'thin':
                    params += [verbose,
                               self.edges[np.abs(self.edges) < self.arclet_lim]
                               * scale, self.center_cut]
                else:
                    params += [self.thth_tau_mask, verbose]
                pars.append(params)
                where.append((cf, ct))
                self.t0s[ct] = time2.mean()
        # the reference's pool.map over the chunks (dynspec.py:1715-1719): here the
        # chunks run back to back on the GPU
        if self.thetatheta_proc == 'thin':
            results = [thth.single_search_thin(p) for p in pars]
        else:
            results = thth.search_batch(pars)
        for (cf, ct), res in zip(where, results):
            self.eta_evo[cf, ct] = U.value(res[0], "s3")
            self.eta_evo_err[cf, ct] = U.value(res[1], "s3")
        f0 = self.f0s[:, np.newaxis]
        if time_avg:
            eta_avg = np.nanmean(self.eta_evo, 1)
            eta_count = np.nansum(self.eta_evo, 1) / eta_avg
            avg_err = np.nanstd(self.eta_evo, 1) / np.sqrt(eta_count - 1)
            tofit = np.isfinite(eta_avg) * np.isfinite(avg_err)
            A = (np.sum(eta_avg[tofit] / (self.f0s * avg_err)[tofit] ** 2) /
                 np.sum(1 / (self.f0s ** 2 * avg_err)[tofit] ** 2))
            A_err = np.sqrt(1 / np.sum(2 / ((self.f0s ** 2) * avg_err)[tofit] ** 2))
        else:
            tofit = np.isfinite(self.eta_evo) * np.isfinite(self.eta_evo_err)
            A = (np.sum(self.eta_evo[tofit] / (f0 * self.eta_evo_err)[tofit] ** 2) /
                 np.sum(1 / ((f0 ** 2) * self.eta_evo_err)[tofit] ** 2))
            A_err = np.sqrt(1 / np.sum(2 / ((f0 ** 2) * self.eta_evo_err)[tofit] ** 2))
        self.ththeta = A / self.fref ** 2
        self.ththetaerr = A_err / self.fref ** 2

    def thetatheta_chunks(self, verbose=False, pool=None, memmap=False, group=None):
        """Phase retrieval on every half-overlapping retrieval chunk
        (reference dynspec.py:1765-1828) -> self.chunks [ncf_ret][nct_ret][cwf][cwt].
        ``pool`` must be None (see fit_thetatheta); memmap is not supported.
        Under torch.distributed (one process per GPU) the chunks are
        block-partitioned over the ranks of ``group`` and all-gathered, the
        counterpart of the reference's ``pool.map`` over chunks (:1815-1828)."""
        if pool is not None:
            raise ValueError("thetatheta_chunks on the B200 path takes pool=None "
                             "(CUDA is not fork-safe)")
        if memmap:
            raise NotImplementedError("memmap chunk storage is outside the B200 path")
        if not hasattr(self, "ththeta"):
            self.fit_thetatheta(verbose=verbose)
        from . import sharding
        pars = []
        for cf in range(self.ncf_ret):
            fs = slice(cf * (self.cwf // 2), cf * (self.cwf // 2) + self.cwf)
            freq2 = np.copy(self.freqs[fs]).astype(np.float64)
            freq = freq2.mean()
            eta = self.ththeta * (self.fref / freq) ** 2
            for ct in range(self.nct_ret):
                ts = slice(ct * (self.cwt // 2), ct * (self.cwt // 2) + self.cwt)
                time2 = np.copy(self.times[ts]).astype(np.float64)
                dspec2 = np.copy(self.dyn[fs, ts]).astype(np.float64)
                dspec2 -= np.nanmean(dspec2)
                dspec2 = np.nan_to_num(dspec2)
                pars.append((dspec2, self.edges * (freq / self.fref), time2, freq2, eta, ct, cf,
                             self.npad, self.thth_tau_mask, verbose))

        def one(p):
            e = thth.single_chunk_retrieval(p)[0]
            return np.concatenate((e.real.ravel(), e.imag.ravel()))

        L = self.cwf * self.cwt
        flat = sharding.sharded_map(one, pars, 2 * L, group)
        self.chunks = (flat[:, :L] + 1j * flat[:, L:]).reshape(
            self.ncf_ret, self.nct_ret, self.cwf, self.cwt)

    def calc_wavefield(self, verbose=False, pool=None, gs=False, memmap=False,
                       niter=1):
        """Mosaic the retrieved chunks into self.wavefield (reference
        dynspec.py:1830-1856); gs=True refines it with ``niter``
        Gerchberg-Saxton iterations."""
        if not hasattr(self, "chunks"):
            self.thetatheta_chunks(verbose=verbose, pool=pool, memmap=memmap)
        self.wavefield = thth.mosaic(self.chunks)
        if gs:
            self.gerchberg_saxton(verbose=verbose, pool=pool, niter=niter)

    def gerchberg_saxton(self, niter=1, verbose=False, pool=None):
        """Gerchberg-Saxton refinement of self.wavefield: measured amplitude
        where the dynamic spectrum is finite and positive, causality
        (tau < 0 zeroed) in between (reference dynspec.py:1858-1896).  The
        wavefield must have power-of-two sizes."""
        import torch
        from . import _device as D, _lib
        self.calc_wavefield(verbose=verbose, pool=pool)
        n0, n1 = self.wavefield.shape
        d = np.asarray(self.dyn[:n0, :n1], dtype=np.float64)
        pos = np.isfinite(d) * (d > 0)
        tau = U.value(thth.fft_axis(np.asarray(self.freqs[:n0], dtype=np.float64), "us"), "us")
        W = self.wavefield * np.sqrt(d[pos].mean() / np.abs(self.wavefield[pos] ** 2).mean())
        W[pos] = np.sqrt(d[pos]) * np.exp(1j * np.angle(W[pos]))
        if niter > 0:
            amp = np.full((n0, n1), np.nan, dtype=np.float32)
            amp[pos] = np.sqrt(d[pos])
            rowmask = np.fft.ifftshift(tau < 0).astype(np.uint8)     # unshifted row order
            wd, ad, md = D.upload_f32(W), D.upload(amp), D.upload(rowmask)
            _lib.check(_lib.lib.sb_gerchberg_saxton_f32(wd.data_ptr(), ad.data_ptr(),
                                                        md.data_ptr(), n0, n1, int(niter),
                                                        D.stream_ptr()))
            a = wd.cpu().numpy()
            W = a[..., 0].astype(np.float64) + 1j * a[..., 1].astype(np.float64)
        self.wavefield = W
