"""CPU restatement of scint_sim.Simulation (float64 / complex128 numpy).

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows (reference = /root/reference/scintools/scint_sim.py):
  set_constants  :137-167      swdsp       :276-292
  get_screen     :169-207      frfilt3     :294-311
  get_intensity  :209-236      get_dynspec :238-252
  __init__ tail  :81-133  (axes, eta, betaeta)

The noise fields are an explicit input (``noise_re``, ``noise_im``) so the GPU
path and the oracle can be fed identical numbers; ``legacy_noise(seed, nx, ny)``
reproduces the reference's ``random.seed`` + two ``randn`` draws (:173,201-202).
Reference quirks kept on purpose: the ky=0 mirror off-by-one (:185), the
Fresnel filter rounded to complex64 (:297), ``spe`` stored as complex64 (:210),
``xyi`` from the last frequency only (:232).
"""
import numpy as np
from scipy.special import gamma


def legacy_noise(seed, nx, ny):
    """scint_sim.py:173,201-202 -- legacy MT19937 global stream."""
    st = np.random.RandomState(seed)
    a = st.randn(nx, ny)
    b = st.randn(nx, ny)
    return a, b


class SimOracle:
    def __init__(self, mb2=2, rf=1, ds=0.01, alpha=5 / 3, ar=1, psi=0,
                 inner=0.001, ns=256, nf=256, dlam=0.25, lamsteps=False,
                 seed=None, nx=None, ny=None, dx=None, dy=None, freq=1400,
                 dt=30, noise_re=None, noise_im=None):
        self.mb2, self.rf, self.ds = mb2, rf, ds
        self.dx = dx if dx is not None else ds
        self.dy = dy if dy is not None else ds
        self.alpha, self.ar, self.psi, self.inner = alpha, ar, psi, inner
        self.nx = nx if nx is not None else ns
        self.ny = ny if ny is not None else ns
        self.nf, self.dlam, self.lamsteps, self.seed = nf, dlam, lamsteps, seed
        self.set_constants()
        if noise_re is None:
            noise_re, noise_im = legacy_noise(seed, self.nx, self.ny)
        self.get_screen(noise_re, noise_im)
        self.get_intensity()
        if nf > 1:
            self.spi = np.real(self.spe * np.conj(self.spe))
        # physical axes (scint_sim.py:92-121)
        self.dt, self.freq = dt, freq
        dyn = self.spi
        self.nsub, self.nchan = dyn.shape
        if not lamsteps:
            self.df = freq * dlam / (self.nchan - 1)
            self.freqs = freq + np.arange(-self.nchan / 2, self.nchan / 2, 1) \
                * self.df
        else:                                   # scint_sim.py:106-112
            from scipy.constants import c as _c
            lam = _c / (freq * 10 ** 6)
            dl = lam * dlam / (self.nchan - 1)
            lams = lam + np.arange(-self.nchan / 2, self.nchan / 2, 1) * dl
            self.freqs = _c / lams / 10 ** 6
            self.freq = freq = (np.max(self.freqs) - np.min(self.freqs)) / 2
        self.times = dt * np.arange(0, self.nsub)
        self.dyn = dyn.T
        V = ds / dt
        k = 2 * np.pi / freq
        L = rf ** 2 * k
        self.eta = L / (2 * V ** 2) / 10 ** 6 / np.cos(psi * np.pi / 180) ** 2

    def set_constants(self):
        lenx = self.nx * self.dx
        leny = self.ny * self.dy
        self.ffconx = (2.0 / (lenx * lenx)) * (np.pi * self.rf) ** 2
        self.ffcony = (2.0 / (leny * leny)) * (np.pi * self.rf) ** 2
        dqx = 2 * np.pi / lenx
        dqy = 2 * np.pi / leny
        a2 = self.alpha * 0.5
        cmb2 = self.alpha * self.mb2 / (4 * np.pi * gamma(1.0 - a2) *
                                        np.cos(self.alpha * np.pi * 0.25))
        self.consp = cmb2 * dqx * dqy / (self.rf ** self.alpha)

    def swdsp(self, kx, ky):
        cs = np.cos(self.psi * np.pi / 180)
        sn = np.sin(self.psi * np.pi / 180)
        r = self.ar
        con = np.sqrt(self.consp)
        alf = -(self.alpha + 2) / 4
        a = (cs ** 2) / r + r * sn ** 2
        b = r * cs ** 2 + sn ** 2 / r
        c = 2 * cs * sn * (1 / r - r)
        q2 = a * np.power(kx, 2) + b * np.power(ky, 2) + c * np.multiply(kx, ky)
        return con * np.multiply(
            np.power(q2, alf),
            np.exp(-(np.add(np.power(kx, 2), np.power(ky, 2))) *
                   self.inner ** 2 / 2))

    def screen_weights(self):
        nx, ny = self.nx, self.ny
        nx2 = int(nx / 2 + 1)
        ny2 = int(ny / 2 + 1)
        w = np.zeros([nx, ny])
        dqx = 2 * np.pi / (self.dx * nx)
        dqy = 2 * np.pi / (self.dy * ny)
        k = np.arange(2, nx2 + 1)
        w[k - 1, 0] = self.swdsp(kx=(k - 1) * dqx, ky=0)
        w[nx + 1 - k, 0] = w[k, 0]              # reference quirk (:185)
        ll = np.arange(2, ny2 + 1)
        w[0, ll - 1] = self.swdsp(kx=0, ky=(ll - 1) * dqy)
        w[0, ny + 1 - ll] = w[0, ll - 1]
        kp = np.arange(2, nx2 + 1)
        k = np.arange((nx2 + 1), nx + 1)
        km = -(nx - k + 1)
        for il in range(2, ny2 + 1):
            w[kp - 1, il - 1] = self.swdsp(kx=(kp - 1) * dqx, ky=(il - 1) * dqy)
            w[k - 1, il - 1] = self.swdsp(kx=km * dqx, ky=(il - 1) * dqy)
            w[nx + 1 - kp, ny + 1 - il] = w[kp - 1, il - 1]
            w[nx + 1 - k, ny + 1 - il] = w[k - 1, il - 1]
        return w

    def get_screen(self, noise_re, noise_im):
        self.w = self.screen_weights()
        xyp = self.w * (noise_re + 1j * noise_im)
        self.xyp = np.real(np.fft.fft2(xyp))

    def freq_scale(self, ifreq):
        if self.lamsteps:
            return 1.0 + self.dlam * (ifreq - 1 - (self.nf / 2)) / self.nf
        return 1 / (1.0 + self.dlam * (-0.5 + ifreq / self.nf))

    def frfilt3(self, xye, scale):
        nx, ny = self.nx, self.ny
        nx2 = int(nx / 2) + 1
        ny2 = int(ny / 2) + 1
        filt = np.zeros([nx2, ny2], dtype=np.csingle)
        q2x = np.linspace(0, nx2 - 1, nx2) ** 2 * scale * self.ffconx
        for ly in range(0, ny2):
            q2 = q2x + (self.ffcony * (ly ** 2) * scale)
            filt[:, ly] = np.cos(q2) - 1j * np.sin(q2)
        xye[0:nx2, 0:ny2] *= filt
        xye[nx:nx2 - 1:-1, 0:ny2] *= filt[1:(nx2 - 1), 0:ny2]
        xye[0:nx2, ny:ny2 - 1:-1] *= filt[0:nx2, 1:(ny2 - 1)]
        xye[nx:nx2 - 1:-1, ny:ny2 - 1:-1] *= filt[1:(nx2 - 1), 1:(ny2 - 1)]
        return xye

    def get_intensity(self):
        spe = np.zeros([self.nx, self.nf], dtype=np.csingle)
        for ifreq in range(self.nf):
            scale = self.freq_scale(ifreq)
            xye = np.fft.fft2(np.exp(1j * self.xyp * scale))
            xye = self.frfilt3(xye, scale)
            xye = np.fft.ifft2(xye)
            spe[:, ifreq] = xye[:, int(np.floor(self.ny / 2))]
        self.xyi = np.real(xye * np.conj(xye))
        self.spe = spe
