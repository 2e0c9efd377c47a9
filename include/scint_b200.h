/* libscint_b200 -- C ABI of the B200-native scintools arc-measurement hot path.
 *
 * The reference (danielreardon/scintools) is pure Python and has no FFI; each
 * entry point below names the reference callable whose arithmetic it replaces
 * (file:line relative to the reference checkout).  INTEGRATION.md shows the
 * ctypes binding a scintools maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on failure (sb_last_error()
 *    holds the message); nothing throws, nothing is printed;
 *  - pointers are DEVICE pointers unless the parameter name ends in _host;
 *    buffers are caller owned; `stream` is a cudaStream_t passed as void*;
 *  - 2-D arrays are C-contiguous (row-major); complex = interleaved
 *    (re, im) float pairs; dyn is [freq][time] like Dynspec.dyn;
 *  - one CUDA context per process, calls into one device from one host
 *    thread at a time (the library keeps a grow-only scratch workspace per
 *    process; sb_release() frees it).  Not fork-safe (CUDA is not).
 *  - sm_100a only; there is no CPU fallback.
 */
#ifndef SCINT_B200_H
#define SCINT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK 0
#define SB_ERR_CUDA (-1)
#define SB_ERR_ARG (-2)
#define SB_ERR_NOMEM (-3)
#define SB_ERR_UNSUPPORTED (-4)

/* per-eta status bits written by sb_eta_sweep */
#define SB_ETA_OK 0
#define SB_ETA_INDEX_ERROR 1   /* numpy would raise IndexError -> NaN (ththmod.py:795-799) */
#define SB_ETA_ZERO_START 2    /* start row all zero -> NaN v0 -> ARPACK error -> NaN */
#define SB_ETA_TOO_SMALL 4     /* cropped matrix smaller than 3x3 -> eigsh raises -> NaN */
#define SB_ETA_NOT_CONVERGED 8 /* Lanczos hit max_iter; best Ritz value returned */

int sb_abi_version(void);
const char* sb_last_error(void);
/* select device, create the context, query SM count. */
int sb_init(int device);
/* free the scratch workspace. */
int sb_release(void);

/* Per-kernel timing with CUDA events recorded on the launching stream.
 * Slots: 0 cs_rows 1 cs_colA 2 cs_colB 3 thth_prep 4 thth_build 5 thth_eig
 * 6 sspec 7 acf 8 sim_screen 9 sim_freq.  sb_profile_collect synchronises the
 * device, writes accumulated milliseconds and launch counts (host arrays of
 * at least 16 entries) and resets the accumulators. */
/* number of kernels this library has launched so far in this process */
int64_t sb_launch_count(void);
int sb_profile_enable(int32_t on);
int sb_profile_collect(double* ms_host, int32_t* count_host, int32_t n);

/* ---- theta-theta ------------------------------------------------------- */

/* Geometry of a conjugate spectrum + theta grid.  Scalars are the
 * reference's own numpy expressions evaluated by the host layer
 * (scintools/ththmod.py:83-97,153-156):
 *   tau0 = tau[0]; dtau = np.diff(tau).mean(); tau_absmax = abs(tau.max())
 *   fd0  = fd[0];  dfd  = np.diff(fd).mean();  fd_half = abs(fd.max())/2
 *   th_cents = recentred bin centres of `edges` (device, float64, n of them)
 */
typedef struct sb_thth_geom {
    const void* cs;        /* float2 conjugate spectrum, rows fftshifted (see cs_half) */
    int64_t ntau, nfd;     /* logical size: len(tau), len(fd) */
    double tau0, dtau, tau_absmax;
    double fd0, dfd, fd_half;
    const double* th_cents;      /* device */
    const double* th_cents_host; /* same values on the host */
    int32_t n_th;
    int32_t coherent;      /* 1: complex CS; 0: |CS| (ththmod.py:801) */
    int64_t cs_pitch;      /* elements per stored row (nfd for a plain full array) */
    int32_t cs_half;       /* 0: full fftshifted [ntau][nfd].  1: Hermitian half as
                              written by sb_cs_f32(half_plane=1): [ntau][cs_pitch],
                              columns k = 0..nfd/2 are the NON-shifted fd >= 0 bins;
                              the rest follows from CS[-tau,-fd] = conj(CS[tau,fd])
                              (valid for the CS of a real dynamic spectrum) */
    int32_t cs_valid_cols; /* cs_half only: how many of the nfd/2+1 stored columns hold data
                              (sb_cs_f32 with ncols_keep > 0 computes only those the theta
                              grid can reach); 0 = all.  Nothing beyond is ever read. */
    const float* cs_bound; /* device scalar: an upper bound of max |re|, |im| over the CS
                              (sb_cs_bound_f32), or NULL -> the sweep scans the CS itself.
                              Only used to scale the fp16 iteration copy of the matrices. */
} sb_thth_geom;

/* Replaces the eta loop of ththmod.single_search (ththmod.py:789-811) /
 * Dynspec.thetatheta_single (dynspec.py:1587-1600), i.e. neta calls of
 * ththmod.Eval_calc (ththmod.py:371-401 -> thth_redmap :119-173 -> thth_map
 * :56-116 -> scipy eigsh(k=1, which="LA")).
 * etas: device float64[neta].  Outputs (device): eigs float64[neta] (NaN where
 * the reference would have produced NaN), status int32[neta] (SB_ETA_*),
 * nred int32[neta] (size of the cropped matrix), iters int32[neta].
 * tol: relative residual tolerance of the Lanczos solve (<=0 -> 2e-5);
 * max_iter: <=0 -> 256. */
int sb_eta_sweep(const sb_thth_geom* geom, const double* etas, int32_t neta,
                 double tol, int32_t max_iter, double* eigs, int32_t* status,
                 int32_t* nred, int32_t* iters, void* stream);

/* Replaces ththmod.thth_map (ththmod.py:56-116) for one eta.  Any output may
 * be NULL.  thth: float2 [n][n]; tau_inv / fd_inv: int32 [n][n]
 * (ththmod.py:94-97, bit exact); pnts: uint8 [n][n] (ththmod.py:100);
 * th_pnts: uint8 [n] crop mask of thth_redmap (ththmod.py:153-156);
 * err: int32[1], SB_ETA_INDEX_ERROR if numpy would have raised. */
int sb_thth_map(const sb_thth_geom* geom, double eta, int32_t hermitian,
                void* thth, int32_t* tau_inv, int32_t* fd_inv, uint8_t* pnts,
                uint8_t* th_pnts, int32_t* err, void* stream);

/* "Thin" (arclet) theta-theta: replaces the eta loop of
 * ththmod.single_search_thin (scintools/ththmod.py:589-627), i.e. neta calls of
 * ththmod.singularvalue_calc (:496-512 -> two_curve_map :1557-1636 ->
 * numpy.linalg.svd, S[0]).  geom describes the CS and the theta1 (column) grid
 * with THIS path's conventions: th_cents = (edges1[1:]+edges1[:-1])/2 (not
 * recentred), tau0 = tau[1], fd0 = fd[1] (ththmod.py:1602-1604), tau_absmax =
 * tau.max(), fd_half unused.  th2_cents: device float64 [n_th2], centres of
 * the arclet grid (rows).  eta1 / eta2: device float64 [neta] (main-arc and
 * arclet curvature per trial).  center_cut: columns with |theta1| < center_cut
 * are zeroed.  power=1 gathers |CS|^2 (incoherent thin search, :609).
 * Outputs: svals float64 [neta] (NaN where numpy would raise), status (SB_ETA_*
 * bits), n1_red / n2_red (cropped sizes), iters. */
int sb_thin_sweep(const sb_thth_geom* geom, const double* th2_cents, int32_t n_th2,
                  double center_cut, int32_t power, const double* eta1, const double* eta2,
                  int32_t neta, double tol, int32_t max_iter, double* svals,
                  int32_t* status, int32_t* n1_red, int32_t* n2_red, int32_t* iters,
                  void* stream);

/* Uncropped two-curvature map thth [n_th2][n_th] (float2) of
 * ththmod.two_curve_map (:1585-1617) for one (eta1, eta2); err as sb_thth_map. */
int sb_thin_map(const sb_thth_geom* geom, const double* th2_cents, int32_t n_th2,
                int32_t power, double eta1, double eta2, void* thth, int32_t* err,
                void* stream);

/* ---- phase retrieval ----------------------------------- */

/* ththmod.rev_map (scintools/ththmod.py:176-258): scatter the n x n theta-theta
 * matrix thth (float2, row-major, device) back into the conjugate spectrum
 * recov [ntau][nfd] (float2, = the reference's recov.T).  Bins are those of
 * np.histogram2d with edges (k - 0.5) * d + x0, bit-exact (x0 = tau[0] / fd[0],
 * d = tau[1]-tau[0] / fd[1]-fd[0], both > 0); weights 1/sqrt|2 eta dtheta|; bin
 * means; empty bins and the (0, 0) bin (zero Jacobian -> NaN -> nan_to_num) are 0.
 * hermitian != 0 also adds the conjugate at (-fd, -tau) (:229-256).
 * th_cents: float64 [n] on the device, already centred as in :208-209. */
int sb_rev_map(const void* thth, int32_t n, const double* th_cents, double eta, double tau0,
               double dtau, int32_t ntau, double fd0, double dfd, int32_t nfd,
               int32_t hermitian, void* recov, void* stream);

/* Largest-algebraic eigenpair of a full Hermitian float2 matrix a [n][ld] on
 * the device: eigsh(thth_red, 1, which='LA') in ththmod.modeler (:300-307).
 * w: float64 [1], v: float2 [n] (unit norm, arbitrary global phase like ARPACK),
 * info: int32 [2] = {lanczos steps, status (SB_ETA_* bits)}.  tol: residual
 * bound relative to w (<= 0: 1e-7); max_iter <= 0: 96. */
int sb_herm_eigvec(const void* a, int32_t n, int32_t ld, double tol, int32_t max_iter,
                   double* w, void* v, int32_t* info, void* stream);

/* out[:crop0, :crop1] = scale * ifft2(ifftshift(in)) (centred != 0) or
 * scale * ifft2(in), in: float2 [n0][n1], powers of two (ththmod.py:321, :1462-1465).
 * real_only != 0 writes float (the real part), else float2.  crop <= 0: full. */
int sb_ifft2_c2c_f32(const void* in, int32_t n0, int32_t n1, int32_t centred, int32_t crop0,
                     int32_t crop1, double scale, int32_t real_only, void* out, void* stream);

/* The loop of Dynspec.gerchberg_saxton (scintools/dynspec.py:1883-1896), niter
 * times, in place on wavefield (float2 [n0][n1], powers of two):
 *   CWF = fft2(w); CWF[rowmask != 0, :] = 0; w = ifft2(CWF);
 *   w = amp * exp(i angle(w)) where amp is not NaN.
 * rowmask: uint8 [n0] over the UNSHIFTED delay rows (1 where tau < 0);
 * amp: float [n0][n1] = sqrt(dyn) where dyn is finite and > 0, NaN elsewhere. */
int sb_gerchberg_saxton_f32(void* wavefield, const float* amp, const uint8_t* rowmask,
                            int32_t n0, int32_t n1, int32_t niter, void* stream);

/* ---- Dynspec 2-D FFT paths ---------------------------------------------- */

/* Dynspec.scale_dyn(scale='lambda')
 * (scintools/dynspec.py:3926-3957): not-a-knot cubic spline of every time
 * column of dyn [nf][nt] at nlam query frequencies, written flipped
 * (out [nlam][nt], wavelength ascending).  The column-independent tables are
 * built by the host (scintools_b200/dynspec.py::_spline_tables, fp64 -> fp32):
 * a, cp, inv, g: float [nf] Thomas factors of the second-derivative system,
 * p0, pn: not-a-knot end ratios, idx: int32 [nlam] interval of each query,
 * w4: float [nlam][4] weights of (y_i, y_i+1, M_i, M_i+1).  flip_rows != 0 when
 * the frequency axis of dyn is descending. */
int sb_scale_dyn_lambda_f32(const float* dyn, int32_t nf, int32_t nt, int32_t flip_rows,
                            const float* a, const float* cp, const float* inv, const float* g,
                            float p0, float pn, const int32_t* idx, const float* w4,
                            int32_t nlam, float* out, void* stream);

/* Dynspec.norm_sspec, the resampling loop (scintools/dynspec.py:2076-2107) and
 * self.powerspectrum (:2120): row ii of the secondary spectrum sspec [nr][nc] (dB,
 * already cropped / masked by the caller like :2040-2046) is resampled with
 * numpy.interp at the nq normalised Doppler values fdopnew[] on the axis
 * fdop / sqrt(tdel[ii] / eta) restricted to |fdop| <= maxnormfac * sqrt(tdel[ii]/eta).
 * out: float [nr][nq], NaN where the reference's mask is set (|fdopnew| beyond the
 * row's reach, or a NaN sample); power: double [nr] = masked mean of 10^(out/10).
 * fdop [nc], tdel [nr], fdopnew [nq]: device float64 (the reference's own axes). */
int sb_norm_sspec_f32(const float* sspec, int32_t nr, int32_t nc, const double* fdop,
                      const double* tdel, double eta, double maxnormfac,
                      const double* fdopnew, int32_t nq, float* out, double* power,
                      void* stream);

/* Delay-scrunched profile of Dynspec.norm_sspec (scintools/dynspec.py:2159-2166):
 * avg[j] = sum_ii w[ii] norm[ii][j] / sum_ii w[ii] over the unmasked (non-NaN)
 * entries of column j (np.ma.average(normSspec, axis=0, weights=w)); NaN where a
 * column is fully masked.  This is fit_arc's power-vs-curvature profile (:1156-1180). */
int sb_norm_sspec_avg_f32(const float* norm, int32_t nr, int32_t nq, const double* weights,
                          double* avg, void* stream);

/* Replaces the arithmetic of Dynspec.calc_sspec (scintools/dynspec.py:3664-3721):
 *   x = win_t[t]*win_f[f]*(dyn - mean(dyn)); x -= mean(x); [prewhite: 2x2
 *   first difference]; |FFT2 zero-padded to nrfft x ncfft|^2; fftshift;
 *   [halve: keep tau >= 0]; [postdark]; [10 log10].
 * nrfft = 2^(ceil(log2 nf)+1), ncfft likewise (dynspec.py:3677-3678).
 * dyn: float32 [nf][nt].  win_t [nt] / win_f [nf]: tapers from
 * scint_utils.get_window (scint_utils.py:810-832) or both NULL (window=None);
 * sum_win_*: their sums.  pd_fd [ncfft], pd_td [nrfft/2]: the sin^2 post-darken
 * vectors of dynspec.py:3706-3711 (only read when prewhite).  sec: float32
 * [nrfft/2 or nrfft][ncfft]; db=0 returns linear power. */
int sb_sspec_f32(const float* dyn, int32_t nf, int32_t nt, const float* win_t,
                 const float* win_f, double sum_win_t, double sum_win_f,
                 int32_t prewhite, int32_t halve, int32_t db, const float* pd_fd,
                 const float* pd_td, float* sec, void* stream);

/* Replaces Dynspec.calc_acf(method='direct') (scintools/dynspec.py:3780-3797):
 * real(fftshift(ifft2(|fft2(dyn - mean(valid), [2nf, 2nt])|^2))) [/ max].
 * acf: float32 [2nf][2nt].  subtract_mean=0 reproduces the input_dyn branch
 * (dynspec.py:3786-3789).  Any 2nf x 2nt is accepted: the transform runs on the
 * next power of two and the lags [-nf,nf) x [-nt,nt) are extracted (identical
 * by the correlation theorem).  The normalisation divides by the zero-lag
 * value, which is the maximum of an autocovariance. */
int sb_acf_f32(const float* dyn, int32_t nf, int32_t nt, int32_t subtract_mean,
               int32_t normalise, float* acf, void* stream);

/* Replaces Dynspec.calc_acf(method='sspec') (scintools/dynspec.py:3798-3807):
 * real(fftshift(fft2(linear un-halved secondary spectrum))) [/ max], with the
 * window arguments of sb_sspec_f32.  acf: float32 [nrfft][ncfft]. */
int sb_acf_sspec_f32(const float* dyn, int32_t nf, int32_t nt, const float* win_t,
                     const float* win_f, double sum_win_t, double sum_win_f,
                     int32_t normalise, float* acf, void* stream);

/* Replaces the CS stage of ththmod.single_search (scintools/ththmod.py:777-787)
 * and Dynspec.thetatheta_single (scintools/dynspec.py:1572-1579):
 *   CS = fftshift(fft2(pad(dspec, npad copies, constant pad_value)));
 *   CS[tau_rowmask] = 0
 * pad_value = NaN pads with the mean of dspec computed on the device
 * (constant_values=dspec2.mean(), ththmod.py:781) without a host pass.
 * dspec float32 [nf][nt]; cs: float2 [(npad+1)nf][(npad+1)nt]; tau_rowmask:
 * uint8 [(npad+1)nf] (1 = zero that fftshifted row) or NULL.  half_plane=1
 * writes only the fd >= 0 half, [(npad+1)nf][cs_pitch] with cs_pitch >=
 * (npad+1)nt/2 + 1 (see sb_thth_geom.cs_half): half the HBM traffic, and all
 * the theta-theta sweep needs.  half_plane=0: full array, cs_pitch ignored.
 * ncols_keep > 0 (half-plane only): compute just the first ncols_keep fd >= 0
 * columns; the others are left untouched.  The sweep gathers at
 * fd = theta_j - theta_i <= max(theta) - min(theta), so a caller that knows
 * its theta grid can skip the columns beyond that (the column passes dominate
 * the transform).  0 = all nfd/2 + 1 columns.
 * Power-of-two padded sizes take the direct radix-16 path (rows <= 65536, cols
 * <= 32768); any other size runs a chirp-z (Bluestein) transform on both axes
 * (rows <= 32768, cols <= 8192, half_plane must be 0). */
int sb_cs_f32(const float* dspec, int32_t nf, int32_t nt, int32_t npad,
              float pad_value, const uint8_t* tau_rowmask, int32_t half_plane,
              int64_t cs_pitch, int32_t ncols_keep, void* cs, void* stream);

/* Upper bound of max |CS| for the conjugate spectrum sb_cs_f32 makes from the same
 * (dspec, nf, nt, npad, pad_value): sum |dspec - c| + |c| (npad+1)^2 nf nt with c the
 * padding constant (the mean when pad_value is NaN) -- the L1 norm bounds every Fourier
 * coefficient.  One pass over dspec.  bound_out: device float.  Feeds
 * sb_thth_geom.cs_bound (the eigen solver's fp16 scale); a loose bound is fine. */
int sb_cs_bound_f32(const float* dspec, int32_t nf, int32_t nt, int32_t npad, float pad_value,
                    float* bound_out, void* stream);

/* ---- scint_sim.Simulation ------------------------------------------------ */

typedef struct sb_sim_params {
    int32_t nx, ny;
    double dx, dy, alpha, ar, psi, inner;
    double consp;   /* Simulation.set_constants (scint_sim.py:137-167), host */
} sb_sim_params;

/* Spectral amplitude w[nx][ny] (float64): the swdsp fill of
 * Simulation.get_screen (scint_sim.py:176-198, swdsp :276-292) including the
 * reference's ky=0 mirror quirk (:185). */
int sb_sim_weights(const sb_sim_params* p, double* w, void* stream);

/* xyp = real(fft2(w * (n1 + i n2))) in float64 (scint_sim.py:201-204).
 * noise_re / noise_im: float64 [nx][ny] (the reference's two randn fields, for
 * seed parity) or both NULL -> counter-based device Gaussian noise from `seed`
 * (statistically equivalent, not the MT19937 stream). */
int sb_sim_screen(int32_t nx, int32_t ny, const double* w, const double* noise_re,
                  const double* noise_im, uint64_t seed, double* xyp, void* stream);

/* Simulation.get_intensity + frfilt3 (scint_sim.py:209-236, 294-311).
 * scales_host: float64[nf] HOST array of the per-frequency `scale`
 * (:218-224).  spe_t: complex64 [nf][nx] = spe transposed (spe[:, f] is
 * column ny//2 of ifft2(filter * fft2(exp(i xyp scale)))); xyi: float32
 * [nx][ny] intensity of the LAST frequency (:232) or NULL. */
int sb_sim_intensity(int32_t nx, int32_t ny, int32_t nf, const double* xyp,
                     const double* scales_host, double ffconx, double ffcony,
                     void* spe_t, float* xyi, void* stream);

/* element-wise float64 -> float32 (n elements); complex128 -> complex64 is the
 * same call with 2n.  Lets the host layer upload the reference's float64
 * arrays unchanged. */
int sb_convert_f64_f32(const double* src, float* dst, int64_t n, void* stream);
int sb_convert_f32_f64(const float* src, double* dst, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCINT_B200_H */
