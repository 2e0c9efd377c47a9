// extern "C" surface of libscint_b200 (see include/scint_b200.h).
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "../../include/scint_b200.h"
#include "common.cuh"
#include "thth.cuh"

namespace sb {

static thread_local char g_err[512] = "";
static void* g_ws[8] = {nullptr};
static size_t g_ws_bytes[8] = {0};
static int g_sms = 148;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }
int num_sms() { return g_sms; }

static long long g_launches = 0;
void count_launch() { ++g_launches; }

// ---- profiling ---------------------------------------------------------
struct ProfPair { int id; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfPair> g_prof_pending;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t g_prof_open[PROF_COUNT];

static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) {
        cudaEvent_t e = g_prof_pool.back();
        g_prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void prof_begin(int id, cudaStream_t st) {
    if (!g_prof_on) return;
    g_prof_open[id] = prof_event();
    cudaEventRecord(g_prof_open[id], st);
}
void prof_end(int id, cudaStream_t st) {
    if (!g_prof_on) return;
    cudaEvent_t b = prof_event();
    cudaEventRecord(b, st);
    g_prof_pending.push_back({id, g_prof_open[id], b});
}

void* workspace(int slot, size_t bytes) {
    if (slot < 0 || slot >= 8) return nullptr;
    if (bytes <= g_ws_bytes[slot] && g_ws[slot]) return g_ws[slot];
    if (g_ws[slot]) {
        cudaDeviceSynchronize();
        cudaFree(g_ws[slot]);
        g_ws[slot] = nullptr;
        g_ws_bytes[slot] = 0;
    }
    size_t want = (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        set_error("workspace slot %d: cudaMalloc(%zu) -> %s", slot, want,
                  cudaGetErrorString(e));
        cudaGetLastError();
        return nullptr;
    }
    g_ws[slot] = p;
    g_ws_bytes[slot] = want;
    return p;
}

void workspace_release() {
    cudaDeviceSynchronize();
    for (int i = 0; i < 8; ++i) {
        if (g_ws[i]) cudaFree(g_ws[i]);
        g_ws[i] = nullptr;
        g_ws_bytes[i] = 0;
    }
}

int eta_sweep(const ThthGeom& g, const double* th_host, const double* d_etas,
              int neta, double tol, int max_iter, double* d_eigs,
              int* d_status, int* d_nred, int* d_iters, cudaStream_t st);
int thth_map(const ThthGeom& g, double eta, int hermitian, float2* d_out,
             int* d_tau_inv, int* d_fd_inv, unsigned char* d_pnts,
             unsigned char* d_th_pnts, int* d_err, cudaStream_t st);

int sspec(const float* dyn, int nf, int nt, const float* wt, const float* wf,
          double swt, double swf, int prewhite, int halve, int db,
          const float* pd1, const float* pd2, float* sec, cudaStream_t st,
          int noshift = 0);
int conj_spectrum(const float* dyn, int nf, int nt, int npad, float pad_value,
                  const unsigned char* rowmask, int half, long pitch, int ncols_keep,
                  float2* CS, cudaStream_t st);
int acf(const float* dyn, int nf, int nt, int subtract_mean, int normalise,
        float* out, cudaStream_t st);
int acf_sspec(const float* dyn, int nf, int nt, const float* wt, const float* wf,
              double swt, double swf, int normalise, float* out, cudaStream_t st);
void twiddle_release();
struct SimParams {
    int nx, ny;
    double dx, dy, alpha, ar, psi, inner, consp;
};
int sim_weights(const SimParams& p, double* w, cudaStream_t st);
int sim_screen(int nx, int ny, const double* w, const double* n1, const double* n2,
               unsigned long long seed, double* xyp, cudaStream_t st);
int sim_intensity(int nx, int ny, int nf, const double* xyp, const double* scales_host,
                  double ffconx, double ffcony, float2* spe_t, float* xyi,
                  cudaStream_t st);

template <typename A, typename B>
__global__ void convert_kernel(const A* __restrict__ a, B* __restrict__ b, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        b[i] = (B)a[i];
}

struct ThinGeom {
    ThthGeom g;
    const double* th2;
    int n2;
    double tau_max;
    double center_cut;
    int power;
};
int thin_sweep(const ThinGeom& t, const double* d_eta1, const double* d_eta2, int neta,
               double tol, int max_iter, double* d_sv, int* d_status, int* d_n1, int* d_n2,
               int* d_iters, cudaStream_t st);
int thin_map(const ThinGeom& t, double e1, double e2, float2* d_out, int* d_err,
             cudaStream_t st);

int rev_map(const float2* thth, int n, const double* th_dev, double eta, double tau0,
            double dtau, int ntau, double fd0, double dfd, int nfd, int hermitian,
            float2* recov, cudaStream_t st);
int herm_eigvec(const float2* A, int n, int ld, double tol, int max_iter, double* w_dev,
                float2* V_dev, int* info_dev, cudaStream_t st);
int ifft2_c2c(const float2* in, int n0, int n1, int centred, int crop0, int crop1,
              double scale, int real_only, void* out, cudaStream_t st);

int gerchberg_saxton(float2* W, const float* amp, const unsigned char* rowmask, int n0, int n1,
                     int niter, cudaStream_t st);

int conj_spectrum_bound(const float* dyn, int nf, int nt, int npad, float pad_value, float* out,
                        cudaStream_t st);
int norm_sspec_rows(const float* sspec, int nr, int nc, const double* fdop, const double* tdel,
                    double eta, double maxnormfac, const double* fdopnew, int nq, float* out,
                    double* power, cudaStream_t st);
int norm_sspec_avg(const float* norm, int nr, int nq, const double* weights, double* avg,
                   cudaStream_t st);
int scale_dyn_lambda(const float* dyn, int nf, int nt, int flip, const float* a,
                     const float* cp, const float* inv, const float* g, float p0, float pn,
                     const int* idx, const float4* W, int nlam, float* out, cudaStream_t st);

static int to_geom(const sb_thth_geom* in, ThthGeom* g) {
    SB_ARG(in != nullptr);
    SB_ARG(in->ntau > 0 && in->nfd > 0);
    SB_ARG(in->th_cents != nullptr && in->th_cents_host != nullptr);
    SB_ARG(in->n_th > 0);
    g->cs = (const float2*)in->cs;
    g->ntau = in->ntau;
    g->nfd = in->nfd;
    g->tau0 = in->tau0;
    g->dtau = in->dtau;
    g->half_dtau = in->dtau / 2;  // dtau / 2 (ththmod.py:95)
    g->tau_absmax = in->tau_absmax;
    g->fd0 = in->fd0;
    g->dfd = in->dfd;
    g->half_dfd = in->dfd / 2;
    g->inv_dtau = 1.0 / in->dtau;
    g->inv_dfd = 1.0 / in->dfd;
    g->fd_half = in->fd_half;
    g->th = in->th_cents;
    g->n = in->n_th;
    g->coherent = in->coherent;
    g->cs_half = in->cs_half;
    g->cs_valid_cols = in->cs_half ? in->cs_valid_cols : 0;
    g->cs_bound = in->cs_bound;
    g->cs_pitch = in->cs_half ? in->cs_pitch : (in->cs_pitch > 0 ? in->cs_pitch : in->nfd);
    SB_ARG(!in->cs_half || (in->cs_pitch >= in->nfd / 2 + 1 && in->nfd % 2 == 0));
    return SB_OK;
}

}  // namespace sb

extern "C" {

int sb_abi_version(void) { return 2; }
const char* sb_last_error(void) { return sb::last_error(); }

int sb_init(int device) {
    SB_CUDA(cudaSetDevice(device));
    SB_CUDA(cudaFree(0));
    cudaDeviceProp prop;
    SB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        sb::set_error("device %d is sm_%d%d; libscint_b200 is sm_100a only",
                      device, prop.major, prop.minor);
        return SB_ERR_UNSUPPORTED;
    }
    sb::g_sms = prop.multiProcessorCount;
    return SB_OK;
}

int64_t sb_launch_count(void) { return sb::g_launches; }

int sb_profile_enable(int32_t on) {
    sb::g_prof_on = on != 0;
    return SB_OK;
}

int sb_profile_collect(double* ms_host, int32_t* count_host, int32_t n) {
    SB_ARG(ms_host && count_host && n >= sb::PROF_COUNT);
    for (int i = 0; i < n; ++i) { ms_host[i] = 0.0; count_host[i] = 0; }
    SB_CUDA(cudaDeviceSynchronize());
    for (auto& p : sb::g_prof_pending) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, p.a, p.b);
        ms_host[p.id] += ms;
        count_host[p.id] += 1;
        sb::g_prof_pool.push_back(p.a);
        sb::g_prof_pool.push_back(p.b);
    }
    sb::g_prof_pending.clear();
    return SB_OK;
}

int sb_release(void) {
    sb::workspace_release();
    sb::twiddle_release();
    return SB_OK;
}

int sb_eta_sweep(const sb_thth_geom* geom, const double* etas, int32_t neta,
                 double tol, int32_t max_iter, double* eigs, int32_t* status,
                 int32_t* nred, int32_t* iters, void* stream) {
    sb::ThthGeom g;
    int rc = sb::to_geom(geom, &g);
    if (rc) return rc;
    SB_ARG(neta >= 0 && etas && eigs && status && nred && iters);
    SB_ARG(geom->cs != nullptr);
    if (!(tol > 0)) tol = 2e-5;
    return sb::eta_sweep(g, geom->th_cents_host, etas, neta, tol, max_iter,
                         eigs, status, nred, iters, (cudaStream_t)stream);
}

int sb_thth_map(const sb_thth_geom* geom, double eta, int32_t hermitian,
                void* thth, int32_t* tau_inv, int32_t* fd_inv, uint8_t* pnts,
                uint8_t* th_pnts, int32_t* err, void* stream) {
    sb::ThthGeom g;
    int rc = sb::to_geom(geom, &g);
    if (rc) return rc;
    const bool wants_map = thth || tau_inv || fd_inv || pnts;
    SB_ARG(!wants_map || (err != nullptr && geom->cs != nullptr));
    return sb::thth_map(g, eta, hermitian, (float2*)thth, tau_inv, fd_inv,
                        pnts, th_pnts, err, (cudaStream_t)stream);
}

static int to_thin(const sb_thth_geom* geom, const double* th2, int32_t n2, double center_cut,
                   int32_t power, sb::ThinGeom* t) {
    int rc = sb::to_geom(geom, &t->g);
    if (rc) return rc;
    SB_ARG(geom->cs != nullptr && th2 != nullptr && n2 > 0);
    t->th2 = th2;
    t->n2 = n2;
    t->tau_max = geom->tau_absmax;     // carries tau.max() for the thin map
    t->center_cut = center_cut;
    t->power = power;
    return SB_OK;
}

int sb_thin_sweep(const sb_thth_geom* geom, const double* th2_cents, int32_t n_th2,
                  double center_cut, int32_t power, const double* eta1, const double* eta2,
                  int32_t neta, double tol, int32_t max_iter, double* svals,
                  int32_t* status, int32_t* n1_red, int32_t* n2_red, int32_t* iters,
                  void* stream) {
    sb::ThinGeom t;
    int rc = to_thin(geom, th2_cents, n_th2, center_cut, power, &t);
    if (rc) return rc;
    SB_ARG(neta >= 0 && eta1 && eta2 && svals && status && n1_red && n2_red && iters);
    if (!(tol > 0)) tol = 2e-5;
    return sb::thin_sweep(t, eta1, eta2, neta, tol, max_iter, svals, status, n1_red, n2_red,
                          iters, (cudaStream_t)stream);
}

int sb_thin_map(const sb_thth_geom* geom, const double* th2_cents, int32_t n_th2,
                int32_t power, double eta1, double eta2, void* thth, int32_t* err,
                void* stream) {
    sb::ThinGeom t;
    int rc = to_thin(geom, th2_cents, n_th2, 0.0, power, &t);
    if (rc) return rc;
    SB_ARG(thth && err);
    return sb::thin_map(t, eta1, eta2, (float2*)thth, err, (cudaStream_t)stream);
}

int sb_rev_map(const void* thth, int32_t n, const double* th_cents, double eta, double tau0,
               double dtau, int32_t ntau, double fd0, double dfd, int32_t nfd,
               int32_t hermitian, void* recov, void* stream) {
    SB_ARG(thth && th_cents && recov && n >= 1 && ntau >= 1 && nfd >= 1);
    return sb::rev_map((const float2*)thth, n, th_cents, eta, tau0, dtau, ntau, fd0, dfd, nfd,
                       hermitian, (float2*)recov, (cudaStream_t)stream);
}

int sb_herm_eigvec(const void* a, int32_t n, int32_t ld, double tol, int32_t max_iter,
                   double* w, void* v, int32_t* info, void* stream) {
    SB_ARG(a && w && v && info && n >= 1 && ld >= n);
    return sb::herm_eigvec((const float2*)a, n, ld, tol, max_iter, w, (float2*)v, info,
                           (cudaStream_t)stream);
}

int sb_ifft2_c2c_f32(const void* in, int32_t n0, int32_t n1, int32_t centred, int32_t crop0,
                     int32_t crop1, double scale, int32_t real_only, void* out, void* stream) {
    SB_ARG(in && out);
    return sb::ifft2_c2c((const float2*)in, n0, n1, centred, crop0, crop1, scale, real_only,
                         out, (cudaStream_t)stream);
}

int sb_gerchberg_saxton_f32(void* wavefield, const float* amp, const uint8_t* rowmask,
                            int32_t n0, int32_t n1, int32_t niter, void* stream) {
    SB_ARG(wavefield && amp && rowmask && niter >= 0);
    return sb::gerchberg_saxton((float2*)wavefield, amp, rowmask, n0, n1, niter,
                                (cudaStream_t)stream);
}

int sb_scale_dyn_lambda_f32(const float* dyn, int32_t nf, int32_t nt, int32_t flip_rows,
                            const float* a, const float* cp, const float* inv, const float* g,
                            float p0, float pn, const int32_t* idx, const float* w4,
                            int32_t nlam, float* out, void* stream) {
    SB_ARG(dyn && a && cp && inv && g && idx && w4 && out && nt >= 1 && nlam >= 1);
    return sb::scale_dyn_lambda(dyn, nf, nt, flip_rows, a, cp, inv, g, p0, pn, idx,
                                (const float4*)w4, nlam, out, (cudaStream_t)stream);
}

int sb_norm_sspec_f32(const float* sspec, int32_t nr, int32_t nc, const double* fdop,
                      const double* tdel, double eta, double maxnormfac,
                      const double* fdopnew, int32_t nq, float* out, double* power,
                      void* stream) {
    SB_ARG(sspec && fdop && tdel && fdopnew && out && power && nr >= 1 && nc >= 2 && nq >= 1);
    return sb::norm_sspec_rows(sspec, nr, nc, fdop, tdel, eta, maxnormfac, fdopnew, nq, out,
                               power, (cudaStream_t)stream);
}

int sb_norm_sspec_avg_f32(const float* norm, int32_t nr, int32_t nq, const double* weights,
                          double* avg, void* stream) {
    SB_ARG(norm && weights && avg && nr >= 1 && nq >= 1);
    return sb::norm_sspec_avg(norm, nr, nq, weights, avg, (cudaStream_t)stream);
}

int sb_sspec_f32(const float* dyn, int32_t nf, int32_t nt, const float* win_t,
                 const float* win_f, double sum_win_t, double sum_win_f,
                 int32_t prewhite, int32_t halve, int32_t db, const float* pd_fd,
                 const float* pd_td, float* sec, void* stream) {
    SB_ARG(dyn && sec && nf >= 2 && nt >= 2);
    SB_ARG((win_t == nullptr) == (win_f == nullptr));
    SB_ARG(!prewhite || (halve && pd_fd && pd_td));
    return sb::sspec(dyn, nf, nt, win_t, win_f, sum_win_t, sum_win_f, prewhite,
                     halve, db, pd_fd, pd_td, sec, (cudaStream_t)stream);
}

int sb_acf_f32(const float* dyn, int32_t nf, int32_t nt, int32_t subtract_mean,
               int32_t normalise, float* acf, void* stream) {
    SB_ARG(dyn && acf && nf >= 1 && nt >= 1);
    return sb::acf(dyn, nf, nt, subtract_mean, normalise, acf, (cudaStream_t)stream);
}

int sb_acf_sspec_f32(const float* dyn, int32_t nf, int32_t nt, const float* win_t,
                     const float* win_f, double sum_win_t, double sum_win_f,
                     int32_t normalise, float* acf, void* stream) {
    SB_ARG(dyn && acf && nf >= 2 && nt >= 2);
    SB_ARG((win_t == nullptr) == (win_f == nullptr));
    return sb::acf_sspec(dyn, nf, nt, win_t, win_f, sum_win_t, sum_win_f, normalise,
                         acf, (cudaStream_t)stream);
}

int sb_cs_f32(const float* dspec, int32_t nf, int32_t nt, int32_t npad,
              float pad_value, const uint8_t* tau_rowmask, int32_t half_plane,
              int64_t cs_pitch, int32_t ncols_keep, void* cs, void* stream) {
    SB_ARG(dspec && cs && nf >= 1 && nt >= 1 && npad >= 0);
    SB_ARG(!half_plane || cs_pitch >= (int64_t)(npad + 1) * nt / 2 + 1);
    return sb::conj_spectrum(dspec, nf, nt, npad, pad_value, tau_rowmask,
                             half_plane, (long)cs_pitch, ncols_keep, (float2*)cs,
                             (cudaStream_t)stream);
}

int sb_cs_bound_f32(const float* dspec, int32_t nf, int32_t nt, int32_t npad, float pad_value,
                    float* bound_out, void* stream) {
    SB_ARG(dspec && bound_out && nf >= 1 && nt >= 1 && npad >= 0);
    return sb::conj_spectrum_bound(dspec, nf, nt, npad, pad_value, bound_out, (cudaStream_t)stream);
}

int sb_sim_weights(const sb_sim_params* p, double* w, void* stream) {
    SB_ARG(p && w);
    sb::SimParams q{p->nx, p->ny, p->dx, p->dy, p->alpha, p->ar, p->psi, p->inner, p->consp};
    return sb::sim_weights(q, w, (cudaStream_t)stream);
}

int sb_sim_screen(int32_t nx, int32_t ny, const double* w, const double* noise_re,
                  const double* noise_im, uint64_t seed, double* xyp, void* stream) {
    SB_ARG(w && xyp && ((noise_re == nullptr) == (noise_im == nullptr)));
    return sb::sim_screen(nx, ny, w, noise_re, noise_im, seed, xyp, (cudaStream_t)stream);
}

int sb_sim_intensity(int32_t nx, int32_t ny, int32_t nf, const double* xyp,
                     const double* scales_host, double ffconx, double ffcony,
                     void* spe_t, float* xyi, void* stream) {
    SB_ARG(xyp && scales_host && spe_t && nf >= 1);
    return sb::sim_intensity(nx, ny, nf, xyp, scales_host, ffconx, ffcony,
                             (float2*)spe_t, xyi, (cudaStream_t)stream);
}

int sb_convert_f64_f32(const double* src, float* dst, int64_t n, void* stream) {
    SB_ARG(src && dst && n >= 0);
    if (n == 0) return SB_OK;
    sb::convert_kernel<double, float><<<sb::num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(src, dst, n);
    SB_LAUNCH_CHECK();
    return SB_OK;
}
int sb_convert_f32_f64(const float* src, double* dst, int64_t n, void* stream) {
    SB_ARG(src && dst && n >= 0);
    if (n == 0) return SB_OK;
    sb::convert_kernel<float, double><<<sb::num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(src, dst, n);
    SB_LAUNCH_CHECK();
    return SB_OK;
}

}  // extern "C"
