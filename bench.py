#!/usr/bin/env python
"""Benchmark of the theta-theta curvature sweep (BASELINE.json metric:
"theta-theta eta-trials/sec on 4096x8192 dynspec").

One *step* = one pass of the hot path over one dynamic spectrum:
  conjugate spectrum of the 4096x8192 chunk (npad=3 -> 16384x32768 c64, 4.3 GB)
  + dominant-eigenvalue sweep over 1024 curvatures on a 512-point theta grid
  (+ one all-gather of the per-eta eigenvalues when N > 1).
Weak scaling: every rank sweeps its own block of 1024 etas of a global
N x 1024 log grid (the CS is recomputed per rank; no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

The b200 arm reports device-resident throughput (`value`), end-to-end
throughput through the public API with pinned host buffers (`e2e`), the
roofline of the dominant kernel and a CPU baseline measured in the same run.
The reference arm times the reference's CPU algorithm (oracle port:
numpy gather + scipy ARPACK, pocketfft CS) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "theta-theta eta-trials/sec on 4096x8192 dynspec"
NF, NT, NPAD = 4096, 8192, 3
NETA, NEDGE = 1024, 512
ETA_TRUE = 0.08          # s^3
DT, DF = 10.0, 0.03125   # s, MHz
EDGE_LIM = 10.0          # mHz
FW = 0.1


def make_dynspec(seed=3, nf=NF, nt=NT):
    """64-image 1-D screen, eta_true = 0.08 s^3.
    E(f,t) = sum_k a_k exp(2 pi i (fd_k t - tau_k f)) is separable per image,
    so it is one (nf x 64) @ (64 x nt) product."""
    rng = np.random.default_rng(seed)
    nimg = 64
    fdk = rng.uniform(-8.0, 8.0, nimg)                       # mHz
    ak = (rng.normal(size=nimg) + 1j * rng.normal(size=nimg)) / np.sqrt(2)
    ak = ak * np.exp(-(fdk / 4.0) ** 2)
    tauk = ETA_TRUE * fdk ** 2                               # us
    t = DT * np.arange(nt)
    f = DF * np.arange(nf)                                   # MHz offset
    U = np.exp(2j * np.pi * 1e-3 * fdk[:, None] * t[None, :])
    V = np.exp(-2j * np.pi * tauk[None, :] * f[:, None]) * ak[None, :]
    E = (V.astype(np.complex64) @ U.astype(np.complex64))
    dyn = (E.real ** 2 + E.imag ** 2).astype(np.float32)
    dyn += rng.normal(0.0, 0.2 * dyn.mean(), dyn.shape).astype(np.float32)
    dyn -= dyn.mean()
    return dyn, 1400.0 + f, t


def eta_grid(n_total):
    return np.logspace(np.log10(ETA_TRUE / 2), np.log10(2 * ETA_TRUE), n_total)


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed
# `ncu --set full` capture of this same command (profiles/
# r1_ncu_full_summary_final.csv); bench.py cannot run ncu on itself.
NCU_TRAFFIC_BYTES = {"thth_eig": 19.325e9 + 0.004e9, "thth_build": 1.368e9 + 1.042e9,
                     "cs_rows": 0.134e9 + 0.164e9, "cs_colA": 0.215e9 + 0.800e9,
                     "cs_colB": 0.858e9 + 0.810e9}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampling in the background (started well before the timed
    region so that it is already polling); `stop(t0, t1)` keeps the samples
    whose timestamp falls inside the timed window (wall clock), falling back to
    the samples taken under load (power above half of the maximum seen)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t0=None, t1=None):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]), float(parts[3]),
                             [nm for nm, v in zip(names, parts[4:8])
                              if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        if not rows:
            return out
        sel = [r for r in rows if t0 is not None and t0 - 0.03 <= r[0] <= t1 + 0.03]
        how = "timed window"
        if not sel:
            pmax = max(r[3] for r in rows)
            sel = [r for r in rows if r[3] >= 0.5 * pmax]
            how = "samples under load around the timed window"
        reasons = sorted({x for r in sel for x in r[4]})
        return {"sm_mhz": float(np.median([r[1] for r in sel])),
                "sm_max_mhz": float(max(r[2] for r in sel)), "reasons": reasons,
                "power_w_max": float(max(r[3] for r in sel)),
                "samples": len(sel), "from": how}


# --------------------------------------------------------------------------
# CPU arms (oracle port of the reference algorithm)
# --------------------------------------------------------------------------
_G = {}


def _eval_one(eta):
    from oracle import thth_oracle as TO
    if _G.get("procs", 1) > 1 and not _G.get("limited"):
        try:    # one BLAS thread per pool worker: no oversubscription
            from threadpoolctl import threadpool_limits
            _G["limited"] = threadpool_limits(1)
        except Exception:
            _G["limited"] = True
    try:
        return TO.Eval_calc(_G["CS"], _G["tau"], _G["fd"], eta, _G["edges"])
    except Exception:
        return float("nan")


def cpu_sample(CS, tau, fd, edges, etas, procs):
    """Time len(etas) eta-trials of the oracle; returns (seconds, eigs)."""
    _G.update(CS=CS, tau=tau, fd=fd, edges=edges, procs=procs)
    t0 = time.perf_counter()
    if procs <= 1:
        eigs = [_eval_one(e) for e in etas]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            eigs = pool.map(_eval_one, list(etas), chunksize=1)
    return time.perf_counter() - t0, np.array(eigs)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import scipy.fft as sfft
    from oracle import thth_oracle as TO
    cores = len(os.sched_getaffinity(0))
    dyn, freq, t = make_dynspec()
    fd = TO.fft_axis(t, "mHz", NPAD)
    tau = TO.fft_axis(freq, "us", NPAD)
    edges = np.linspace(-EDGE_LIM, EDGE_LIM, NEDGE)
    t0 = time.perf_counter()
    pad = np.zeros(((NPAD + 1) * NF, (NPAD + 1) * NT), dtype=np.float32)
    pad[:NF, :NT] = dyn
    CS = sfft.fftshift(sfft.fft2(pad, workers=cores))       # pocketfft, c64
    del pad
    t_cs = time.perf_counter() - t0
    etas = eta_grid(NETA * args.gpus)
    nsamp = max(cores, 4)
    times = []
    rng = np.random.default_rng(0)
    for it in range(args.warmup + args.steps):
        sel = np.sort(rng.choice(len(etas), nsamp, replace=False))
        dt_, _ = cpu_sample(CS, tau, fd, edges, etas[sel], cores)
        if it >= args.warmup:
            times.append(dt_)
    per_step = float(np.mean(times))
    value = nsamp / per_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "eta-trials/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3 eta-sweep: 4096x8192 dynspec, npad=3, "
                               "512-pt theta grid, %d etas" % NETA,
                   "note": "reference algorithm (numpy gather + scipy ARPACK "
                           "eigsh) via the oracle port; astropy unavailable"},
        "cpu_baseline": {"value": value, "unit": "eta-trials/s", "cores": cores,
                         "kind": "port",
                         "sample": "%d eta-trials per step over a %d-process "
                                   "fork pool on the full-size CS; the CS "
                                   "(scipy pocketfft c64, %d threads) took "
                                   "%.1f s once and is NOT counted"
                                   % (nsamp, cores, cores, t_cs)},
        "e2e": {"value": value, "unit": "eta-trials/s",
                "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------
def b200_arm(args):
    import torch
    import torch.distributed as dist
    from scintools_b200 import _device as D
    from scintools_b200 import _lib
    from scintools_b200 import ththmod as thth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = D.device()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    clocks = ClockSampler(local)      # polling from the start; windowed later
    dyn, freq, t = make_dynspec()
    fd = np.asarray(thth.fft_axis(t, "mHz", NPAD))
    tau = np.asarray(thth.fft_axis(freq, "us", NPAD))
    edges = np.linspace(-EDGE_LIM, EDGE_LIM, NEDGE)
    etas_all = eta_grid(NETA * world)
    etas = np.ascontiguousarray(etas_all[rank * NETA:(rank + 1) * NETA])

    # device-resident inputs
    d_dyn = D.upload(dyn)
    ntau, nfd = (NPAD + 1) * NF, (NPAD + 1) * NT
    pitch = nfd // 2 + 16          # Hermitian half-plane CS (fd >= 0)
    d_cs = D.empty((ntau, pitch, 2), torch.float32)
    # the sweep gathers at fd = theta_j - theta_i <= 2*EDGE_LIM: only those fd
    # columns of the CS are computed (exactly what single_search does)
    keep = thth.needed_fd_columns(fd, edges) or 0
    cs = thth.DeviceCS(d_cs, nfd=nfd, ncols_valid=keep or None)
    geom = thth._Geom(cs, tau, fd, edges, True)
    d_etas = D.upload(etas)
    d_eigs = D.empty((NETA,), torch.float64)
    d_stat = D.empty((NETA,), torch.int32)
    d_nred = D.empty((NETA,), torch.int32)
    d_iter = D.empty((NETA,), torch.int32)
    gathered = D.empty((world * NETA,), torch.float64) if world > 1 else None
    stream = D.stream_ptr()
    L = _lib.lib

    def step():
        _lib.check(L.sb_cs_f32(d_dyn.data_ptr(), NF, NT, NPAD, 0.0, 0, 1, pitch, keep,
                               d_cs.data_ptr(), stream))
        _lib.check(L.sb_eta_sweep(geom.ref, d_etas.data_ptr(), NETA, thth.DEFAULT_TOL,
                                  0, d_eigs.data_ptr(), d_stat.data_ptr(),
                                  d_nred.data_ptr(), d_iter.data_ptr(), stream))
        if world > 1:
            dist.all_gather_into_tensor(gathered, d_eigs)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()
    launches0 = L.sb_launch_count()
    L.sb_profile_enable(1)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    sync_all()
    wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    sync_all()
    wall1 = time.time()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = int(L.sb_launch_count() - launches0)
    prof_ms = (np.zeros(16), np.zeros(16, dtype=np.int32))
    _lib.check(L.sb_profile_collect(prof_ms[0].ctypes.data, prof_ms[1].ctypes.data, 16))
    L.sb_profile_enable(0)
    clk = clocks.stop(wall0, wall1)

    eigs = d_eigs.cpu().numpy()
    nred = d_nred.cpu().numpy().astype(np.int64)
    iters = d_iter.cpu().numpy()
    status = d_stat.cpu().numpy()
    ms_step = ms_total / args.steps
    value = world * NETA / (ms_step * 1e-3)

    names = ["cs_rows", "cs_colA", "cs_colB", "thth_prep", "thth_build",
             "thth_eig", "sspec", "acf", "sim_screen", "sim_freq"]
    kern = {n: (prof_ms[0][i] / max(1, prof_ms[1][i]))
            for i, n in enumerate(names) if prof_ms[1][i]}
    # algorithmic bytes of one launch of the sweep kernels: one c64 gather of the
    # strict upper triangle + one f64 eigenvalue per eta
    alg_bytes = float(np.sum(8 * nred * (nred - 1) // 2 + 8))
    dom = max((k for k in kern if k.startswith("thth")), key=lambda k: kern[k])
    peak, peak_src = peak_hbm()
    ach = alg_bytes / (kern[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak,
                "traffic": NCU_TRAFFIC_BYTES.get(dom),
                "traffic_source": "profiles/r1_ncu_full_summary_final.csv (bytes per launch)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "iterative solver: the 1 MB triangle is streamed once per Lanczos step "
                        "(~18 steps) = `traffic`; the on-chip variant that reads it once "
                        "(eig_cluster.cu, traffic 1.10 GB) measured slower, see DESIGN.md",
                "kernel_ms": kern}

    # ---- end to end through the public API, pinned host input ----------
    h_dyn = torch.from_numpy(dyn).pin_memory()
    params = [h_dyn.numpy(), freq, t, etas, edges, None, False, FW, NPAD, True,
              0.0, False]
    res = None
    for _ in range(min(2, args.warmup)):
        res = thth.single_search(params)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = thth.single_search(params)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = world * NETA / (float(e2e_t.item()) / args.steps)
    e2e = {"value": e2e_val, "unit": "eta-trials/s",
           "h2d_bytes_per_step": int(dyn.nbytes + etas.nbytes + 8 * (NEDGE - 1)),
           "d2h_bytes_per_step": int(8 * NETA),
           "api": "scintools_b200.ththmod.single_search(params) incl. host "
                  "parabola fit; dyn float32 in pinned host memory",
           "eta_fit": float(res[0]) if res is not None else None}

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            from oracle import thth_oracle as TO   # checker / CPU baseline only
            full = thth.conjugate_spectrum(dyn, NPAD, 0.0)      # all columns, for the CPU leg
            CS_host = full.numpy().astype(np.complex64)
            del full
            sel = np.linspace(0, NETA - 1, 8).astype(int)
            secs, ref = cpu_sample(CS_host, tau, fd, edges, etas[sel], 1)
            rel = np.abs(eigs[sel] - ref) / np.abs(ref)
            cpu = {"value": len(sel) / secs, "unit": "eta-trials/s", "cores": 1,
                   "kind": "port",
                   "sample": "8 of 1024 eta-trials (oracle Eval_calc: numpy "
                             "gather + scipy ARPACK) on the GPU-built 16384x32768 "
                             "CS; the CPU fft2 of the CS is not counted",
                   "max_rel_err_vs_gpu": float(np.nanmax(rel))}
        line = {
            "metric": METRIC, "value": value, "unit": "eta-trials/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3 eta-sweep: 4096x8192 dynspec (1-D screen, "
                                   "64 images, eta_true=0.08 s^3), npad=3 -> CS "
                                   "16384x32768 (fd>=0 half stored, 2.15 GB c64) recomputed every step, "
                                   "512-pt theta grid, 1024 etas per GPU",
                       "etas_total": world * NETA,
                       "cs_columns": "%d of %d fd>=0 columns computed (those the "
                                     "512-pt theta grid can reach)" % (keep or nfd // 2 + 1,
                                                                      nfd // 2 + 1),
                       "l2": "inputs larger than L2 (CS half-plane 2.15 GB, matrices 1.07 GB)",
                       "tol": thth.DEFAULT_TOL,
                       "parallelism": "eta blocks per rank, CS replicated, one "
                                      "NCCL all-gather of eigenvalues per step"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clk,
            "sweep": {"nred_min": int(nred.min()), "nred_max": int(nred.max()),
                      "iters_mean": float(iters.mean()), "iters_max": int(iters.max()),
                      "status_nonzero": int((status != 0).sum()),
                      "eta_peak": float(etas[np.nanargmax(eigs)])},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true",
                    help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)
    return b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
