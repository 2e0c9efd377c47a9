"""Device time of the other BASELINE.json configurations (run on the GPU box):
C2  4096x8192 calc_sspec + calc_acf      C4  one 8192^2 Simulation realisation.
Prints one JSON line with CUDA-event times from sb_profile_collect and the
algorithmic-byte rooflines of SURVEY.md section 8(d)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scintools_b200 import _device as D, _lib, BasicDyn, Dynspec
from scintools_b200.scint_sim import Simulation

D.device()
L = _lib.lib
NF, NT = 4096, 8192
rng = np.random.default_rng(2)
dyn = rng.exponential(1.0, (NF, NT)).astype(np.float32)
ds = Dynspec(dyn=BasicDyn(dyn, times=10.0 * np.arange(NT), freqs=1400 + 0.03125 * np.arange(NF),
                          dt=10.0, df=0.03125), verbose=False)
def prof(fn, reps):
    fn(); torch.cuda.synchronize()
    L.sb_profile_enable(1)
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    ms = np.zeros(16); cnt = np.zeros(16, dtype=np.int32)
    L.sb_profile_collect(ms.ctypes.data, cnt.ctypes.data, 16)
    L.sb_profile_enable(0)
    return wall, ms, cnt
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
out = {"peak_gbs": peak}
w, ms, cnt = prof(lambda: ds.calc_sspec(dtype=np.float32), 3)
t = ms[6] / cnt[6]
alg = 4 * NF * NT + 4 * NF * 2 * NT
out["c2_sspec"] = {"device_ms": t, "api_wall_ms": w * 1e3, "algorithmic_GB": alg / 1e9,
                   "achieved_GBs": alg / t / 1e6, "frac": alg / t / 1e6 / peak}
w, ms, cnt = prof(lambda: ds.calc_acf(dtype=np.float32), 3)
t = ms[7] / cnt[7]
alg = 4 * NF * NT + 4 * 2 * NF * 2 * NT
out["c2_acf"] = {"device_ms": t, "api_wall_ms": w * 1e3, "algorithmic_GB": alg / 1e9,
                 "achieved_GBs": alg / t / 1e6, "frac": alg / t / 1e6 / peak}
nfreq = 16
def sim():
    return Simulation(mb2=2, ns=8192, nf=nfreq, dlam=0.25, seed=1, device_rng=True)
w, ms, cnt = prof(sim, 1)
n = 8192
out["c4_sim"] = {"nf_timed": nfreq, "screen_ms": ms[8] / cnt[8], "per_freq_ms": ms[9] / cnt[9],
                 "api_wall_s": w,
                 "per_freq_algorithmic_GB": 52 * n * n / 1e9, "per_freq_collapsed_GB": 20 * n * n / 1e9,
                 "per_freq_achieved_GBs_vs_52n2": 52 * n * n / (ms[9] / cnt[9]) / 1e6,
                 "realisation_nf256_est_s": (ms[8] / cnt[8] + 256 * ms[9] / cnt[9]) / 1e3}
print(json.dumps(out))
