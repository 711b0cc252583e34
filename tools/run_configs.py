"""Runs the BASELINE.json configurations end to end on one B200 through the host mirror and records timings and
size-independent checks -> gpurun_out/configs.json (copied to profiles/).  Development / evidence script."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200  # noqa: E402
from fungp_b200 import synth, model as M  # noqa: E402

out = {}
ctx = fungp_b200.Context(1)
which = sys.argv[1:] or ["c2", "c3", "c5"]


def tic():
    return time.perf_counter()


if "c2" in which:
    # C2: fgpm train + predict, n = 2048, 2 scalar + 2 functional (len 100, B-spline k = 5), matern5_2, 10-point multistart
    n, m = 2048, 512
    sIn, fIn, y = synth.make_data(2048, n + m, 2, [100, 100])
    t0 = tic()
    mod = M.fgpm(sIn=sIn[:n], fIn=[f[:n] for f in fIn], sOut=y[:n], kerType="matern5_2", f_pdims=5, n_starts=10, n_presample=20,
                 rng=M.NumpyRng(1), ctx=ctx)
    t_fit = tic() - t0
    t0 = tic()
    pr = M.predict(mod, sIn_pr=sIn[n:], fIn_pr=[f[n:] for f in fIn])
    t_pr = tic() - t0
    rmse = float(np.sqrt(np.mean((pr["mean"] - y[n:]) ** 2)))
    # the same fit with all 10 starts advanced together by the lock-step driver (one candidate = the full structure)
    from fungp_b200 import factory as F
    ant = F.encode_ant(2, 2, [0, 1], [0, 1], ["L2_bygroup"] * 2, [5, 5], ["B-splines"] * 2, "matern5_2")
    F.score_candidates(ctx, sIn[:n], [f[:n] for f in fIn], y[:n], ant[None, :], M.NumpyRng(1), n_starts=10, n_presample=20)
    t0 = tic()
    rl = F.score_candidates(ctx, sIn[:n], [f[:n] for f in fIn], y[:n], ant[None, :], M.NumpyRng(1), n_starts=10, n_presample=20)
    t_lock = tic() - t0
    out["c2_lockstep"] = dict(fit_and_q2_seconds=t_lock, nll=float(rl["nll"][0]), q2=float(rl["fitness"][0]))
    out["c2"] = dict(n=n, m=m, fit_seconds=t_fit, predict_seconds=t_pr, nll=mod.negLogLik, sig2=mod.kern.varHyp,
                     ls=np.r_[mod.kern.s_lsHyps, mod.kern.f_lsHyps].tolist(), rmse=rmse, sd_range=[float(pr["sd"].min()), float(pr["sd"].max())],
                     kernel_launches=ctx.launch_count())
    print("c2", out["c2"], flush=True)
    mod.close()

if "c3" in which:
    # C3: fgpm + simulate, n = 8192 train / 2048 conditional simulation points, 4 scalar + 2 functional, gauss
    n, m, nsim = 8192, 2048, 100
    sIn, fIn, y = synth.make_data(8192, n + m, 4, [100, 100])
    proj = [ctx.project(f[:n], 5, "B-splines") for f in fIn]
    coefs = [c for (_, _, c) in proj]; J = [j for (_, j, _) in proj]
    cnew = [ctx.project(f[n:], 5, "B-splines", B=B)[2] for f, (B, _, _) in zip(fIn, proj)]
    P = fungp_b200.Problem(ctx, sIn[:n], coefs, J, ["L2_bygroup"] * 2, y[:n])
    theta = 0.25 * 2.0 * P.dist_max()
    t0 = tic(); nll, s2, info = P.nll_batch("gauss", 1e-8, theta); t_nll = tic() - t0
    P.premats("gauss", 1e-8, s2[0], theta, want_L=False)            # warm: factor storage
    t0 = tic(); P.premats("gauss", 1e-8, s2[0], theta, want_L=False); t_pm = tic() - t0
    t0 = tic(); pr = P.predict(m, sIn[n:], cnew); t_pr = tic() - t0
    Z = np.random.default_rng(1).standard_normal((m, nsim))
    P.predict(m, sIn[n:], cnew); t0 = tic(); pr = P.predict(m, sIn[n:], cnew); t_pr = tic() - t0
    P.simulate(m, Z, sIn[n:], cnew, nugget_sm=1e-8)          # warm: workspaces
    t0 = tic(); sm = P.simulate(m, Z, sIn[n:], cnew, nugget_sm=1e-8); t_sm = tic() - t0
    t0 = tic(); Lh, zh = P.premats("gauss", 1e-8, s2[0], theta, want_L=True); t_pmL = tic() - t0
    t0 = tic(); prf = P.predict(m, sIn[n:], cnew, detail="full"); t_prf = tic() - t0
    # properties: simulate's mean/sd equal predict's; the sample spread follows the predictive sd
    emp_sd = sm["sims"].std(axis=0)
    out["c3"] = dict(n=n, m=m, nsim=nsim, nll_seconds=t_nll, premats_seconds=t_pm, predict_seconds=t_pr, simulate_seconds=t_sm,
                     premats_with_L_out_seconds=t_pmL, L_out_GBps=8.0 * n * n / max(t_pmL - t_pm, 1e-9) / 1e9,
                     predict_full_seconds=t_prf, Ktp_Kpp_GBps=8.0 * (n * m + m * m) / max(t_prf - t_pr, 1e-9) / 1e9,
                     nll=float(nll[0]), mean_diff=float(np.max(np.abs(sm["mean"] - pr["mean"]))), sd_diff=float(np.max(np.abs(sm["sd"] - pr["sd"]))),
                     emp_vs_pred_sd_ratio_median=float(np.median(emp_sd / np.maximum(pr["sd"], 1e-12))),
                     rmse=float(np.sqrt(np.mean((pr["mean"] - y[n:]) ** 2))))
    print("c3", out["c3"], flush=True)
    P.close()

if "c5" in which:
    # C5: large single-model fit n = 32768, 2 scalar + 2 functional (len 200, k = 10), matern3_2, likelihood + predict
    n, m = 32768, 2048
    sIn, fIn, y = synth.make_data(32768, n + m, 2, [200, 200])
    proj = [ctx.project(f[:n], 10, "B-splines") for f in fIn]
    coefs = [c for (_, _, c) in proj]; J = [j for (_, j, _) in proj]
    cnew = [ctx.project(f[n:], 10, "B-splines", B=B)[2] for f, (B, _, _) in zip(fIn, proj)]
    P = fungp_b200.Problem(ctx, sIn[:n], coefs, J, ["L2_bygroup"] * 2, y[:n])
    t0 = tic(); dm = P.dist_max(); t_dm = tic() - t0
    theta = 0.5 * 2.0 * dm
    P.nll_batch("matern3_2", 1e-8, theta)
    t0 = tic(); nll, s2, info = P.nll_batch("matern3_2", 1e-8, theta); t_nll = tic() - t0
    P.premats("matern3_2", 1e-8, s2[0], theta, want_L=False)            # warm: allocates the factor storage (8.6 GB)
    t0 = tic(); P.premats("matern3_2", 1e-8, s2[0], theta, want_L=False); t_pm = tic() - t0
    t0 = tic(); pr = P.predict(m, sIn[n:], cnew); t_pr1 = tic() - t0
    t0 = tic(); pr = P.predict(m, sIn[n:], cnew); t_pr = tic() - t0
    t0 = tic(); Lh, zh = P.premats("matern3_2", 1e-8, s2[0], theta, want_L=True); t_pmL = tic() - t0
    del Lh
    out["c5"] = dict(n=n, m=m, premats_with_L_out_seconds=t_pmL, L_out_seconds=t_pmL - t_pm, L_out_GBps=8.0 * n * n / max(t_pmL - t_pm, 1e-9) / 1e9,
                     distmax_seconds=t_dm, nll_seconds=t_nll, chol_tflops=n ** 3 / 3 / t_nll / 1e12, premats_seconds=t_pm,
                     predict_first_call_seconds=t_pr1, predict_seconds=t_pr, predict_trsm_tflops=float(n) * n * m / t_pr / 1e12, nll=float(nll[0]), info=int(info[0]),
                     rmse=float(np.sqrt(np.mean((pr["mean"] - y[n:]) ** 2))), sd_max=float(pr["sd"].max()))
    print("c5", out["c5"], flush=True)
    P.close()

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1, default=float)
