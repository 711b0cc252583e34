#!/usr/bin/env python
"""bench.py -- headline benchmark of the funGp-b200 hot path (contract: see the task statement / DESIGN.md).

Metric (BASELINE.json): negLogLik evaluations / second at n = 8192 on one GPU.
Workload = BASELINE.json configs[2] sizes ("C3": n = 8192 training points, 4 scalar + 2 functional inputs of length
100 projected on B-splines with p = 5, L2_bygroup, gauss kernel, nugget 1e-8, FP64).  One *step* = one
finite-difference gradient of the likelihood as stats::optim evaluates it: 2d+1 = 13 length-scale vectors
(theta, theta +- 1e-3 e_i), i.e. 13 independent evaluations of  assembly -> Cholesky -> L^-1 y -> log det  in one
batched C-ABI call (fgp_nll_batch).  Every evaluation builds and factors its own 8192 x 8192 FP64 matrix (512 MiB,
larger than the 126 MB L2, so nothing is served from cache between steps).

  value   evaluations/s with the data set resident in HBM (fgp_problem created once); each step still copies the
          13 x d length-scales in and 13 x (nll, sigma^2, info) out -- that is the call's interface.
  e2e     the same through the reference-facing call sequence with HOST buffers only: every step creates the device
          problem from host arrays (H2D of coordinates, Gram matrices and y), evaluates the batch, reads the results
          back and destroys the problem.
With --gpus N (torchrun, one process per GPU) every rank evaluates its own 13-point batch on its own GPU: independent
work items, no collective on the data path; value = all ranks' evaluations / max-over-ranks time.
torch.distributed is used for the barrier and the max over ranks only.

--impl reference runs the CPU restatement of the reference (oracle/, numpy/scipy, all host threads) on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "negLogLik evals/sec at n=8192 (1 GPU)"
UNIT = "evals/s"
CFG = dict(n=8192, ds=4, f_lens=[100, 100], pdims=[5, 5], basis="B-splines", dis="L2_bygroup", ker="gauss", nugget=1e-8,
           theta_frac=0.25, seed=8192)


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class ClockSampler:
    """SM clock / throttle reasons / power during the timed region (B200_PROFILING.md recipe), sampled in-process through
    NVML every 250 ms (spawning nvidia-smi -lms next to the timed loop perturbs kernel launches)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        except Exception:
            self.nvml = None
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nvml
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((sm, mx, pw, rs))
            except Exception:
                pass
            self.stop_flag.wait(0.25)

    def stop(self):
        if self.nvml is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvml unavailable"])
        self.stop_flag.set()
        self.thread.join(timeout=2)
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm = [x[0] for x in self.samples]; pw = [x[2] for x in self.samples]
        busy = [c for c, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        reasons = sorted({name for (_, _, _, rs) in self.samples for bit, name in self.REASONS.items() if rs & bit})
        return dict(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(x[1] for x in self.samples)), reasons=reasons,
                    samples=len(sm), power_w_max=float(max(pw)))


FACTORY = dict(n=1024, ds=5, f_lens=[10, 22, 50], seed=1024, n_starts=1, n_presample=20, nugget=1e-8, n_cand=2000, list_seed=7)


def factory_candidates(n_cand, seed=7):
    """Fixed pre-generated list of candidate structures for BASELINE.json configs[3] (5 scalar + 3 functional inputs):
    random draws over the reference's solution space (state, distance, projection dimension, basis, kernel)."""
    from fungp_b200 import factory as F
    rng = np.random.default_rng(seed)
    ds, df, lens = FACTORY["ds"], len(FACTORY["f_lens"]), FACTORY["f_lens"]
    ants = []
    while len(ants) < n_cand:
        s_act = [i for i in range(ds) if rng.random() < 0.6]
        f_act = [j for j in range(df) if rng.random() < 0.6]
        if not s_act and not f_act:
            continue
        dis = [["L2_bygroup", "L2_byindex"][rng.integers(2)] for _ in f_act]
        pd = [int(rng.integers(1, min(6, lens[j]) + 1)) for j in f_act]
        bas = [["B-splines", "PCA"][rng.integers(2)] for _ in f_act]
        ker = ["gauss", "matern5_2", "matern3_2"][rng.integers(3)]
        ants.append(F.encode_ant(ds, df, s_act, f_act, dis, pd, bas, ker))
    return np.stack(ants)


def factory_draws(ants, seed=11):
    """The presample uniforms of every candidate of the fixed list, drawn in list order from ONE stream (as the sequential
    reference loop would consume R's runif): a candidate's draws do not depend on how the list is sharded."""
    from fungp_b200 import factory as F
    rng = np.random.default_rng(seed)
    ds, df, lens = FACTORY["ds"], len(FACTORY["f_lens"]), FACTORY["f_lens"]
    npre = max(FACTORY["n_presample"], FACTORY["n_starts"])
    return [rng.uniform(size=F.n_lengthscales(a, ds, df, lens) * npre) for a in ants]


def run_factory(ctx, rank, world, n_cand, barrier):
    """fgpm_factory candidate scoring (second headline metric), BASELINE.json configs[3]: n = 1024, a FIXED list of n_cand
    structures sharded over the ranks (strong scaling: the list does not grow with the number of GPUs)."""
    from fungp_b200 import synth
    sIn, fIn, y = synth.make_data(FACTORY["seed"], FACTORY["n"], FACTORY["ds"], FACTORY["f_lens"])
    ants = factory_candidates(n_cand, FACTORY["list_seed"])
    u01 = factory_draws(ants)
    idx = np.arange(rank, n_cand, world)        # round-robin deal: candidate costs vary with the structure, this evens them out
    kw = dict(nugget=FACTORY["nugget"], n_starts=FACTORY["n_starts"], n_presample=FACTORY["n_presample"])
    warm = idx[:32]
    ctx.fit_batch(sIn, fIn, y, ants[warm], [u01[i] for i in warm], **kw)       # warm-up: workspaces, streams, projections
    w0 = [ctx.work_count(i) for i in range(4)]
    l0 = ctx.launch_count()
    barrier()
    t0 = time.perf_counter()
    res = ctx.fit_batch(sIn, fIn, y, ants[idx], [u01[i] for i in idx], **kw)
    barrier()
    dt = time.perf_counter() - t0
    w1 = [ctx.work_count(i) for i in range(4)]
    return dict(seconds=dt, n_cand=n_cand, mine=len(idx), ok=int((res["info"] == 0).sum()), best=float(np.nanmax(res["fitness"])),
                nll_evals=w1[0] - w0[0], loocv=w1[1] - w0[1], launches=ctx.launch_count() - l0,
                first=(int(idx[0]), float(res["fitness"][0]), float(res["nll"][0])), data=(sIn, fIn, y, ants, u01))


def factory_cpu_baseline(data, budget_s=5.0):
    """The oracle (numpy/scipy restatement of fgpm() + Q2loocv, oracle/) on candidates of the same fixed list, with the same
    presample draws, on the host cores: a bounded sample (whole candidates until the time budget is spent -- one oracle
    candidate at n = 1024 takes tens of seconds, so in practice the first candidate of the list)."""
    from oracle import fungp_ref as ref
    from fungp_b200 import factory as F
    sIn, fIn, y, ants, u01 = data

    class _Draws:
        def __init__(self, u):
            self.u = u

        def runif(self, k):
            assert k == self.u.size
            return self.u

    done, t0, fits = 0, time.perf_counter(), []
    for c in range(len(ants)):
        kw = F.decode_ant(ants[c], FACTORY["ds"], len(FACTORY["f_lens"]))
        try:
            m = ref.fgpm(sIn=sIn[:, kw["s_active"]] if kw["s_active"] else None,
                         fIn=[fIn[j] for j in kw["f_active"]] if kw["f_active"] else None, sOut=y, kerType=kw["kerType"],
                         f_disType=kw["f_disType"] or "L2_bygroup", f_pdims=kw["f_pdims"] or 3, f_basType=kw["f_basType"] or "B-splines",
                         nugget=FACTORY["nugget"], n_starts=FACTORY["n_starts"], n_presample=FACTORY["n_presample"], rng=_Draws(u01[c]))
            fits.append(float(ref.q2_loocv(m)))
        except Exception:
            fits.append(float("nan"))
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "candidates/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"the first {done} candidates of the fixed 2000-structure list, full oracle fgpm() (presample 20, scipy L-BFGS-B with "
                      "R-style finite differences, preMats) + Q2loocv each; numpy element-wise passes single-threaded as in R, OpenBLAS all cores",
            "seconds": dt, "q2": fits}


def build_inputs(rank=0):
    """Host arrays of the C3 workload; projection through the library's own dimReduction (fgp_project)."""
    from fungp_b200 import synth
    sIn, fIn, y = synth.make_data(CFG["seed"], CFG["n"], CFG["ds"], CFG["f_lens"])
    return sIn, fIn, y


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import fungp_b200
    from fungp_b200 import synth
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # the data path has no collective (independent work items per GPU): the process group only carries the barrier
        # and the max-over-ranks of three timing scalars, so gloo on host tensors is enough (and prints no banner)
        dist.init_process_group(backend="gloo")
    ctx = fungp_b200.Context(device_ids=[local_rank])
    ctx.set_option("streams", args.streams)
    ctx.set_option("outer", args.outer)
    ctx.set_option("fit_blocking", args.fit_blocking)
    ctx.set_option("graphs", args.graphs)
    ctx.set_option("fit_mode", args.fit_mode)
    ctx.set_option("pdl", args.pdl)
    ctx.set_option("dag", args.dag)
    ctx.set_option("fused_assembly", args.fused)
    if args.fit_lanes > 0:
        ctx.set_option("fit_lanes", args.fit_lanes)
    if args.fit_points > 0:
        ctx.set_option("fit_points", args.fit_points)
    if args.fit_workers > 0:
        ctx.set_option("fit_workers", args.fit_workers)
    sIn, fIn, y = build_inputs()
    n, ker, nug = CFG["n"], CFG["ker"], CFG["nugget"]
    proj = [ctx.project(f, p, CFG["basis"]) for f, p in zip(fIn, CFG["pdims"])]
    coefs = [c for (_, _, c) in proj]
    J = [j for (_, j, _) in proj]
    dis = [CFG["dis"]] * len(fIn)
    P = fungp_b200.Problem(ctx, sIn, coefs, J, dis, y)
    d = P.d
    ub = 2.0 * P.dist_max()                       # setBounds_SF: upper = 2 max(D_l)
    theta0 = CFG["theta_frac"] * ub * (1.0 + 0.003 * rank)
    thetas = synth.fd_points(theta0, np.full(d, 1e-10), ub)        # d x (2d+1)
    B = thetas.shape[1]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident():
        nll, s2, info = P.nll_batch(ker, nug, thetas)
        if info.any() or not np.all(np.isfinite(nll)):
            raise SystemExit(f"bench.py: factorisation failed info={info}")
        return nll

    def step_e2e():
        Pe = fungp_b200.Problem(ctx, sIn, coefs, J, dis, y)
        nll, s2, info = Pe.nll_batch(ker, nug, thetas)
        Pe.close()
        if info.any() or not np.all(np.isfinite(nll)):
            raise SystemExit(f"bench.py: factorisation failed in the end-to-end arm, info={info}")
        return nll

    # ---- device-resident arm
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ctx.mark(0)
    t0 = time.perf_counter()
    dag_ms = []                                   # device time of each step's dataflow-Cholesky launch (events on its stream)
    for _ in range(args.steps):
        nll_last = step_resident()
        dag_ms.append(ctx.last_factor_ms())
    ctx.mark(1)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ctx.elapsed_ms(0, 1)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    # ---- end-to-end arm (host buffers in, scalars out, every step)
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    wall_e2e = time.perf_counter() - t1
    # ---- one evaluation at a time (what optim's fn calls and every line-search step are): look-ahead path
    single_ms = None
    if rank == 0:
        for _ in range(2):
            P.nll_batch(ker, nug, thetas[:, :1])
        ts = time.perf_counter()
        for _ in range(5):
            P.nll_batch(ker, nug, thetas[:, :1])
        single_ms = (time.perf_counter() - ts) / 5 * 1e3

    fac = None
    if args.factory_cands > 0:
        fac = run_factory(ctx, rank, world, args.factory_cands, barrier)
    # max over ranks
    tt = torch.tensor([wall, dev_ms, wall_e2e], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    wall, dev_ms, wall_e2e = [float(x) for x in tt.cpu()]
    if fac is not None:
        ft = torch.tensor([fac["seconds"]], dtype=torch.float64)
        fo = torch.tensor([float(fac["ok"]), float(fac["nll_evals"]), float(fac["loocv"]), float(fac["launches"])], dtype=torch.float64)
        fb = torch.tensor([fac["best"]], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ft, op=dist.ReduceOp.MAX)
            dist.all_reduce(fo, op=dist.ReduceOp.SUM)
            dist.all_reduce(fb, op=dist.ReduceOp.MAX)
        fac.update(seconds=float(ft.item()), ok=int(fo[0].item()), nll_evals=int(fo[1].item()), loocv=int(fo[2].item()),
                   launches=int(fo[3].item()), best=float(fb.item()))
    # ---- the path a single R session uses: ONE process, one context over all N GPUs (fgp_ctx_create(N)); rank 0 scores a
    # slice of the same list through it while the other ranks wait, and checks it against its own per-rank results
    inproc = None
    if world > 1 and fac is not None and args.inproc_cands > 0:
        barrier()
        if rank == 0:
            sInF, fInF, yF, ants, u01 = fac["data"]
            k = min(args.inproc_cands, len(ants))
            ctxN = fungp_b200.Context(device_ids=list(range(world)))
            ctxN.fit_batch(sInF, fInF, yF, ants[:2 * world], u01[:2 * world], nugget=FACTORY["nugget"])
            ti = time.perf_counter()
            rN = ctxN.fit_batch(sInF, fInF, yF, ants[:k], u01[:k], nugget=FACTORY["nugget"], n_starts=FACTORY["n_starts"],
                                n_presample=FACTORY["n_presample"])
            ti = time.perf_counter() - ti
            thN = np.tile(thetas, (1, world))
            nllN, _, _ = fungp_b200.Problem(ctxN, sIn, coefs, J, dis, y).nll_batch(ker, nug, thN)
            same_first = bool(rN["fitness"][0] == fac["first"][1] and rN["nll"][0] == fac["first"][2]) if fac["first"][0] == 0 else None
            inproc = {"devices": ctxN.device_count, "candidates": int(k), "candidates_per_s": k / ti, "fitted_ok": int((rN["info"] == 0).sum()),
                      "first_candidate_bit_identical_to_rank0": same_first,
                      "nll_batch_sharded_bit_identical": bool(np.array_equal(nllN[:B], nll_last) and np.array_equal(nllN[:B], nllN[-B:]))}
            ctxN.close()
        barrier()
    # one 2d+1 batch split over the ranks and gathered on the host (fungp_b200/shard.py: the one-process-per-GPU analogue of
    # the split fgp_nll_batch makes inside a multi-GPU context): must equal rank 0's own evaluation of the whole batch
    sharded_ok = None
    if world > 1:
        from fungp_b200 import shard
        th0 = synth.fd_points(CFG["theta_frac"] * ub, np.full(d, 1e-10), ub)
        sn, ss, si = shard.sharded_nll(P, ker, nug, th0, rank, world)
        if rank == 0:
            fn, fs, fi = P.nll_batch(ker, nug, th0)
            sharded_ok = bool(np.array_equal(sn, fn) and np.array_equal(ss, fs) and np.array_equal(si, fi))
    total_evals = world * B * args.steps
    # the step ends with a host-synchronous read of its results, so wall time >= device time; report the wall clock
    value = total_evals / wall
    e2e_value = total_evals / wall_e2e
    h2d = 8 * (n * CFG["ds"] + sum(c.size for c in coefs) + sum(j.size for j in J) + n) + 8 * thetas.size
    d2h = B * (8 + 8 + 4)

    out = None
    if rank == 0:
        # ---- FP64 denominator.  DMMA and DFMA share one FP64 datapath on sm_100a (profiles/r01_peaks.md), so it is the
        # best of both register-resident loops over three attempts (a single attempt right after the timed region has come
        # out 20 % low); never below this repo's own 4-second sustained measurement (37.08, tools/dmma_sustained.cu) -- a low
        # denominator would overstate the fraction
        attempts = []
        for _ in range(3):
            attempts.append(ctx.measure_peaks())
            time.sleep(0.3)
        peaks = dict(attempts[0])
        peaks["dmma_tflops"] = max(a["dmma_tflops"] for a in attempts)
        peaks["dfma_tflops"] = max(a["dfma_tflops"] for a in attempts)
        peaks["copy_gbs"] = max(a["copy_gbs"] for a in attempts)
        peaks["dmma_attempts"] = [a["dmma_tflops"] for a in attempts]
        live = max(peaks["dmma_tflops"], peaks["dfma_tflops"])
        fp64_peak = max(live, 37.08)
        peak_src = ("measured live: best of register-resident DMMA / DFMA loops over 3 attempts (fgp_measure_peaks)" if live >= 37.08 else
                    "this repo's sustained DMMA measurement on this pool (37.08, profiles/r01_peaks.md); the live loops gave %.2f" % live)
        # per-launch CUDA events (profile mode: one stream, two single evaluations) -- explains the step, is not the step
        ctx.set_option("profile", 1)
        P.nll_batch(ker, nug, thetas)
        P.nll_batch(ker, nug, thetas)
        tm = ctx.timing()
        ctx.set_option("profile", 0)
        upd_tf = tm["update_flops"] / (tm["update_ms"] * 1e-3) / 1e12
        asm_bytes = B * 8.0 * n * (n + 1) / 2          # the B matrices of the profiled call: one batch-fused assembly launch
        asm_gbs = asm_bytes / (tm["assembly_ms"] * 1e-3) / 1e9
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            hbm_peak = 6650.0
        traffic, traffic_note = None, None
        dag_on = bool(dag_ms) and min(dag_ms) > 0
        for name in (("r02_traffic_dag.json",) if dag_on else ()) + ("r02_traffic.json", "r01_traffic.json"):
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", name)))
                if name == "r02_traffic_dag.json":
                    traffic = tj["dram_bytes_per_launch"]
                    traffic_note = ("%s: ncu --set full capture of one chol_dag_kernel launch of this workload (13 matrices): %.2f GB read + "
                                    "%.2f GB written, algorithmic %.2f GB, %.1f TFLOP/s and DMMA pipe %.1f %% active under ncu"
                                    % (name, tj["dram_bytes_read"] / 1e9, tj["dram_bytes_write"] / 1e9, tj["algorithmic_bytes_per_launch"] / 1e9,
                                       tj["tflops_under_ncu"], tj["dmma_pipe_active_pct"]))
                else:
                    traffic = tj["update_kernel_dram_bytes_per_launch"]
                    traffic_note = ("%s: ncu capture of the K=%d top-level update launch of the launch-per-step engine: %.1f GFLOP, %.1f MB "
                                    "algorithmic, %.1f TFLOP/s and DMMA pipe %.1f %% active under ncu"
                                    % (name, tj.get("K", 2048), tj["flops_per_launch"] / 1e9, tj["algorithmic_bytes_per_launch"] / 1e6,
                                       tj["tflops_under_ncu"], tj["dmma_pipe_active_pct"]))
                break
            except Exception:
                pass
        chol_flops = n ** 3 / 3.0
        step_tf = chol_flops * total_evals / wall / world / 1e12      # per GPU, whole step, everything else included in the time
        dag_avg = float(np.mean(dag_ms)) if dag_ms and min(dag_ms) > 0 else None
        if dag_avg:
            # the dominant kernel is ONE launch per step: chol_dag_kernel factors the step's 13 matrices (tile tasks: left-looking
            # DMMA accumulation through the TMA ring + the diagonal-block / triangular-solve epilogues)
            kern_tf = chol_flops * B / (dag_avg * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": "chol_dag_kernel (the whole blocked Cholesky of the step's 13 matrices as one persistent "
                                                     "kernel: FP64 DMMA m8n8k4, operands by TMA, tile-ready flags between CTAs)",
                        "achieved": kern_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": kern_tf / fp64_peak,
                        "launch_ms_avg": dag_avg, "launches_timed": len(dag_ms),
                        "flops_per_launch": chol_flops * B,
                        "note": "algorithmic flops (13 n^3/3) over the kernel's own duration, CUDA events on its stream inside the timed region",
                        "whole_step": {"achieved": step_tf, "frac": step_tf / fp64_peak,
                                       "note": "the same flops over the whole timed step (assembly, reductions, copies, host included)"},
                        "traffic": traffic, "traffic_note": traffic_note,
                        "peak_source": peak_src + "; MEASURED_PEAKS.json has no FP64 entry; nominal B200 FP64 = 37.2 TFLOP/s at 1965 MHz"}
        else:
            roofline = {"bound": "tensor", "kernel": "gemm_dmma_tma_kernel (trailing update of the blocked Cholesky: FP64 DMMA m8n8k4, "
                                                     "operands by TMA) -- measured as the whole timed step",
                        "achieved": step_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": step_tf / fp64_peak,
                        "flops_per_step_per_gpu": chol_flops * B,
                        "traffic": traffic, "traffic_note": traffic_note,
                        "peak_source": peak_src + "; MEASURED_PEAKS.json has no FP64 entry; nominal B200 FP64 = 37.2 TFLOP/s at 1965 MHz"}
        roofline["update_kernel_profiled"] = {
            "note": "launch-per-step engine (option dag 0, profile mode): per-launch CUDA events, the step's 13 points as one batch on ONE stream; "
                    "flops count the triangle of diagonal tiles",
            "tflops": upd_tf, "frac": upd_tf / fp64_peak,
            "flops_per_launch_avg": tm["update_flops"] / max(1, tm["update_launches"]),
            "launch_ms_avg": tm["update_ms"] / max(1, tm["update_launches"])}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[2] sizes: fgpm likelihood at n=8192, 4 scalar + 2 functional "
                                   "(len 100, B-splines p=5, L2_bygroup), gauss kernel, nugget 1e-8; one step = the 13 points of one "
                                   "finite-difference gradient (2d+1, d=6) in one fgp_nll_batch call",
                       "n": n, "d": int(d), "evals_per_step_per_gpu": int(B), "l2_policy": "inputs larger than L2: every evaluation "
                       "writes and factors its own 512 MiB matrix", "streams": args.streams, "outer_panel_tiles": args.outer,
                       "parallelism": f"theta-points sharded, {world} GPU(s), no collective"},
            "device_ms_per_step": dev_ms / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": wall_e2e / args.steps * 1e3},
            "gpu_launches": int(launches),
            "clocks": clocks,
            # achieved = ALGORITHMIC flops of the blocked Cholesky (n^3/3 per evaluation) over the timed step itself: assembly,
            # diagonal blocks, panel solves and reductions are all inside the time, so this can only understate the kernel
            "roofline": roofline,
            "single_eval": {"ms": single_ms, "tflops": chol_flops / (single_ms * 1e-3) / 1e12, "frac": chol_flops / (single_ms * 1e-3) / 1e12 / fp64_peak,
                            "note": "one fgp_nll_batch call with B = 1, host wall clock, mean of 5"},
            "roofline_assembly": {"bound": "hbm", "kernel": "assemble_fused_kernel<gauss> (distance + correlation, lower triangles of the "
                                                            "step's 13 matrices in one launch)",
                                  "achieved": asm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": asm_gbs / hbm_peak,
                                  "algorithmic_bytes_per_launch": asm_bytes, "launch_ms": tm["assembly_ms"],
                                  "bytes_per_matrix": asm_bytes / B, "ms_per_matrix": tm["assembly_ms"] / B,
                                  "note": "FP64-issue bound: DMMA and DFMA share one datapath (DESIGN.md 2.3)"},
            "kernel_ms_profiled_13evals_1stream_launch_per_step_engine": {k: tm[k] for k in ("assembly_ms", "factor_ms", "update_ms", "potf2_ms", "trsm_ms")},
            "fp64_peaks_measured": peaks,
            "nll_check": float(nll_last[0]),
        }
        if sharded_ok is not None:
            out["sharded_batch_bit_identical"] = sharded_ok
        if fac is not None:
            nf = FACTORY["n"]
            agg_flops = fac["nll_evals"] * nf ** 3 / 3.0 + fac["loocv"] * 2.0 * nf ** 3 / 3.0
            agg_tf = agg_flops / fac["seconds"] / 1e12
            out["factory"] = {"metric": "fgpm_factory candidates/sec", "value": fac["n_cand"] / fac["seconds"], "unit": "candidates/s",
                              "candidates": fac["n_cand"], "fitted_ok": fac["ok"], "seconds": fac["seconds"], "n_gpus": world,
                              "scaling": "strong", "best_q2": fac["best"],
                              "likelihood_evaluations": fac["nll_evals"], "loocv_factorisations": fac["loocv"],
                              "gpu_launches": fac["launches"],
                              "roofline": {"bound": "tensor", "achieved": agg_tf, "peak": fp64_peak * world, "unit": "TFLOP/s",
                                           "frac": agg_tf / (fp64_peak * world),
                                           "note": "aggregate algorithmic FP64 flops: likelihood evaluations x n^3/3 + leave-one-out "
                                                   "factorisations x 2 n^3/3, over the whole call (host optimiser, projections included)"},
                              "config": "BASELINE.json configs[3]: n=1024, 5 scalar + 3 functional (len 10/22/50), the fixed pre-generated list "
                                        "of 2000 candidate structures (seed 7) sharded round-robin over the ranks; per candidate: presample 20, "
                                        "L-BFGS-B (lock-step over all in-flight candidates, fgp_fit_batch), Q2loocv; no collective"}
            if inproc is not None:
                out["factory"]["in_process_multi_gpu"] = inproc
            if world == 1 and not args.skip_cpu:
                out["factory"]["cpu_baseline"] = factory_cpu_baseline(fac["data"])
        if world == 1 and not args.skip_cpu:
            out["cpu_baseline"] = cpu_baseline(sIn, coefs, J, y, thetas[:, 0], nll_last[0])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    P.close()
    ctx.close()
    if rank == 0:
        print(json.dumps(out))


# ------------------------------------------------------------------------------------------------------------------
def _oracle_lists(sIn, coefs, J, rows=None):
    from oracle import fungp_ref as ref
    s1 = sIn if rows is None else sIn[:rows]
    c1 = coefs if rows is None else [c[:rows] for c in coefs]
    Ms = ref.setDistMatrix_S(s1, sIn) + ref.setDistMatrix_F(c1, coefs, J, [CFG["dis"]] * len(coefs))
    return Ms


def cpu_baseline(sIn, coefs, J, y, theta, nll_gpu):
    """The oracle (literal numpy/scipy restatement of the reference's operation sequence) on the host cores: ONE
    likelihood evaluation at n = 8192 with the materialised distance lists (built once, untimed, as fgpm() does)."""
    from oracle import fungp_ref as ref
    Ms = _oracle_lists(sIn, coefs, J)
    t0 = time.perf_counter()
    nll, s2 = ref.negLogLik(theta, [], Ms, y, CFG["ker"], CFG["nugget"])
    dt = time.perf_counter() - t0
    return {"value": 1.0 / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": "1 full evaluation at n=8192 (setR over 6 materialised 8192^2 distance matrices, dpotrf, backsolve); numpy "
                      "element-wise passes are single-threaded as in R, OpenBLAS uses all cores",
            "seconds": dt, "nll": float(nll), "nll_rel_diff_vs_gpu": abs(nll - nll_gpu) / max(abs(nll), CFG["n"])}


def run_reference(args):
    """Reference arm: the CPU restatement of the reference (oracle/) on this box's host cores, rank 0 only.
    Each step = a bounded sample of one evaluation: setR on the first f*n rows of every distance matrix (timed, scaled
    by 1/f) + the full dpotrf / backsolve / log det on the assembled 8192^2 matrix (timed in full)."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import scipy.linalg as sla
    from oracle import fungp_ref as ref
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm is to use every host thread it can
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    sIn, fIn, y = build_inputs()
    bcj = ref.dimReduction(fIn, CFG["pdims"], [CFG["basis"]] * len(fIn))
    J, coefs = bcj["J"], bcj["coefs"]
    n = CFG["n"]
    budget = 150.0 / max(1, args.steps + args.warmup)
    # full matrix once (untimed) for the factorisation part
    Ms_full = _oracle_lists(sIn, coefs, J)
    ub = np.array([2.0 * M.max() for M in Ms_full])
    theta = CFG["theta_frac"] * ub
    R = ref.corr_matrix(theta, Ms_full, CFG["ker"], CFG["nugget"])
    del Ms_full

    def chol_part():
        t0 = time.perf_counter()
        L = ref.r_chol_lower(R)
        s2 = ref.analyticVar_llik(L, y, n)
        ld = float(np.sum(np.log(np.diag(L))))
        return time.perf_counter() - t0, 0.5 * (n * np.log(2 * np.pi * s2) + 2 * ld + n)
    t_chol, nll = chol_part()
    rows_probe = 256
    Msp = _oracle_lists(sIn, coefs, J, rows_probe)
    t0 = time.perf_counter(); ref.setR(theta, Msp, CFG["ker"]); t_probe = time.perf_counter() - t0
    t_asm_full = t_probe * n / rows_probe
    frac = min(1.0, max(0.03, (budget - t_chol) / t_asm_full))
    rows = max(64, int(frac * n))
    Ms = _oracle_lists(sIn, coefs, J, rows)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        Rs = ref.setR(theta, Ms, CFG["ker"])
        Rs = Rs + 0.0   # the '+ diag(nugget)' pass of the reference touches the whole matrix once more
        t_asm = (time.perf_counter() - t0) * n / rows
        tc, nll = chol_part()
        if it >= args.warmup:
            times.append(t_asm + tc)
    sec = float(np.mean(times))
    out = {"impl": "reference", "metric": METRIC, "value": 1.0 / sec, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": "BASELINE.json configs[2] sizes: fgpm likelihood at n=8192, 4 scalar + 2 functional (len 100, "
                                  "B-splines p=5, L2_bygroup), gauss kernel, nugget 1e-8; one step = one evaluation (CPU)", "n": n},
           "cpu_baseline": {"value": 1.0 / sec, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                            "sample": f"per step: setR on the first {rows} of {n} rows of each of the 6 distance matrices (scaled by "
                                      f"{n / rows:.2f}) + full dpotrf/backsolve/logdet on the 8192^2 matrix; R is not installed on the "
                                      "box, so this is the numpy/scipy restatement of the reference (oracle/)"},
           "e2e": {"value": 1.0 / sec, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "nll_check": float(nll)}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--outer", type=int, default=16)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--graphs", type=int, default=0, help="replay small-n likelihood calls as CUDA graphs in the factory fits (0/1)")
    ap.add_argument("--fit-workers", type=int, default=0, help="fit_batch worker contexts per GPU (0 = library default, 8)")
    ap.add_argument("--fit-blocking", type=int, default=-1, help="fit_batch workers: 1 sleep / 0 spin in host waits, -1 auto")
    ap.add_argument("--factory-cands", type=int, default=FACTORY["n_cand"],
                    help="length of the fixed candidate list for the factory metric, sharded over the GPUs (0 = skip)")
    ap.add_argument("--fit-mode", type=int, default=1, help="fgp_fit_batch: 1 lock-step driver (one host thread per GPU), 0 one thread per candidate")
    ap.add_argument("--dag", type=int, default=1, help="dataflow Cholesky: one persistent kernel per likelihood batch (0 = launch-per-step engine)")
    ap.add_argument("--pdl", type=int, default=1, help="programmatic dependent launch along the factorisation chain (0/1)")
    ap.add_argument("--fused", type=int, default=1, help="batch-fused assembly kernel (0/1)")
    ap.add_argument("--fit-lanes", type=int, default=0, help="lock-step rounds in flight per GPU (0 = library default)")
    ap.add_argument("--fit-points", type=int, default=0, help="lock-step likelihood points per round and lane (0 = library default)")
    ap.add_argument("--inproc-cands", type=int, default=256, help="N > 1: candidates rank 0 also scores through ONE context over all N GPUs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
