#!/usr/bin/env python
"""bench.py -- headline benchmark of the GP-BO hot path (BASELINE.json metric).

metric   EI candidates scored/sec at n=4096, d=20 (whole job, all GPUs)
step     one pass of the hot path over one candidate pool per GPU:
         posterior mean/variance -> EI -> top-k (+ one all-gather of per-GPU top-k when N > 1)
value    pool already resident in HBM when the timed region starts
e2e      same metric through the C-ABI host entry point (bbo_gp_score_pool_host): the pool lives
         in pinned HOST memory; H2D of the pool and D2H of the selection are inside the timed region
also     "fit": GP fit ms/iter (value + analytic gradient, FP64) next to cuSOLVER and the CPU s/iter;
         "roofline" for the dominant kernel; "cpu_baseline" (the UNMODIFIED reference on the host cores)

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision tf32+refine|tf32|fp64] [--impl reference]
                  [--scaling weak|strong --pool-total Q] [--nobs N --dim D --objective ackley|levy|rastrigin]
For N > 1 launch with torchrun (one rank per GPU); rank 0 prints ONE JSON line.

precision  "tf32+refine" (default): tcgen05 TF32 panel over the whole pool keeps a short-list, the short-list is
           re-scored in float64, the selection and the returned scores are float64 (identical to --precision fp64:
           checked in the run, `selection_matches_fp64`).  "tf32": TF32-accurate scores.  "fp64": DMMA throughout.
scaling    "weak" (default): --pool candidates per GPU.  "strong": --pool-total candidates split over the GPUs
           (BASELINE config 5: --scaling strong --pool-total 67108864 --nobs 8192 --dim 100 --objective levy).
reference  --impl reference runs the UNMODIFIED reference (vendored to baseline/_ref by
           baseline/vendor_reference.py): GP.predict in chunks -> .diagonal() -> EI -> top-k
           (bbo/algorithms/surrogates/gp/gp.py:65-83, designers/bo/acquisitions.py:18-23) on all host cores, on a
           bounded sample of the same pool; the oracle port is timed beside it (`cpu_port`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_OBS, DIM, TOPK = 4096, 20, 8
NOISE, LENGTHSCALE = 1e-4, 0.5
SEED = 20251019


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="tf32+refine", choices=["fp64", "tf32", "tf32+refine"],
                    help="arithmetic of the n^2 panel in the headline number (fp64 is reported beside it)")
    ap.add_argument("--no-other-mode", action="store_true")
    # (--nobs / --dim: torchrun's own parser rejects a script argument spelled --n as an ambiguous abbreviation)
    ap.add_argument("--nobs", "--n", dest="n", type=int, default=N_OBS)
    ap.add_argument("--dim", "--d", dest="d", type=int, default=DIM)
    ap.add_argument("--objective", default="ackley", choices=["ackley", "levy", "rastrigin"])
    ap.add_argument("--lengthscale", type=float, default=None, help="default 0.5 (1.5 for d >= 50)")
    ap.add_argument("--pool", type=int, default=1 << 20, help="candidates per GPU per step (weak scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--pool-total", type=int, default=1 << 26, help="candidates per step, all GPUs (strong scaling)")
    ap.add_argument("--cpu-sample", type=int, default=4096, help="candidates per CPU-baseline step")
    ap.add_argument("--no-fit", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the single-GPU re-score of the global pool")
    a = ap.parse_args()
    if a.lengthscale is None:
        a.lengthscale = LENGTHSCALE if a.d < 50 else 1.5
    return a


def world_size(a):
    return int(os.environ.get("WORLD_SIZE", str(max(1, a.gpus)))) if a.impl == "b200" else max(1, a.gpus)


def pool_per_gpu(a, world):
    return a.pool if a.scaling == "weak" else -(-a.pool_total // world)


def config_dict(a, world):
    """The same dict from both arms (the driver compares them)."""
    per = pool_per_gpu(a, world)
    total = per * world if a.scaling == "weak" else a.pool_total
    return {"workload": (f"GP-BO EI pool scoring: {a.objective.capitalize()}-{a.d}D, n={a.n} observations, ARD RBF "
                         f"(l={a.lengthscale:g}, s2=1, noise=1e-4), uniform candidate pool, top-{TOPK}"),
            "n": a.n, "d": a.d, "k": TOPK, "pool_per_gpu": per, "pool_total": total, "scaling": a.scaling,
            "parallelism": f"pool sharded x{world}, 1 all-gather of top-k",
            "l2": "inputs larger than L2 (pool + L^-1 per pass)"}


def problem(a):
    from bbo_b200.benchmarks import Ackley, Levy, Rastrigin, make_observations
    func = {"ackley": Ackley, "levy": Levy, "rastrigin": Rastrigin}[a.objective](a.d)
    X, y = make_observations(func, a.n, SEED)
    raw_ls = float(np.log(np.expm1(a.lengthscale)))      # softplus^-1
    raw_var = float(np.log(np.expm1(1.0)))
    return X, y, raw_ls, raw_var


# ------------------------------------------------------------------------------------------
# CPU arms: the unmodified reference (baseline/_ref) and the oracle port
# ------------------------------------------------------------------------------------------
def cpu_pool_sample(a, count):
    rng = np.random.default_rng(SEED + 1)
    return rng.random((count, a.d))


def reference_available():
    from baseline.ref_env import reference_path
    return reference_path() is not None


class ReferenceArm:
    """The reference's own classes on this host's cores, float64, fixed hyper-parameters."""

    def __init__(self, a):
        from baseline.ref_env import setup_reference
        self.path = setup_reference()
        from bbo.algorithms.designers.bo.acquisitions import EI
        from bbo.algorithms.surrogates.gp import GP
        from bbo.algorithms.surrogates.gp.kernels import RBFKernel, ScaleKernel
        from bbo.algorithms.surrogates.gp.means import ConstantMean
        from bbo.algorithms.surrogates.gp.objectives import neg_log_marginal_likelihood
        X, y, raw_ls, raw_var = problem(a)
        self.a = a
        self.X = torch.as_tensor(X)
        self.Y = torch.as_tensor(y).reshape(-1, 1)
        self.cov = ScaleKernel(RBFKernel(ard_dim=a.d)).double()
        self.mean = ConstantMean().double()
        with torch.no_grad():
            self.cov.base_kernel.lengthscale.fill_(raw_ls)
            self.cov.variance.fill_(raw_var)
        self.gp = GP(self.X, self.Y, self.mean, self.cov, noise_variance=NOISE, epochs=0)
        self.ei = EI(float(y.max()))
        self.nll = neg_log_marginal_likelihood

    def score(self, Xq: torch.Tensor, chunk: int):
        """GP.predict (gp.py:65-83) in 2-D chunks -> .diagonal() -> EI (acquisitions.py:18-23) -> top-k."""
        out = []
        with torch.no_grad():
            for c0 in range(0, Xq.shape[0], chunk):
                mu, cov = self.gp.predict(Xq[c0:c0 + chunk])
                out.append(self.ei(mu.reshape(-1), cov.diagonal().sqrt()))
        sc = torch.cat(out)
        return torch.topk(sc, min(TOPK, sc.numel()))

    def fit_iter(self):
        """objectives.py:34-45 + backward (gp.py:58-62): one fit iteration."""
        for p in list(self.mean.parameters()) + list(self.cov.parameters()):
            p.grad = None
        loss = self.nll(self.X, self.Y, self.mean, self.cov, NOISE)
        loss.backward()
        return float(loss)


class PortArm:
    def __init__(self, a):
        from oracle import gp_oracle as O
        X, y, raw_ls, raw_var = problem(a)
        self.O, self.X, self.y = O, X, y
        self.p = O.GPParams(kind=O.RBF, raw_lengthscale=np.full(a.d, raw_ls), raw_variance=raw_var, constant=0.0,
                            noise=NOISE)
        self.state = O.solve_gp_linear_system(self.p, X, y)
        self.best_f = float(y.max())

    def score(self, Xq, chunk):
        return self.O.score_pool(self.p, self.X, self.y, Xq, "ei", self.best_f, k=TOPK, state=self.state, chunk=chunk)

    def fit_iter(self):
        return self.O.objective_value_and_grad(self.p, self.X, self.y)


def cpu_rate(arm, Xq, steps, warmup, chunks=(64, 256, 1024)):
    """(cand/s, s/step, chunk): quick sweep of the chunk size on a slice, then `warmup` + `steps` timed passes."""
    probe = Xq[:min(len(Xq), 1024)]
    best = None
    arm.score(probe[:64], 64)            # the first predict caches the factorisation (gp.py:66-74)
    for chunk in chunks:
        t0 = time.perf_counter()
        arm.score(probe, chunk)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, chunk)
    chunk = best[1]
    for _ in range(warmup):
        arm.score(Xq, chunk)
    t0 = time.perf_counter()
    for _ in range(steps):
        arm.score(Xq, chunk)
    dt = (time.perf_counter() - t0) / steps
    return len(Xq) / dt, dt, chunk


def emit_line(line):   # rebound by main() to the saved stdout descriptor
    print(json.dumps(line), flush=True)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    world = world_size(a)
    use_ref = reference_available()
    arm = ReferenceArm(a) if use_ref else PortArm(a)
    Xq_np = cpu_pool_sample(a, a.cpu_sample)
    Xq = torch.as_tensor(Xq_np) if use_ref else Xq_np
    rate, dt, chunk = cpu_rate(arm, Xq, a.steps, min(a.warmup, 1))
    cores = torch.get_num_threads()
    kind = "reference" if use_ref else "port"
    what = ("unmodified reference (baseline/_ref): GP.predict 2-D chunks -> diagonal -> EI -> topk, torch CPU float64"
            if use_ref else "oracle port (torch CPU float64); baseline/_ref absent")
    sample = f"{a.cpu_sample} candidates/step of the same pool, chunk {chunk}, {what}"
    line = {
        "impl": "reference", "metric": "EI candidates scored/sec", "value": rate, "unit": "candidates/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(a, world),
        "cpu_baseline": {"value": rate, "unit": "candidates/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and clocks-event reasons of one GPU sampled DURING the timed region: NVML from a thread every
    ~2 ms (nvidia-smi -lms needs ~100 ms to start and 50 ms per sample, the timed region of a default run is ~100 ms);
    nvidia-smi is the fallback when NVML cannot be loaded."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []          # (sm_mhz, max_mhz, watts, set of reasons)
        self.proc = None
        self.nvml = None
        self.stop_flag = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[self.index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:  # noqa: BLE001
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _poll(self):
        nv = self.nvml
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        mx = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                self.rows.append((sm, mx, pw, {k for k, b in bits.items() if mask & b}))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) < 7:
                continue
            try:
                self.rows.append((float(r[0]), float(r[1]), float(r[2]),
                                  {nm for nm, v in zip(names, r[3:7]) if v.lower().startswith("active")}))
            except ValueError:
                continue

    def stop(self):
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML, no nvidia-smi"]}
        self.stop_flag.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()
        else:
            self.t.join(timeout=1)
        rows = list(self.rows)
        sm = [r[0] for r in rows]
        reasons = set()
        for r in rows:
            reasons |= r[3]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": max(r[1] for r in rows) if rows else None,
                "power_w_max": max(r[2] for r in rows) if rows else None, "samples": len(rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi", "reasons": sorted(reasons)}


def load_json(path):
    try:
        return json.load(open(path))
    except Exception:  # noqa: BLE001
        return None


def stored_peaks():
    """profiles/measured_peaks_tf32_f64.json: cuBLAS TF32 / FP64 GEMM peaks measured ONCE on this pool's B200 with the
    method of MEASURED_PEAKS.json (tools/measure_peaks.py) -- the driver's file has bf16 only.  A stored denominator
    keeps roofline.frac comparable from run to run (a live one moved 708 -> 758 TFLOP/s across round 1's runs)."""
    return load_json(os.path.join(ROOT, "profiles", "measured_peaks_tf32_f64.json"))


def live_matmul_peak(dtype, tf32, n=4096, reps=10):
    """cuBLAS GEMM throughput, best of `reps`, CUDA events (fallback when the stored peaks are absent)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, dtype=dtype, device="cuda")
    b = torch.randn(n, n, dtype=dtype, device="cuda")
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = old
    del a, b
    return 2.0 * n ** 3 / best / 1e9   # TFLOP/s


def peaks_for(tag):
    """(burst, sustained, source) TFLOP/s for the panel dtype.  The tensor-core scoring panel multiplies binary16
    operands (TF32's 10-bit mantissa in a 16-bit container) into float32 accumulators, i.e. it runs at the dense 16-bit
    rate: its denominator is the driver's MEASURED_PEAKS.json (cuBLAS bf16 GEMM), else the same measurement stored in
    profiles/.  The float64 kernels use the stored cuBLAS f64 GEMM peak; a live measurement is the last resort."""
    sp = stored_peaks()
    if tag != "vnorm":
        mp = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"))
        if mp and "bf16_tflops" in mp:
            return mp["bf16_tflops"], mp.get("bf16_tflops_sustained"), \
                f"MEASURED_PEAKS.json dense bf16 ({mp.get('how', 'driver-written')})"
        if sp and "bf16_tflops" in sp:
            return sp["bf16_tflops"], sp.get("bf16_tflops_sustained"), \
                f"profiles/measured_peaks_tf32_f64.json dense bf16 ({sp.get('how', '')})"
        return live_matmul_peak(torch.bfloat16, False, n=8192), None, "measured live: cuBLAS bf16 GEMM 8192^3 best of 10"
    key = "f64"
    if sp and f"{key}_tflops" in sp:
        return sp[f"{key}_tflops"], sp.get(f"{key}_tflops_sustained"), \
            f"profiles/measured_peaks_tf32_f64.json ({sp.get('how', 'cuBLAS GEMM, MEASURED_PEAKS.json method')})"
    return live_matmul_peak(torch.float64, False), None, "measured live: cuBLAS f64 GEMM 4096^3 best of 10"


def cusolver_fit_ms(n, reps=3):
    """Library bar for the fit's dense part: cuSOLVER potrf + potri through torch.linalg on an SPD n x n float64
    matrix (the Cholesky and the K^-1 the gradient needs; the fit also builds K and reduces the gradient)."""
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)
    K = A @ A.T / n + torch.eye(n, dtype=torch.float64, device="cuda")
    del A
    for _ in range(2):
        L = torch.linalg.cholesky(K)
        torch.cholesky_inverse(L)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        L = torch.linalg.cholesky(K)
        torch.cholesky_inverse(L)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run_b200(a):
    import torch.distributed as dist
    import bbo_b200 as B
    from bbo_b200 import _lib
    from bbo_b200.distributed import allgather_merge_topk, broadcast_state, shard_range
    from bbo_b200.pool import uniform_pool

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; bbo_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    X, y, raw_ls, raw_var = problem(a)
    cov = B.ScaleKernel(B.RBFKernel(ard_dim=a.d)).double()
    with torch.no_grad():
        cov.base_kernel.lengthscale.fill_(raw_ls)
        cov.variance.fill_(raw_var)
    gp = B.CudaGP(torch.as_tensor(X), torch.as_tensor(y).reshape(-1, 1), B.ConstantMean().double(), cov,
                  device=dev, noise_variance=NOISE, epochs=1)
    bc0, bc1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    bc0.record()
    if world > 1:
        broadcast_state(gp, src=0)       # hyper-parameters from the fitting rank; every rank factorises once
    gp.state
    bc1.record()
    barrier()
    state_ms = max_over_ranks(bc0.elapsed_time(bc1))
    acq = B.EI(float(y.max()))
    if a.scaling == "weak":
        q = a.pool
        lo = rank * q
        q_total = world * q
    else:
        q_total = a.pool_total
        lo, hi = shard_range(q_total, rank, world)
        q = hi - lo
    # device-resident pool (float64); when the shard does not fit comfortably it is generated in slabs per step
    slab = min(q, max(1 << 20, (2 << 30) // (8 * a.d)))
    resident = q <= slab
    Xq = uniform_pool(SEED + 1, lo, min(q, slab), a.d, device=dev)
    lib = _lib.load()
    coll_ev = []

    def score_shard(precision):
        """This rank's shard -> (scores[k], indices[k]) on the device (slabs merged through `running`)."""
        if resident:
            return gp.score_pool(Xq, acq, TOPK, precision=precision, index_offset=lo)
        run = None
        for s0 in range(0, q, slab):
            qs = min(slab, q - s0)
            uniform_pool(SEED + 1, lo + s0, qs, a.d, device=dev, out=Xq[:qs])
            run = gp.score_pool(Xq[:qs], acq, TOPK, precision=precision, index_offset=lo + s0, running=run)
        return run

    def step(precision=a.precision, timed=True):
        s, i = score_shard(precision)
        if world == 1:
            return s, i
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        out = allgather_merge_topk(s, i, TOPK)           # pack -> ONE all_gather_into_tensor -> merge
        c1.record()
        if timed:
            coll_ev.append((c0, c1))
        return out

    for _ in range(a.warmup):
        step(timed=False)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = lib.bbo_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        top_s, top_i = step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.bbo_launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    if clk is not None:
        # context for the reader: what the driver's own peak measurement saw on this pool under tensor load
        mpk = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json")) or {}
        cul = mpk.get("clocks_under_load")
        if cul:
            clk["measured_peaks_clocks_under_load"] = {k: cul.get(k) for k in ("sm_mhz_median", "sm_max_mhz",
                                                                              "power_w_max", "throttled")}
    value = q_total * a.steps / (ms * 1e-3)
    # collective cost: `collective_in_step_ms` is what the events around pack / all-gather / merge saw inside the timed
    # steps -- it includes waiting for the slowest rank to ARRIVE (per-chip power caps make the ranks' steps differ by
    # a few hundred microseconds); `collective_ms` is the same three operations timed back to back after a barrier
    collective_ms = collective_in_step_ms = None
    if world > 1:
        collective_in_step_ms = max_over_ranks(sum(c0.elapsed_time(c1) for c0, c1 in coll_ev) / max(len(coll_ev), 1))
        for _ in range(3):
            allgather_merge_topk(top_s, top_i, TOPK)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(50):
            allgather_merge_topk(top_s, top_i, TOPK)
        c1.record()
        torch.cuda.synchronize()
        collective_ms = max_over_ranks(c0.elapsed_time(c1) / 50)
    cert = gp.refine_certificate() if a.precision == "tf32+refine" else None

    # ---- e2e: pool in pinned host memory (float32, the dtype of the reference's features), through
    #      bbo_gp_score_pool_host: H2D of every candidate and D2H of the selection inside the timed region
    e2e = None
    if not a.no_e2e:
        qh = min(q, slab)
        Xh = torch.empty(qh, a.d, dtype=torch.float32).pin_memory()
        if not resident:
            uniform_pool(SEED + 1, lo, qh, a.d, device=dev, out=Xq[:qh])
        Xh.copy_(Xq[:qh])
        torch.cuda.synchronize()

        def step_host():
            hs, hi = gp.score_pool(Xh, acq, TOPK, precision=a.precision, index_offset=lo, host_pool=True)
            if world > 1:   # same single all-gather + merge on the gathered lists
                ms_, mi_ = allgather_merge_topk(hs.to(dev), hi.to(dev), TOPK)
                return ms_.cpu(), mi_.cpu()
            return hs, hi

        for _ in range(min(a.warmup, 3)):
            step_host()
        barrier()
        e0.record()
        for _ in range(a.steps):
            hs, hi = step_host()
        e1.record()
        barrier()
        ms_h = max_over_ranks(e0.elapsed_time(e1))
        # the device run scored the same candidates in float64; compare selections on the float32 pool
        ds, di = gp.score_pool(Xh.to(dev), acq, TOPK, precision="fp64", index_offset=lo)
        if world > 1:
            ds, di = allgather_merge_topk(ds, di, TOPK)
        e2e = {"value": world * qh * a.steps / (ms_h * 1e-3), "unit": "candidates/s",
               "h2d_bytes_per_step": qh * a.d * 4, "d2h_bytes_per_step": TOPK * 16,
               "ms_per_step": ms_h / a.steps, "host_pool_dtype": "float32", "candidates_per_gpu_per_step": qh,
               "selection_matches_fp64": bool(torch.equal(hi.cpu(), di.cpu()))}
        del Xh

    # ---- float64 mode on the same candidates: the parity anchor of the headline mode, and its throughput.  When the
    #      shard is generated in slabs (strong scaling at config-5 sizes) the comparison runs on its first slab.
    other = None
    sel_match = None
    if not a.no_other_mode and a.precision != "fp64":
        qs = min(q, slab)
        if not resident:
            uniform_pool(SEED + 1, lo, qs, a.d, device=dev, out=Xq[:qs])

        def sub_step(precision):
            s_, i_ = gp.score_pool(Xq[:qs], acq, TOPK, precision=precision, index_offset=lo)
            if world > 1:
                return allgather_merge_topk(s_, i_, TOPK)
            return s_, i_

        hs_, hi_ = sub_step(a.precision)
        osteps = 2
        sub_step("fp64")
        barrier()
        e0.record()
        for _ in range(osteps):
            os_, oi_ = sub_step("fp64")
        e1.record()
        barrier()
        oms = max_over_ranks(e0.elapsed_time(e1))
        sel_match = bool(torch.equal(oi_, hi_))
        other = {"precision": "fp64", "value": world * qs * osteps / (oms * 1e-3), "unit": "candidates/s",
                 "ms_per_step": oms / osteps, "steps": osteps, "candidates_per_gpu_per_step": qs,
                 "max_rel_score_diff_vs_headline": float(((os_ - hs_).abs() / os_.abs().clamp_min(1e-300)).max())}

    # ---- N > 1: the sharded selection against ONE GPU scoring the whole pool alone (after the timed region)
    single = None
    if world > 1 and not a.no_verify and rank == 0:
        run = None
        for s0 in range(0, q_total, slab):
            qs = min(slab, q_total - s0)
            buf = uniform_pool(SEED + 1, s0, qs, a.d, device=dev)
            run = gp.score_pool(buf, acq, TOPK, precision=a.precision, index_offset=s0, running=run)
            del buf
        single = bool(torch.equal(run[1], top_i) and torch.equal(run[0], top_s))

    roof = fit = cpu = cpu_port = None
    if rank == 0:
        # ---- roofline of the dominant kernel: separate profiled pass (CUDA events on the launch stream)
        tag = "vnorm" if a.precision == "fp64" else "tf32"
        npasses = max(1, min(a.steps, 3))
        qprof = min(q, slab)
        lib.bbo_profile_enable(1)
        for _ in range(npasses):
            gp.score_pool(Xq[:qprof], acq, TOPK, precision=a.precision, index_offset=lo)
        regions, tot_ms = _lib.profile_read(tag)
        kregions, ktot = _lib.profile_read("kstar")
        sregions, stot = _lib.profile_read("select")
        if tag == "tf32":
            vregions, vtot = _lib.profile_read("vnorm")       # the float64 re-score of the short-list
        else:
            vregions, vtot = 0, 0.0
        lib.bbo_profile_enable(0)
        flops_per_launch = float(a.n) * a.n * (qprof * npasses / max(regions, 1))   # n^2 per candidate
        avg_ms = tot_ms / max(regions, 1)
        achieved = flops_per_launch / (avg_ms * 1e-3) / 1e12
        peak, peak_sus, peak_src = peaks_for(tag)
        traffic = None
        tr = load_json(os.path.join(ROOT, "profiles", "roofline_traffic.json"))
        if tr:
            traffic = tr.get(tag, {}).get("dram_bytes_per_launch")
        sp = stored_peaks() or {}
        share = tot_ms / max(tot_ms + ktot + stot + vtot, 1e-9)
        roof = {"bound": "tensor",
                "kernel": "k_vnorm (DMMA.8x8x4)" if tag == "vnorm" else "k_vnorm_tf32_2sm (tcgen05.mma cta_group::2 kind::f16, f32 accumulators in TMEM)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "peak_sustained": peak_sus, "frac_of_sustained": (achieved / peak_sus) if peak_sus else None,
                "operands": "f64" if tag == "vnorm" else
                "binary16 containers of 10-bit-mantissa (TF32-width) values: correlations K*/s2 and 2^e L^-1; f32 accumulate",
                "frac_vs_tf32_peak": (achieved / sp["tf32_tflops"]) if (tag == "tf32" and "tf32_tflops" in sp) else None,
                "peak_source": peak_src, "launches_timed": regions, "avg_launch_ms": avg_ms,
                "algorithmic_flops_per_launch": flops_per_launch, "share_of_step": share,
                "other_kernels_ms_per_pass": {"kstar": ktot / npasses, "select": stot / npasses,
                                              "refine_fp64_panel": vtot / npasses}}
        step_flops = (float(a.n) * a.n + 3.0 * a.n * a.d + 4.0 * a.n) * q_total / world
        roof["whole_step_tflops_per_gpu"] = step_flops / (ms / a.steps * 1e-3) / 1e12

        # ---- GP fit s/iter (metric part ii): value + analytic gradient, FP64
        if not a.no_fit:
            for _ in range(3):   # warm-up (the library builds its cached graph on the second identical call)
                gp.value_and_grad()
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nit = 5
            f0.record()
            for _ in range(nit):
                gp.value_and_grad()
            f1.record()
            torch.cuda.synchronize()
            fit_ms = f0.elapsed_time(f1) / nit
            fit_flops = float(a.n) ** 3 + 3.0 * a.n * a.n * a.d
            peak64, _, src64 = peaks_for("vnorm")
            fit = {"ms_per_iter": fit_ms, "n": a.n, "d": a.d, "dtype": "f64",
                   "achieved_tflops": fit_flops / (fit_ms * 1e-3) / 1e12,
                   "algorithmic_flops_per_iter": fit_flops, "peak_tflops": peak64, "peak_source": src64}
            fit["frac"] = fit["achieved_tflops"] / peak64
            try:
                fit["cusolver_potrf_potri_ms"] = cusolver_fit_ms(a.n)
                fit["vs_cusolver"] = fit["cusolver_potrf_potri_ms"] / fit_ms
            except Exception as e:  # noqa: BLE001
                fit["cusolver_potrf_potri_ms"] = None
                fit["cusolver_error"] = str(e)[:200]
            # the same iteration inside GP.train (gp.py:54-63): value + gradient + Adam update per epoch, the epoch
            # replayed as a CUDA graph (what a suggest() pays 100 times); second train() call timed, 20 epochs
            cov_t = B.ScaleKernel(B.RBFKernel(ard_dim=a.d)).double()
            with torch.no_grad():
                cov_t.base_kernel.lengthscale.fill_(raw_ls)
                cov_t.variance.fill_(raw_var)
            gpt = B.CudaGP(torch.as_tensor(X), torch.as_tensor(y).reshape(-1, 1), B.ConstantMean().double(), cov_t,
                           device=dev, noise_variance=NOISE, epochs=20)
            gpt.train()
            torch.cuda.synchronize()
            f0.record()
            gpt.train()
            f1.record()
            torch.cuda.synchronize()
            fit["train_ms_per_epoch"] = f0.elapsed_time(f1) / 20
            del gpt

        # ---- CPU baseline on this host's cores (bounded sample; N=1 only): the unmodified reference, port beside it
        if world == 1 and not a.no_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            Xs = cpu_pool_sample(a, a.cpu_sample)
            port = PortArm(a)
            prate, pdt, pchunk = cpu_rate(port, Xs, 2, 1, chunks=(256, 1024))
            cpu_port = {"value": prate, "unit": "candidates/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{a.cpu_sample} candidates x 2 steps, chunk {pchunk}, oracle port (torch CPU float64)"}
            if reference_available():
                ref = ReferenceArm(a)
                rrate, rdt, rchunk = cpu_rate(ref, torch.as_tensor(Xs), 2, 1)
                cpu = {"value": rrate, "unit": "candidates/s", "cores": torch.get_num_threads(), "kind": "reference",
                       "sample": f"{a.cpu_sample} candidates x 2 steps, chunk {rchunk}, unmodified reference "
                                 "(baseline/_ref): GP.predict -> diagonal -> EI -> topk, torch CPU float64"}
                if fit is not None and a.n <= 4096:
                    ref.fit_iter()
                    t0 = time.perf_counter()
                    ref.fit_iter()
                    fit["cpu_s_per_iter"] = time.perf_counter() - t0
                    fit["cpu_kind"] = "reference (neg_log_marginal_likelihood + backward)"
            else:
                cpu = cpu_port
            if fit is not None and "cpu_s_per_iter" not in fit and a.n <= 4096:
                t0 = time.perf_counter()
                port.fit_iter()
                fit["cpu_s_per_iter"] = time.perf_counter() - t0
                fit["cpu_kind"] = "port"
            if fit is not None:
                fit["cpu_cores"] = torch.get_num_threads()

    if rank == 0:
        line = {
            "metric": "EI candidates scored/sec", "value": value, "unit": "candidates/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": a.scaling, "vs_baseline": None, "dtype": "f64" if a.precision == "fp64" else
            ("f16 operands (TF32-width mantissa) x f32 accumulate" + (" + f64 re-score of the short-list"
                                                                     if a.precision == "tf32+refine" else "")),
            "data": "synthetic", "config": config_dict(a, world), "precision": a.precision,
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "cpu_port": cpu_port, "fit": fit, "other_mode": other, "selection_matches_fp64": sel_match,
            "refine": cert, "collective_ms": collective_ms, "collective_in_step_ms": collective_in_step_ms,
            "state_replication_ms": state_ms,
            "selection_matches_single_gpu": single,
            "top": {"scores": [float(v) for v in top_s.cpu()], "indices": [int(v) for v in top_i.cpu()]},
        }
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
    # rank 0), so everything but the result line is sent to stderr: fd 1 is pointed at fd 2 for the run and the
    # line goes to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global emit_line

    def emit_line(line):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
