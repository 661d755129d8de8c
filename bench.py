#!/usr/bin/env python
"""bench.py -- headline benchmark of the GP-BO hot path (BASELINE.json metric).

metric   EI candidates scored/sec at n=4096, d=20 (whole job, all GPUs)
step     one pass of the hot path over one candidate pool per GPU:
         posterior mean/variance -> EI -> top-k (+ one all-gather of per-GPU top-k when N > 1)
value    pool already resident in HBM when the timed region starts
e2e      same metric through the C-ABI host entry point (bbo_gp_score_pool_host): the pool lives
         in pinned HOST memory; H2D of the pool and D2H of the selection are inside the timed region
also     "fit": GP fit ms/iter (value + analytic gradient, FP64) next to the CPU oracle's s/iter;
         "roofline" for the dominant kernel; "cpu_baseline" (oracle port on the host cores)

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp64|tf32] [--impl reference]
For N > 1 launch with torchrun (one rank per GPU); rank 0 prints ONE JSON line.
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
    ap.add_argument("--precision", default="tf32", choices=["fp64", "tf32"],
                    help="arithmetic of the n^2 panel in the headline number (the other mode is reported too)")
    ap.add_argument("--no-other-mode", action="store_true")
    ap.add_argument("--n", type=int, default=N_OBS)
    ap.add_argument("--d", type=int, default=DIM)
    ap.add_argument("--pool", type=int, default=1 << 20, help="candidates per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=4096, help="candidates per CPU-baseline step")
    ap.add_argument("--no-fit", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return (f"GP-BO EI pool scoring: Ackley-{a.d}D, n={a.n} observations, ARD RBF (l=0.5, s2=1, "
            f"noise=1e-4), {a.pool} uniform candidates per GPU per step, top-{TOPK}")


def problem(a):
    from bbo_b200.benchmarks import Ackley, make_observations
    X, y = make_observations(Ackley(a.d), a.n, SEED)
    raw_ls = float(np.log(np.expm1(LENGTHSCALE)))      # softplus^-1(0.5)
    raw_var = float(np.log(np.expm1(1.0)))
    return X, y, raw_ls, raw_var


# ------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port (the reference is pure Python and cannot travel
# to the GPU box; oracle/gp_oracle.py restates its arithmetic and is pinned to it by goldens)
# ------------------------------------------------------------------------------------------
def cpu_pool_sample(a, count):
    rng = np.random.default_rng(SEED + 1)
    return rng.random((count, a.d))


def cpu_state(a):
    from oracle import gp_oracle as O
    X, y, raw_ls, raw_var = problem(a)
    p = O.GPParams(kind=O.RBF, raw_lengthscale=np.full(a.d, raw_ls), raw_variance=raw_var, constant=0.0,
                   noise=NOISE)
    state = O.solve_gp_linear_system(p, X, y)
    return O, p, X, y, state


def cpu_score_rate(a, steps, warmup):
    """cand/s of the CPU oracle on `cpu_sample` candidates per step, best chunk of {256, 1024}."""
    O, p, X, y, state = cpu_state(a)
    Xq = cpu_pool_sample(a, a.cpu_sample)
    best_f = float(y.max())
    best = None
    for chunk in (256, 1024):
        for _ in range(max(1, min(warmup, 1))):
            O.score_pool(p, X, y, Xq[:chunk], "ei", best_f, k=TOPK, state=state, chunk=chunk)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.score_pool(p, X, y, Xq, "ei", best_f, k=TOPK, state=state, chunk=chunk)
        dt = (time.perf_counter() - t0) / steps
        if best is None or dt < best[0]:
            best = (dt, chunk)
    return a.cpu_sample / best[0], best[0], best[1]


def emit_line(line):   # rebound by main() to the saved stdout descriptor
    print(json.dumps(line), flush=True)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    steps = max(1, min(a.steps, 5))
    rate, dt, chunk = cpu_score_rate(a, steps, a.warmup)
    cores = torch.get_num_threads()
    sample = f"{a.cpu_sample} candidates/step, chunk {chunk}, oracle port (torch CPU float64)"
    line = {
        "impl": "reference", "metric": "EI candidates scored/sec", "value": rate, "unit": "candidates/s",
        "n_gpus": a.gpus, "steps": steps, "warmup": a.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "n": a.n, "d": a.d},
        "cpu_baseline": {"value": rate, "unit": "candidates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)


# ------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
            except ValueError:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return json.load(open(path))
    except Exception:  # noqa: BLE001
        return None


def live_matmul_peak(dtype, tf32, n=4096, reps=10):
    """cuBLAS GEMM throughput measured the way MEASURED_PEAKS.json measured bf16 (best of 10,
    CUDA events) -- the driver file has no FP64/TF32 entry, so the denominator is measured here."""
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


def live_matmul_sustained(dtype, tf32, n=4096, seconds=2.0):
    """The same GEMM back to back for `seconds` (MEASURED_PEAKS.json's "sustained" method, there 4 s for bf16): the
    denominator for a kernel that is timed inside a long step, where the chip sits at its power cap."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, dtype=dtype, device="cuda")
    b = torch.randn(n, n, dtype=dtype, device="cuda")
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    a @ b
    e1.record()
    torch.cuda.synchronize()
    reps = max(8, int(seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3)))
    e0.record()
    for _ in range(reps):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    torch.backends.cuda.matmul.allow_tf32 = old
    del a, b
    return 2.0 * n ** 3 * reps / ms / 1e9, reps, time.perf_counter() - t0


def run_b200(a):
    import torch.distributed as dist
    import bbo_b200 as B
    from bbo_b200 import _lib
    from bbo_b200.distributed import broadcast_state, sharded_score_pool
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

    X, y, raw_ls, raw_var = problem(a)
    cov = B.ScaleKernel(B.RBFKernel(ard_dim=a.d)).double()
    with torch.no_grad():
        cov.base_kernel.lengthscale.fill_(raw_ls)
        cov.variance.fill_(raw_var)
    gp = B.CudaGP(torch.as_tensor(X), torch.as_tensor(y).reshape(-1, 1), B.ConstantMean().double(), cov,
                  device=dev, noise_variance=NOISE, epochs=1)
    if world > 1:
        broadcast_state(gp, src=0)       # fit/factorise on rank 0 only, replicate L^-1 and alpha
    else:
        gp.state
    acq = B.EI(float(y.max()))
    q = a.pool
    lo = rank * q                                              # weak scaling: q candidates per GPU
    Xq = uniform_pool(SEED + 1, lo, q, a.d, device=dev)        # float64 (q,d): 168 MB at 2^20 x 20 > L2
    lib = _lib.load()

    def step():
        return sharded_score_pool(gp, acq, TOPK, Xq, lo, precision=a.precision)

    for _ in range(a.warmup):
        step()
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
    ms = e0.elapsed_time(e1)
    launches = lib.bbo_launch_count() - l0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clk = clocks.stop() if rank == 0 else None
    value = world * q * a.steps / (ms * 1e-3)

    # ---- e2e: pool in pinned host memory, through bbo_gp_score_pool_host
    e2e = None
    if not a.no_e2e:
        Xh = torch.empty(q, a.d, dtype=torch.float64).pin_memory()
        Xh.copy_(Xq)
        torch.cuda.synchronize()

        def step_host():
            hs, hi = gp.score_pool(Xh, acq, TOPK, precision=a.precision, index_offset=lo, host_pool=True)
            if world > 1:   # same single all-gather + merge on the gathered lists
                from bbo_b200.distributed import gather_topk, merge_topk_device
                gs, gi = gather_topk(hs.to(dev), hi.to(dev))
                ms_, mi_ = merge_topk_device(gs, gi, TOPK)
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
        ms_h = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_h], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_h = float(t.item())
        e2e = {"value": world * q * a.steps / (ms_h * 1e-3), "unit": "candidates/s",
               "h2d_bytes_per_step": q * a.d * 8, "d2h_bytes_per_step": TOPK * 16,
               "ms_per_step": ms_h / a.steps,
               "selection_matches_device_run": bool(torch.equal(hi.cpu(), top_i.cpu()))}
        del Xh

    # ---- the other precision mode, same workload, fewer steps (reported beside the headline)
    other = None
    if not a.no_other_mode:
        oprec = "fp64" if a.precision == "tf32" else "tf32"
        osteps = 2 if oprec == "fp64" else a.steps
        sharded_score_pool(gp, acq, TOPK, Xq, lo, precision=oprec)
        barrier()
        e0.record()
        for _ in range(osteps):
            os_, oi_ = sharded_score_pool(gp, acq, TOPK, Xq, lo, precision=oprec)
        e1.record()
        barrier()
        oms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([oms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            oms = float(t.item())
        other = {"precision": oprec, "value": world * q * osteps / (oms * 1e-3), "unit": "candidates/s",
                 "ms_per_step": oms / osteps, "steps": osteps,
                 "top1_index_matches_headline_mode": bool(int(oi_[0]) == int(top_i[0]))}

    # ---- roofline of the dominant kernel: separate profiled pass (CUDA events on the launch stream)
    roof = None
    fit = None
    cpu = None
    if rank == 0:
        tag = "vnorm" if a.precision == "fp64" else "tf32"
        lib.bbo_profile_enable(1)
        for _ in range(max(1, min(a.steps, 3))):
            gp.score_pool(Xq, acq, TOPK, precision=a.precision, index_offset=lo)
        regions, tot_ms = _lib.profile_read(tag)
        kregions, ktot = _lib.profile_read("kstar")
        sregions, stot = _lib.profile_read("select")
        lib.bbo_profile_enable(0)
        npasses = max(1, min(a.steps, 3))
        flops_per_launch = float(a.n) * a.n * (q * npasses / max(regions, 1))   # n^2 per candidate
        avg_ms = tot_ms / max(regions, 1)
        achieved = flops_per_launch / (avg_ms * 1e-3) / 1e12
        if a.precision == "fp64":
            peak = live_matmul_peak(torch.float64, False)
            peak_src = "measured live: cuBLAS f64 GEMM 4096^3 best of 10 (MEASURED_PEAKS.json has no f64 entry)"
        else:
            peak = live_matmul_peak(torch.float32, True, n=8192)
            peak_src = "measured live: cuBLAS tf32 GEMM 8192^3 best of 10 (MEASURED_PEAKS.json has no tf32 entry)"
        # `peak` stays the burst figure (best of 10: the conservative denominator).  The panel is timed inside a long
        # step (dozens of back-to-back launches at the power cap), where the matching figure is the SUSTAINED GEMM
        # rate: reported beside it as peak_sustained / frac_of_sustained.
        if a.precision == "fp64":
            peak_sus, sus_reps, _ = live_matmul_sustained(torch.float64, False, n=4096, seconds=1.5)
        else:
            peak_sus, sus_reps, _ = live_matmul_sustained(torch.float32, True, n=8192, seconds=1.5)
        peak_src += (f"; peak_sustained = the same GEMM back to back for {sus_reps} launches (~1.5 s), the method of "
                     "MEASURED_PEAKS.json's bf16_tflops_sustained")
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            traffic = tr.get(tag, {}).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass
        if other is not None:   # panel roofline of the other mode as well (same method)
            otag = "vnorm" if other["precision"] == "fp64" else "tf32"
            lib.bbo_profile_enable(1)
            gp.score_pool(Xq, acq, TOPK, precision=other["precision"], index_offset=lo)
            oreg, otot = _lib.profile_read(otag)
            lib.bbo_profile_enable(0)
            opeak = live_matmul_peak(torch.float64, False) if otag == "vnorm" else live_matmul_peak(torch.float32, True, n=8192)
            oach = float(a.n) * a.n * q / (otot * 1e-3) / 1e12
            other["roofline"] = {"bound": "tensor", "kernel": "k_vnorm (DMMA.8x8x4)" if otag == "vnorm" else "k_vnorm_tf32_2sm (tcgen05)",
                                 "achieved": oach, "peak": opeak, "unit": "TFLOP/s", "frac": oach / opeak,
                                 "launches_timed": oreg}
        share = tot_ms / max(tot_ms + ktot + stot, 1e-9)
        roof = {"bound": "tensor", "kernel": "k_vnorm (DMMA.8x8x4)" if tag == "vnorm" else "k_vnorm_tf32_2sm (tcgen05.mma cta_group::2 kind::tf32)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "peak_sustained": peak_sus, "frac_of_sustained": achieved / peak_sus,
                "peak_source": peak_src, "launches_timed": regions, "avg_launch_ms": avg_ms,
                "algorithmic_flops_per_launch": flops_per_launch, "share_of_step": share,
                "other_kernels_ms": {"kstar": ktot / max(kregions, 1), "select": stot / max(sregions, 1)}}

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
            fit = {"ms_per_iter": fit_ms, "n": a.n, "d": a.d, "dtype": "f64",
                   "achieved_tflops": fit_flops / (fit_ms * 1e-3) / 1e12,
                   "algorithmic_flops_per_iter": fit_flops}
            if a.precision != "fp64":
                peak64 = live_matmul_peak(torch.float64, False)
            else:
                peak64 = peak
            fit["peak_tflops"] = peak64
            fit["frac"] = fit["achieved_tflops"] / peak64
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

        # ---- CPU baseline on this host's cores (bounded sample; N=1 only)
        if world == 1 and not a.no_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            rate, dt, chunk = cpu_score_rate(a, 2, 1)
            cpu = {"value": rate, "unit": "candidates/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{a.cpu_sample} candidates x 2 steps, chunk {chunk}, oracle port (torch CPU float64)"}
            if fit is not None:
                O, p, Xc, yc, _ = cpu_state(a)
                t0 = time.perf_counter()
                O.objective_value_and_grad(p, Xc, yc)
                fit["cpu_s_per_iter"] = time.perf_counter() - t0
                fit["cpu_cores"] = torch.get_num_threads()

    if rank == 0:
        line = {
            "metric": "EI candidates scored/sec", "value": value, "unit": "candidates/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64" if a.precision == "fp64" else "tf32",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "n": a.n, "d": a.d, "pool_per_gpu": q, "k": TOPK,
                       "precision": a.precision, "parallelism": f"pool sharded x{world}, 1 all-gather of top-k",
                       "l2": "inputs larger than L2 (pool 168 MB + L^-1 134 MB per pass)"},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "fit": fit, "other_mode": other, "top1": {"score": float(top_s[0]), "index": int(top_i[0])},
        }
        emit_line(line)
    if world > 1:
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
