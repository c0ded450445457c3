#!/usr/bin/env python
"""
bench.py -- collocation points / second through the fused fwd+bwd loss (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload heat|burgers|kdv|ks|rd2d]
                    [--activation tanh|rational] [--impl ours|reference]

A "step" is one closure-equivalent of the reference training loop (Code/Test_Train.py:131-190): L^p loss +
per network {collocation loss, data loss, L2 penalty} -> weighted total -> gradients of every network
parameter and of Xi (+ one flat NCCL all-reduce of the gradients when N > 1).  Point sampling and the
optimiser update are outside the step.

Default workload = BASELINE.json configs[1]: Heat 1-D, 2^20 collocation points, MLP 5x40 (tanh), one B200.
Weak scaling: every rank owns the full per-GPU row count; the mean runs over the global count.

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every key.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "collocation points/sec (fused fwd+bwd loss, device-timed)"
UNIT = "points/s"
WEIGHTS = {"Data": 1.0, "Coll": 1.0, "Lp": 2e-4, "L2": 5e-6}
P_EXP = 0.1
N_DATA = 10_000
CPU_CHUNK = 20_000            # rows per reference-CPU step (memory ~160 KB/row on the KS shape)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="heat")
    ap.add_argument("--activation", default="tanh", choices=["tanh", "rational"])
    ap.add_argument("--points", type=int, default=0, help="rows per GPU (default: the config's size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def per_gpu_points(workload, n_gpus):
    from pde_learn_b200 import configs
    total = configs.CONFIG_POINTS[workload]
    # configs quoted on several GPUs (kdv: 2/4, ks: 8) keep their per-GPU share; others are per-GPU sizes
    share = {"kdv": 2, "ks": 8}.get(workload, 1)
    return total // share // configs.CONFIG_NETWORKS[workload]


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port (same algorithm as the reference: nested autograd) on CPU
# ------------------------------------------------------------------------------------------------------

def cpu_reference_rate(workload, activation, steps, warmup, chunk=CPU_CHUNK):
    import torch
    from oracle import pde_oracle as O                      # cpu_baseline leg: allowed user of oracle/
    from pde_learn_b200 import configs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    d, lhs, rhs, bounds = O.config_library(workload)
    S = configs.CONFIG_NETWORKS[workload]
    K = len(rhs)
    gen = torch.Generator().manual_seed(0)
    nets = [O.init_network([d] + [40] * 5 + [1], activation, seed=i).to(torch.float32, True) for i in range(S)]
    xi = (0.1 * torch.randn(K, generator=gen)).requires_grad_(True)
    mask = [False] * K
    lo = torch.tensor([b[0] for b in bounds], dtype=torch.float32); hi = torch.tensor([b[1] for b in bounds], dtype=torch.float32)
    pts = [(lo + (hi - lo) * torch.rand(chunk, d, generator=gen)) for _ in range(S)]
    x_in = [(lo + (hi - lo) * torch.rand(N_DATA, d, generator=gen)) for _ in range(S)]
    y = [torch.randn(N_DATA, generator=gen, dtype=torch.float64) for _ in range(S)]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        O.closure_losses(nets, xi, mask, [p.clone() for p in pts], x_in, y, lhs, rhs, P_EXP, WEIGHTS, backward=True)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    best = min(times)
    return {"value": S * chunk / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d closure(s) of %d rows x %d network(s) (+%d data rows), float32, torch %s, best of %d"
                      % (steps, chunk, S, N_DATA, torch.__version__, steps),
            "ms_per_step": 1e3 * statistics.mean(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5)); warmup = max(1, min(args.warmup, 2))
    cb = cpu_reference_rate(args.workload, args.activation, steps, warmup)
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args.workload, args.activation), "rows_per_step": CPU_CHUNK,
                       "note": "reference algorithm (nested torch.autograd) on the host cores, bounded sample"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(w, act):
    names = {"burgers": "Burgers 1D shipped Library.txt (J=5,K=17)", "heat": "Heat 1D order-2 library (J=4,K=9)",
             "kdv": "KdV 1D order-3 library (J=5,K=14)", "ks": "Kuramoto-Sivashinsky 1D order-4 library (J=6,K=30)",
             "rd2d": "2D reaction-diffusion, 2 networks sharing Xi (J=6,K=21)"}
    return "%s, MLP 5x40 %s" % (names[w], act)


# ------------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------------

class ClockSampler:
    """Samples SM clock, power and throttle reasons of one GPU through NVML every ~5 ms while a region runs."""

    def __init__(self, index):
        self.index, self.rows, self._stop, self._thr, self.err = index, [], threading.Event(), None, None

    def start(self):
        try:
            import pynvml as N
            N.nvmlInit()
            self.N = N
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            except Exception:
                pass
            self.h = None
            if uuid:
                try:
                    self.h = N.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
                except Exception:
                    self.h = None
            if self.h is None:
                self.h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM)
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        except Exception as e:                      # noqa: BLE001
            self.err = repr(e)

    def _run(self):
        N = self.N
        while not self._stop.is_set():
            try:
                sm = N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM)
                pw = N.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = N.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.rows.append((sm, pw, rs))
            except Exception as e:                  # noqa: BLE001
                self.err = repr(e)
                return
            time.sleep(0.005)

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples: %s" % self.err], "samples": 0}
        N = self.N
        names = {"hw_slowdown": N.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": N.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": N.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": N.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(k for k, bit in names.items() if any(r[2] & bit for r in self.rows))
        pmax = max(r[1] for r in self.rows)
        load = [r[0] for r in self.rows if r[1] >= 0.6 * pmax] or [r[0] for r in self.rows]
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": self.max_sm, "reasons": reasons,
                "power_w_max": pmax, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------

def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import pde_learn_b200 as P
    from pde_learn_b200 import configs, engine
    from pde_learn_b200 import _lib as L
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl ours) needs CUDA devices; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    w = args.workload
    d, bounds, derivs, lhs, rhs = configs.config(w)
    S = configs.CONFIG_NETWORKS[w]
    K = len(rhs)
    B = args.points if args.points > 0 else per_gpu_points(w, world)
    act = "Tanh" if args.activation == "tanh" else "Rational"
    torch.manual_seed(0)
    U_List = [P.Network([d] + [40] * 5 + [1], act, Device=dev) for _ in range(S)]
    if world > 1:                                       # replicas must start identical
        for U in U_List:
            for q in U.parameters():
                dist.broadcast(q.data, 0)
    gen = torch.Generator().manual_seed(0)
    Xi = (0.1 * torch.randn(K, generator=gen)).to(dev).requires_grad_(True)
    Mask = torch.zeros(K, dtype=torch.bool)
    n_glob = B * world
    # rank r owns global rows [r*B, (r+1)*B) of every network's batch; counter-based sampling
    coll = [P.Generate_Points(bounds, B, dev, Seed=100 + s, First_Index=rank * B) for s in range(S)]
    nd_loc = N_DATA // world
    x_all = [P.Generate_Points(bounds, N_DATA, dev, Seed=200 + s) for s in range(S)]
    y_all = [torch.randn(N_DATA, generator=torch.Generator().manual_seed(1 + s), dtype=torch.float64).to(dev) for s in range(S)]
    x_in = [x[rank * nd_loc:(rank + 1) * nd_loc].contiguous() for x in x_all]
    y_in = [y[rank * nd_loc:(rank + 1) * nd_loc].contiguous() for y in y_all]
    params = [q for U in U_List for q in U.parameters()] + [Xi]

    def closure(points_list):
        for q in params:
            q.grad = None
        # one fused autograd node: Lp + per network {Coll, Data, L2} + weighted gradient combination (engine._ClosureFn)
        total, _scalars, _res = P.Closure_Loss(U_List, Xi, Mask, points_list, x_in, y_in, derivs, lhs, rhs, P_EXP, WEIGHTS,
                                               inv_n_coll=[1.0 / n_glob] * S, inv_n_data=[1.0 / (nd_loc * world)] * S, world=world)
        total.backward()
        if group is not None:
            P.allreduce_gradients(params, group)
        return total

    flush = torch.empty(192 * 1024 * 1024, dtype=torch.float32, device=dev)      # 768 MB > 126 MB L2

    def barrier():
        if group is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed steps, inputs resident in HBM ----
    sampler = ClockSampler(local); sampler.start()
    for _ in range(args.warmup):
        closure(coll)
    barrier()
    sampler.rows.clear()                               # keep only samples taken during the timed region
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()                                   # evict L2 between timed iterations (inputs are 8 MB)
        ev[i][0].record()
        closure(coll)
        ev[i][1].record()
    barrier()
    clocks = sampler.stop()
    ms = [a.elapsed_time(b) for a, b in ev]
    t_total = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    if group is not None:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_total) / args.steps
    value = S * n_glob / (ms_per_step * 1e-3)

    # ---- dominant kernel alone (roofline numerator/denominator) ----
    plan = engine.collocation_plan(U_List[0], derivs, lhs, rhs)
    prm0 = engine.network_param_tensors(U_List[0])
    mask_dev = engine.device_mask(Mask, K, dev)
    resid = torch.empty(B, dtype=torch.float32, device=dev)
    stats = torch.empty(4, dtype=torch.float64, device=dev)
    gp = torch.empty(plan.n_param_scalars, dtype=torch.float32, device=dev)
    gx = torch.empty(K, dtype=torch.float32, device=dev)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.warmup + args.steps):
        flush.zero_()
        if i >= args.warmup:
            kev[i - args.warmup][0].record()
        plan.launch(coll[0], [q.detach() for q in prm0], Xi.detach(), mask_dev, None, 1.0 / n_glob, True, resid, None, gp, gx, stats)
        if i >= args.warmup:
            kev[i - args.warmup][1].record()
    torch.cuda.synchronize()
    k_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)
    achieved_tf = plan.flops_per_point * B / (k_ms * 1e-3) / 1e12
    peak = C.c_double(0.0)
    L.check(L.lib().pdl_measure_fp32_peak(C.byref(peak), torch.cuda.current_stream().cuda_stream))
    hbm_bytes = B * (4 * d + 4)                                          # algorithmic: read points, write residual
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(w)
    except Exception:
        pass
    roofline = {"bound": "fp32", "kernel": "pdl_main_kernel<true> (fused jet-MLP fwd+adjoint, %s)" % w,
                "achieved": achieved_tf, "peak": float(peak.value), "unit": "TFLOP/s",
                "frac": achieved_tf / float(peak.value) if peak.value > 0 else None,
                "peak_source": "FFMA-chain microbenchmark pdl_measure_fp32_peak on this GPU in this run "
                               "(MEASURED_PEAKS.json has no FP32 entry; nominal 148*128*2*1.965GHz = 74.4)",
                "flops_per_point": plan.flops_per_point, "points_per_launch": B, "kernel_ms": k_ms,
                "hbm_gbs_algorithmic": hbm_bytes / (k_ms * 1e-3) / 1e9,
                "hbm_peak_gbs": peaks.get("hbm_gbs"), "traffic": traffic}
    # the tensor-pipe view of the same launch: executed MMA instructions (split products, padding included) against the measured
    # mma.sync rate.  A TF32 m16n8k8 and a bf16 m16n8k16 occupy the pipe equally long (277 vs 554 TFLOP/s measured), so the unit
    # is the TF32-equivalent instruction: 2048 flop.
    try:
        import re
        src = L.plan_source(plan.desc)
        df = {k: int(v) for k, v in re.findall(r"#define (PDL_WP|PDL_J|PDL_H|PDL_DGRAD_MMA|PDL_WGRAD_MMA|PDL_CROSS_BF16) (\d+)", src)}
        if "jet_mlp_kernel_mma.cuh" in src:
            nt8, nm, nmt = df["PDL_WP"] // 8, (df["PDL_J"] + 1) // 2, (df["PDL_WP"] + 15) // 16
            xb = df.get("PDL_CROSS_BF16", 0)
            per_prod = {"forward": 2 if xb & 1 else 3, "dgrad": 2 if xb & 2 else 3, "wgrad": 2 if xb & 4 else 3}
            per_layer = (per_prod["forward"] * nt8 * nt8 * nm + (per_prod["dgrad"] * nt8 * nt8 * nm if df["PDL_DGRAD_MMA"] else 0) +
                         (per_prod["wgrad"] * nmt * nt8 * df["PDL_J"] if df["PDL_WGRAD_MMA"] else 0))
            mma_flops_pt = per_layer * (df["PDL_H"] - 1) * (16 * 8 * 8 * 2) / 8.0
            exe = mma_flops_pt * B / (k_ms * 1e-3) / 1e12
            roofline["tensor_pipe"] = {"executed_tflops": exe, "peak": 277.0, "frac": exe / 277.0, "mma_per_tile_layer": per_layer,
                                       "mma_per_product": per_prod,
                                       # transparency: the same algorithmic flops against the tcgen05-class dense TF32 peak (half of the
                                       # measured bf16 rate in MEASURED_PEAKS.json) -- a pipe this kernel does not use at width 40
                                       "algorithmic_frac_of_tcgen05_tf32_peak": (achieved_tf / (float(peaks["bf16_tflops_sustained"]) / 2.0)
                                                                                 if peaks.get("bf16_tflops_sustained") else None),
                                       "contractions_on_tensor_pipe": ["forward"] + (["dgrad"] if df["PDL_DGRAD_MMA"] else []) +
                                                                      (["wgrad"] if df["PDL_WGRAD_MMA"] else []),
                                       "peak_source": "mma.sync.m16n8k8 TF32 microbenchmark, profiles/r01_ubench.txt (277 TFLOP/s = "
                                                      "1.35e11 MMA instructions/s; bf16 m16n8k16 issues at the same rate); split "
                                                      "products (3 MMAs, or TF32 hi*hi + one bf16 cross-term MMA) and padding are "
                                                      "executed, not algorithmic, work; executed_tflops counts 2048 flop per MMA"}
    except Exception as ex:                                                  # informational only
        roofline["tensor_pipe"] = {"error": str(ex)}

    # ---- end to end through the public API: pinned host points -> device, closure, loss back to host ----
    host_pts = [c.cpu().pin_memory() for c in coll]
    dev_pts = [[torch.empty_like(c) for c in coll] for _ in range(2)]              # double buffer
    copy_stream = torch.cuda.Stream(dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(slot):
        """H2D copy of one step's collocation rows on the copy stream (overlaps the previous step's kernels)."""
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[slot])
            for i in range(S):
                dev_pts[slot][i].copy_(host_pts[i], non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_run(n_steps):
        for ev_ in freed:
            ev_.record()
        upload(0)
        for it in range(n_steps):
            slot = it & 1
            if it + 1 < n_steps:
                upload(slot ^ 1)                            # next step's inputs travel while this step computes
            torch.cuda.current_stream(dev).wait_event(ready[slot])
            total = closure(dev_pts[slot])
            freed[slot].record()
            _ = float(total.detach())                        # device -> host read of the step's result

    e2e_run(max(3, args.warmup // 2))
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if group is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = S * n_glob * args.steps / float(e2e_t)
    h2d = sum(int(c.numel()) * 4 for c in coll)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(w, args.activation), "rows_per_gpu": B, "networks": S,
                       "data_rows": N_DATA, "weights": WEIGHTS, "p": P_EXP,
                       "l2_policy": "768 MB buffer zeroed between timed iterations (inputs are smaller than L2)",
                       "parallelism": "dp%d over collocation rows, one flat gradient all-reduce" % world},
            "clocks": clocks, "roofline": roofline,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": (1 + 6 * S) * args.steps}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_rate(w, args.activation, steps=3, warmup=1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if group is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
