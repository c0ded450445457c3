#!/usr/bin/env python
"""
bench.py -- collocation points / second through the fused fwd+bwd loss (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload heat|burgers|kdv|ks|rd2d]
                    [--activation tanh|rational] [--impl ours|reference] [--scaling weak|strong] [--total-points T]

A "step" is one closure-equivalent of the reference training loop (Code/Test_Train.py:131-190): L^p loss +
per network {collocation loss, data loss, L2 penalty} -> weighted total -> gradients of every network
parameter and of Xi (+ one flat NCCL all-reduce of the gradients when N > 1).  Point sampling and the
optimiser update are outside the step.

Default workload = BASELINE.json configs[1]: Heat 1-D, 2^20 collocation points, MLP 5x40 (tanh), one B200.
Weak scaling (default): every rank owns the full per-GPU row count; the mean runs over the global count.
Strong scaling (--scaling strong): the configuration's TOTAL row count (BASELINE configs 3-5: KdV 2^22, KS 2^24, rd2d 2^26) is fixed
and split over the ranks, rows_per_gpu = total / N.
--impl reference times the UNMODIFIED reference (baseline/_ref, see baseline/install_ref.py) on the host cores.

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
    ap.add_argument("--points", type=int, default=0, help="rows per GPU and network (default: the config's size)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--total-points", type=int, default=0, help="strong scaling: total rows over all GPUs and networks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rational", action="store_true", help="skip the Rational-activation sub-record of the default line")
    return ap.parse_args()


def kernel_stamp(plan_source):
    """Identifies the kernel a DRAM-traffic capture belongs to: hash of the generated plan source + the hand-written kernel template
    + the launch helpers (scripts/record_traffic.py writes it next to the ncu numbers, bench.py reports them only while it matches)."""
    import hashlib
    csrc = os.path.join(ROOT, "pde_learn_b200", "csrc")
    return hashlib.sha1((plan_source + open(os.path.join(csrc, "jet_mlp_kernel_mma.cuh")).read() +
                         open(os.path.join(csrc, "jet_common.cuh")).read() +
                         open(os.path.join(csrc, "kernel_params.h")).read()).encode()).hexdigest()[:16]


def per_gpu_points(workload, n_gpus, scaling="weak", total_points=0):
    """Rows per GPU and per network."""
    from pde_learn_b200 import configs
    total = total_points if total_points > 0 else configs.CONFIG_POINTS[workload]
    S = configs.CONFIG_NETWORKS[workload]
    if scaling == "strong":
        return max(8, total // S // n_gpus)
    # weak: configs quoted on several GPUs (kdv: 2/4, ks: 8) keep their per-GPU share; others are per-GPU sizes
    share = {"kdv": 2, "ks": 8}.get(workload, 1)
    return total // share // S


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference (baseline/_ref) through its own Training(); the oracle port only
# when no copy of the reference is available (then kind = "port")
# ------------------------------------------------------------------------------------------------------

def _oracle_port_rate(workload, activation, steps, warmup, chunk=CPU_CHUNK):
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
    return {"value": S * chunk * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port", "steps": len(times), "warmup": warmup,
            "sample": "%d closure(s) of %d rows x %d network(s) (+%d data rows), oracle port of the reference algorithm, float32, torch %s"
                      % (steps, chunk, S, N_DATA, torch.__version__),
            "ms_per_step": 1e3 * statistics.mean(times), "rows": chunk, "networks": S}


def cpu_reference_rate(workload, activation, steps, warmup, chunk=CPU_CHUNK, in_process=False):
    """Times the reference's own CPU implementation on a bounded sample (chunk rows per step).  Runs baseline/reference_arm.py in a
    subprocess unless `in_process` (the --impl reference leg imports nothing of ours, so it can run it directly)."""
    arm = os.path.join(ROOT, "baseline", "reference_arm.py")
    res = None
    if in_process:
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import reference_arm
        res = reference_arm.run(workload, activation, chunk, steps, warmup)
    else:
        try:
            r = subprocess.run([sys.executable, arm, "--workload", workload, "--activation", activation, "--rows", str(chunk),
                                "--steps", str(steps), "--warmup", str(warmup), "--max-seconds", "60"],
                               capture_output=True, text=True, timeout=600, cwd=ROOT)
            res = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:                                  # noqa: BLE001
            res = {"unavailable": "reference arm failed: %r" % (e,)}
    if "unavailable" in res:
        port = _oracle_port_rate(workload, activation, steps, warmup, chunk)
        port["note"] = res["unavailable"]
        return port
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_rate(args.workload, args.activation, max(1, args.steps), max(1, args.warmup), in_process=True)
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": cb["steps"], "warmup": cb["warmup"],
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args.workload, args.activation), "rows_per_step": cb["rows"], "networks": cb["networks"],
                       "note": "the reference's own CPU implementation (Code/Test_Train.py:22 Training: closure + Adam step) on the host "
                               "cores, bounded sample of the workload (the reference needs ~160 KB per row); points/s is size-normalised"},
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
    B = args.points if args.points > 0 else per_gpu_points(w, world, args.scaling, args.total_points)

    def make_networks(activation):
        torch.manual_seed(0)
        nets = [P.Network([d] + [40] * 5 + [1], "Tanh" if activation == "tanh" else "Rational", Device=dev) for _ in range(S)]
        if world > 1:                                   # replicas must start identical
            for U in nets:
                for q in U.parameters():
                    dist.broadcast(q.data, 0)
        return nets

    U_List = make_networks(args.activation)
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

    def make_closure(nets):
        prms = [q for U in nets for q in U.parameters()] + [Xi]

        def closure(points_list):
            for q in prms:
                q.grad = None
            # one fused autograd node: Lp + per network {Coll, Data, L2} + weighted gradient combination; its backward all-reduces
            # the flat gradient buffer once when N > 1 (engine._ClosureFn)
            total, _scalars, _res = P.Closure_Loss(nets, Xi, Mask, points_list, x_in, y_in, derivs, lhs, rhs, P_EXP, WEIGHTS,
                                                   inv_n_coll=[1.0 / n_glob] * S, inv_n_data=[1.0 / (nd_loc * world)] * S,
                                                   world=world, group=group)
            total.backward()
            return total
        return closure

    closure = make_closure(U_List)

    flush = torch.empty(192 * 1024 * 1024, dtype=torch.float32, device=dev)      # 768 MB > 126 MB L2

    def barrier():
        if group is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(fn, steps, warmup, sampler=None):
        """W untimed + K device-timed steps (CUDA events on the launching stream), barrier + synchronize on both sides, max over ranks."""
        for _ in range(warmup):
            fn(coll)
        barrier()
        if sampler is not None:
            sampler.rows.clear()                       # keep only samples taken during the timed region
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            flush.zero_()                               # evict L2 between timed iterations (inputs may be smaller than L2)
            ev[i][0].record()
            fn(coll)
            ev[i][1].record()
        barrier()
        t_total = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
        if group is not None:
            dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
        return float(t_total) / steps

    # ---- device-timed steps, inputs resident in HBM ----
    sampler = ClockSampler(local); sampler.start()
    ms_per_step = timed_steps(closure, args.steps, args.warmup, sampler)
    clocks = sampler.stop()
    value = S * n_glob / (ms_per_step * 1e-3)

    # ---- dominant kernel alone (roofline numerator/denominator) ----
    mask_dev = engine.device_mask(Mask, K, dev)

    def kernel_alone(nets, steps, warmup, bwd=True):
        """Average duration of the collocation fwd+adjoint kernel (+ its tiny reduce) launched on its own: CUDA events on the
        launching stream, L2 flushed between launches.  bwd=False: the launch without gradients (Testing / evaluation)."""
        plan_ = engine.collocation_plan(nets[0], derivs, lhs, rhs)
        prm0 = [q.detach() for q in engine.network_param_tensors(nets[0])]
        resid = torch.empty(B, dtype=torch.float32, device=dev)
        stats = torch.empty(4, dtype=torch.float64, device=dev)
        gp = torch.empty(plan_.n_param_scalars, dtype=torch.float32, device=dev)
        gx = torch.empty(K, dtype=torch.float32, device=dev)
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(warmup + steps):
            flush.zero_()
            if i >= warmup:
                kev[i - warmup][0].record()
            plan_.launch(coll[0], prm0, Xi.detach(), mask_dev, None, 1.0 / n_glob, bwd, resid, None, gp if bwd else None,
                         gx if bwd else None, stats)
            if i >= warmup:
                kev[i - warmup][1].record()
        torch.cuda.synchronize()
        return plan_, statistics.mean(a.elapsed_time(b) for a, b in kev)

    plan, k_ms = kernel_alone(U_List, args.steps, args.warmup)
    achieved_tf = plan.flops_per_point * B / (k_ms * 1e-3) / 1e12
    peak = C.c_double(0.0)
    L.check(L.lib().pdl_measure_fp32_peak(C.byref(peak), torch.cuda.current_stream().cuda_stream))
    hbm_bytes = B * (4 * d + 4)                                          # algorithmic: read points, write residual
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    src = L.plan_source(plan.desc)
    on_tensor_pipe = "jet_mlp_kernel_mma.cuh" in src
    # The contractions run on the tensor pipe, so the denominator is the tensor-pipe peak for their input type: dense TF32 =
    # half of the measured cuBLAS bf16 rate; the kernel is timed alone, hence the burst figure.  The numerator stays the algorithmic F.
    if peaks.get("bf16_tflops"):
        tf32_peak, tf32_src = float(peaks["bf16_tflops"]) / 2.0, "MEASURED_PEAKS.json bf16_tflops (burst, cuBLAS) / 2 = dense TF32, of measured"
    else:
        tf32_peak, tf32_src = 1590.0 / 2.0, "B200_PROFILING.md fallback 1.59 PFLOP/s bf16 / 2 = dense TF32, of fallback"
    # DRAM traffic of this kernel is an ncu number (profiles/traffic.json); it is reported only while the kernel it was captured for
    # is still the one that runs (stamp = hash of the generated plan source + the kernel template).
    stamp = kernel_stamp(src)
    traffic, traffic_note = None, "no ncu capture recorded for this workload"
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("%s_%s" % (w, args.activation))
        if rec and rec.get("points_per_launch") == B:
            if rec.get("stamp") == stamp:
                traffic, traffic_note = float(rec["bytes"]), rec.get("source", "")
            else:
                traffic_note = "stale: captured for kernel stamp %s, running %s" % (rec.get("stamp"), stamp)
    except Exception:
        pass
    fp32_peak = float(peak.value)
    roofline = {"bound": "tensor" if on_tensor_pipe else "fp32",
                "kernel": "pdl_main_kernel<true> (fused jet-MLP fwd+adjoint, %s %s)" % (w, args.activation),
                "achieved": achieved_tf, "unit": "TFLOP/s",
                "peak": tf32_peak if on_tensor_pipe else fp32_peak,
                "frac": achieved_tf / (tf32_peak if on_tensor_pipe else fp32_peak),
                "peak_source": tf32_src if on_tensor_pipe else "FFMA-chain microbenchmark on this GPU in this run",
                "frac_fp32": achieved_tf / fp32_peak if fp32_peak > 0 else None, "peak_fp32": fp32_peak,
                "peak_fp32_source": "FFMA-chain microbenchmark pdl_measure_fp32_peak on this GPU in this run (nominal 148*128*2*1.965GHz = 74.4); "
                                    "the FP32-accurate view: what the same flops would need on the CUDA cores",
                "flops_per_point": plan.flops_per_point, "points_per_launch": B, "kernel_ms": k_ms, "kernel_stamp": stamp,
                "hbm_gbs_algorithmic": hbm_bytes / (k_ms * 1e-3) / 1e9, "algorithmic_bytes": hbm_bytes,
                "hbm_peak_gbs": peaks.get("hbm_gbs"), "traffic": traffic, "traffic_note": traffic_note}
    # the tensor-pipe view of the same launch: executed MMA instructions (split products, padding included) against the measured
    # mma.sync rate.  A TF32 m16n8k8 and a bf16 m16n8k16 occupy the pipe equally long (277 vs 554 TFLOP/s measured), so the unit
    # is the TF32-equivalent instruction: 2048 flop.
    try:
        import re
        df = {k: int(v) for k, v in re.findall(r"#define (PDL_WP|PDL_J|PDL_H|PDL_DGRAD_MMA|PDL_WGRAD_MMA|PDL_CROSS_BF16) (\d+)", src)}
        if "jet_mlp_kernel_mma.cuh" in src:
            nt8, nm, nmt = df["PDL_WP"] // 8, (df["PDL_J"] + 1) // 2, (df["PDL_WP"] + 15) // 16
            xb = df.get("PDL_CROSS_BF16", 0)
            per_prod = {"forward": 2 if xb & 9 else 3, "dgrad": 2 if xb & 2 else 3, "wgrad": 2 if xb & 4 else 3}
            per_layer = (per_prod["forward"] * nt8 * nt8 * nm + (per_prod["dgrad"] * nt8 * nt8 * nm if df["PDL_DGRAD_MMA"] else 0) +
                         (per_prod["wgrad"] * nmt * nt8 * df["PDL_J"] if df["PDL_WGRAD_MMA"] else 0))
            mma_flops_pt = per_layer * (df["PDL_H"] - 1) * (16 * 8 * 8 * 2) / 8.0
            exe = mma_flops_pt * B / (k_ms * 1e-3) / 1e12
            roofline["tensor_pipe"] = {"executed_tflops": exe, "peak": 277.0, "frac": exe / 277.0, "mma_per_tile_layer": per_layer,
                                       "mma_per_product": per_prod,
                                       "algorithmic_frac_of_sustained_tf32_peak": (achieved_tf / (float(peaks["bf16_tflops_sustained"]) / 2.0)
                                                                                   if peaks.get("bf16_tflops_sustained") else None),
                                       "contractions_on_tensor_pipe": ["forward"] + (["dgrad"] if df["PDL_DGRAD_MMA"] else []) +
                                                                      (["wgrad"] if df["PDL_WGRAD_MMA"] else []),
                                       "peak_source": "mma.sync.m16n8k8 TF32 microbenchmark, profiles/r01_ubench.txt (277 TFLOP/s = "
                                                      "1.35e11 MMA instructions/s; bf16 m16n8k16 issues at the same rate); split "
                                                      "products (TF32 hi*hi + one 16-bit cross-term MMA: scaled fp16 forward, bf16 adjoint) and padding are "
                                                      "executed, not algorithmic, work; executed_tflops counts 2048 flop per MMA"}
    except Exception as ex:                                                  # informational only
        roofline["tensor_pipe"] = {"error": str(ex)}

    # ---- end to end through the public API (P.Closure_Loss + backward, the closure of Training(); not the Adam step): every step
    #      copies its collocation rows from pinned host memory to the device and reads its loss back to the host.  The loss travels
    #      with a non-blocking copy into pinned memory and the host consumes it one step later (a logging loop), so launch latency
    #      is not exposed; every loss has been read when the timed region ends ----
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

    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_run(n_steps):
        seen = []
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
            loss_host[slot:slot + 1].copy_(total.detach().reshape(1), non_blocking=True)   # device -> host read of the step's result
            loss_ready[slot].record()
            if it > 0:                                       # consume the previous step's loss on the host
                loss_ready[slot ^ 1].synchronize()
                seen.append(float(loss_host[slot ^ 1]))
        loss_ready[(n_steps - 1) & 1].synchronize()
        seen.append(float(loss_host[(n_steps - 1) & 1]))
        assert len(seen) == n_steps and all(v == v for v in seen)

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

    # ---- the reference's default activation (Settings.txt:26 "Rational") on the same workload: device-timed closure + kernel alone ----
    rational = None
    if args.activation == "tanh" and not args.no_rational:
        R_List = make_networks("rational")
        r_steps, r_warm = max(3, args.steps // 2), max(3, args.warmup // 2)
        r_ms = timed_steps(make_closure(R_List), r_steps, r_warm)
        r_plan, r_kms = kernel_alone(R_List, r_steps, r_warm)
        r_tf = r_plan.flops_per_point * B / (r_kms * 1e-3) / 1e12
        rational = {"value": S * n_glob / (r_ms * 1e-3), "unit": UNIT, "ms_per_step": r_ms, "steps": r_steps, "warmup": r_warm,
                    "kernel_ms": r_kms, "achieved_tflops": r_tf, "frac": r_tf / (tf32_peak if on_tensor_pipe else fp32_peak),
                    "frac_fp32": r_tf / fp32_peak if fp32_peak > 0 else None,
                    "workload": workload_name(w, "rational")}

    # ---- launches without gradients (Testing, Code/Test_Train.py:205-330; residual evaluation): the tcgen05 / TMEM engine when the
    # plan has one.  Not part of BASELINE.json's metric (that is the fused fwd+bwd loss); reported because it is the Blackwell-native path.
    f_steps, f_warm = max(3, args.steps // 2), max(3, args.warmup // 2)
    _fp, f_ms = kernel_alone(U_List, f_steps, f_warm, bwd=False)
    f_tf = plan.flops_per_point / 3.0 * B / (f_ms * 1e-3) / 1e12          # forward contraction = one third of F
    forward_only = {"value": world * B / (f_ms * 1e-3), "unit": "points/s (one network, launch without gradients, kernel alone)",
                    "kernel_ms": f_ms, "steps": f_steps, "warmup": f_warm,
                    "engine": "tcgen05/TMEM (pdl_fwd_tc_kernel)" if "#define PDL_FWD_TC 1" in src else "mma.sync (pdl_main_kernel<false>)",
                    "achieved_tflops": f_tf, "frac": f_tf / (tf32_peak if on_tensor_pipe else fp32_peak)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(w, args.activation), "rows_per_gpu": B, "rows_total": S * n_glob, "networks": S,
                       "data_rows": N_DATA, "weights": WEIGHTS, "p": P_EXP,
                       "l2_policy": "768 MB buffer zeroed between timed iterations (inputs may be smaller than L2)",
                       "parallelism": "dp%d over collocation rows, one flat gradient all-reduce inside the closure node" % world},
            "clocks": clocks, "roofline": roofline,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "api": "pde_learn_b200.Closure_Loss(...).backward() = the closure of Training() (Code/Test_Train.py:143-188); "
                           "the reference's Training also runs the optimiser and returns the residual on the host, which is not part "
                           "of BASELINE.json's metric"},
            "gpu_launches": (2 + 6 * S) * args.steps}     # per step: Lp, per network {collocation + reduce, data + reduce, L2, axpby3}, closure tail
    line["forward_only"] = forward_only
    if rational is not None:
        line["rational"] = rational
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_rate(w, args.activation, steps=5, warmup=1)
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
