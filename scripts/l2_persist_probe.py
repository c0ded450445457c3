"""Experiment: does an L2 persisting access-policy window over the kernel's workspace (per-warp adjoint scratch + weight-gradient
partial sums, ~62 MB, rewritten by every tile) stop the L2 from writing the dirty lines back to DRAM (2.4 GB per launch)?
Run under `ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum -k regex:pdl_main_kernel`.
usage: python scripts/l2_persist_probe.py [0|1] [workload]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pde_learn_b200 as P
from pde_learn_b200 import configs, engine


class Window(C.Structure):
    _fields_ = [("base_ptr", C.c_void_p), ("num_bytes", C.c_size_t), ("hitRatio", C.c_float), ("hitProp", C.c_int), ("missProp", C.c_int)]


class AttrValue(C.Union):
    _fields_ = [("accessPolicyWindow", Window), ("pad", C.c_char * 64)]


on = len(sys.argv) > 1 and sys.argv[1] == "1"
w = sys.argv[2] if len(sys.argv) > 2 else "heat"
dev = torch.device("cuda", 0)
d, bounds, derivs, lhs, rhs = configs.config(w)
torch.manual_seed(0)
U = P.Network([d] + [40] * 5 + [1], "Tanh", Device=dev)
K = len(rhs); B = 1 << 20
xi = (0.1 * torch.randn(K)).to(dev); mask = torch.zeros(K, dtype=torch.uint8, device=dev)
pts = P.Generate_Points(bounds, B, dev, Seed=1)
plan = engine.collocation_plan(U, derivs, lhs, rhs)
ws, n_sm = plan.workspace(dev)
rt = C.CDLL("libcudart.so.12")
if on:
    prop_max = C.c_int(0); win_max = C.c_int(0)
    rt.cudaDeviceGetAttribute(C.byref(prop_max), 108, 0)      # cudaDevAttrMaxPersistingL2CacheSize
    rt.cudaDeviceGetAttribute(C.byref(win_max), 109, 0)       # cudaDevAttrMaxAccessPolicyWindowSize
    print("max persisting L2 %d MB, max window %d MB, workspace %.1f MB" % (prop_max.value >> 20, win_max.value >> 20, ws.numel() / 2**20))
    print("cudaDeviceSetLimit ->", rt.cudaDeviceSetLimit(6, C.c_size_t(prop_max.value)))   # cudaLimitPersistingL2CacheSize
    v = AttrValue()
    v.accessPolicyWindow = Window(ws.data_ptr(), min(ws.numel(), win_max.value), 1.0, 2, 1)   # hit: persisting, miss: streaming
    st = torch.cuda.current_stream().cuda_stream
    print("cudaStreamSetAttribute ->", rt.cudaStreamSetAttribute(C.c_void_p(st), 1, C.byref(v)))
prm = [q.detach() for q in engine.network_param_tensors(U)]
resid = torch.empty(B, dtype=torch.float32, device=dev); stats = torch.empty(4, dtype=torch.float64, device=dev)
gp = torch.empty(plan.n_param_scalars, dtype=torch.float32, device=dev); gx = torch.empty(K, dtype=torch.float32, device=dev)
ts = []
for i in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    plan.launch(pts, prm, xi, mask, None, 1.0 / B, True, resid, None, gp, gx, stats)
    b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print("persist", on, w, "ms", ["%.3f" % t for t in ts], "loss", float(stats[0]))
