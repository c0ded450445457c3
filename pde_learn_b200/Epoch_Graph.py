"""
Captured_Epochs -- the whole training epoch of Code/main.py:281-405 as ONE CUDA graph.

Per replay, for every data set i: draw Num_Train_Coll_Points fresh collocation points (Philox key read from device
memory), append the targeted points kept from the previous epoch (count on the device), evaluate the closure of
Code/Test_Train.py:143-188 (L^p + collocation + data + L2 losses and all gradients: the same fused kernels Training()
launches), apply the Adam update of Code/Test_Train.py:193 / main.py:232-235, select the next targeted points
(|r| >= mean + 3 std over the random rows, Code/main.py:365-384) and append the epoch's losses to a device-side history.
Nothing returns to the host between epochs: one cudaGraphLaunch per epoch instead of ~60 kernel launches, ~40 tensor
allocations and 2 host synchronisations.  This is what small batches need (config 1, 10 000 points, is launch bound).

The optimiser stays a torch.optim.Adam: its state tensors (exp_avg, exp_avg_sq) are updated in place by the fused
kernel and state["step"] is kept in sync, so Optimizer.state_dict() / Save_State keep the reference's format
(Code/main.py:519).  Restrictions (checked, loud): Adam without weight decay / amsgrad / maximize, at most `Max_Targeted`
targeted rows per data set and rank are carried over (the reference has no bound; the count actually selected is recorded,
and exceeding the capacity raises when the history is read).

Data parallel (`Process_Group`, one process per GPU, NCCL): the graph then holds the whole {sample -> loss -> all-reduce -> Adam}
step of SURVEY 8(f) rank 2.  Rank g draws rows [g*n, (g+1)*n) of the epoch's Philox stream (n = Num_Train_Coll_Points / world),
keeps the targeted rows of its own shard, and two collectives are captured per epoch: one of 5 doubles per data set (row count,
sum r^2, the |r| moments of the random rows, the data-loss partial) and one of the flat gradient buffer [networks | xi].  The local
means the kernels produce are rescaled by rows_local / rows_global on the device before the gradient all-reduce, the targeted
cutoff uses the global moments (pdl_select_targeted_dev), and every rank applies the same Adam update: replicas stay identical and
the numbers equal the one-process run.  `Inputs_List` / `Targets_List` are this rank's shard of the data rows.
"""
import ctypes as C
from typing import Dict, List

import numpy
import torch

from . import _lib as L
from .engine import collocation_plan, data_plan, device_mask, network_param_tensors, _check_cuda_f32


class Captured_Epochs:
    def __init__(self, U_List: List, Xi: torch.Tensor, Mask: torch.Tensor, Inputs_List: List[torch.Tensor],
                 Targets_List: List[torch.Tensor], Bounds_List: List[numpy.ndarray], Derivatives: List, LHS_Term, RHS_Terms: List,
                 p: float, Weights: Dict[str, float], Optimizer: torch.optim.Optimizer, Num_Train_Coll_Points: int,
                 Max_Targeted: int = None, Seed: int = 0, History_Capacity: int = 4096, First_Epoch: int = 0, Process_Group=None):
        assert len(U_List) == len(Inputs_List) == len(Targets_List) == len(Bounds_List)
        assert torch.numel(Xi) == len(RHS_Terms)
        assert 0 < p < 2                                                    # Loss.py:279
        if not isinstance(Optimizer, torch.optim.Adam):
            raise TypeError("Captured_Epochs fuses the Adam update; use Run_Epochs/Training for other optimisers")
        self.group, self.rank, self.world = Process_Group, 0, 1
        if Process_Group is not None:
            import torch.distributed as dist
            self.rank, self.world = dist.get_rank(Process_Group), dist.get_world_size(Process_Group)
            if int(Num_Train_Coll_Points) % self.world:
                raise ValueError("Num_Train_Coll_Points must be a multiple of the world size (%d)" % self.world)
        self.S, self.N_global = len(U_List), int(Num_Train_Coll_Points)
        self.N = self.N_global // self.world                                # random rows this rank draws per epoch
        self.cap_t = int(Max_Targeted) if Max_Targeted is not None else self.N
        self.U_List, self.Xi, self.opt = U_List, Xi, Optimizer
        self.p, self.W = float(p), dict(Weights)
        self.dev = dev = Xi.device
        if dev.type != "cuda":
            raise RuntimeError("Captured_Epochs needs CUDA tensors: this engine has no CPU path")
        self.K = len(RHS_Terms)
        self.mask_dev = device_mask(Mask, self.K, dev)
        self.coll_plans = [collocation_plan(U, Derivatives, LHS_Term, RHS_Terms) for U in U_List]
        self.data_plans = [data_plan(U) for U in U_List]
        self.params = [[_check_cuda_f32(q, "network parameter") for q in network_param_tensors(U)] for U in U_List]
        _check_cuda_f32(Xi, "Xi")
        self.inputs = [_check_cuda_f32(X, "Inputs").detach() for X in Inputs_List]
        self.targets = [Y.reshape(-1).to(device=dev, dtype=torch.float64).contiguous() for Y in Targets_List]
        self.bounds = [numpy.ascontiguousarray(numpy.asarray(B, dtype=numpy.float64).reshape(-1)) for B in Bounds_List]
        self.d = int(len(self.bounds[0]) // 2)
        for U in U_List:
            U.train()

        # ---- optimiser state, shared with torch.optim.Adam ----
        groups = self.opt.param_groups
        if len(groups) != 1:
            raise ValueError("Captured_Epochs expects one parameter group (Code/main.py:232-235)")
        g = groups[0]
        if g.get("weight_decay", 0) != 0 or g.get("amsgrad", False) or g.get("maximize", False):
            raise ValueError("Captured_Epochs implements plain Adam (no weight decay / amsgrad / maximize)")
        self.lr, (self.b1, self.b2), self.eps = float(g["lr"]), g["betas"], float(g["eps"])
        mine = [q for prm in self.params for q in prm] + [Xi]
        theirs = {id(q) for q in g["params"]}
        if {id(q) for q in mine} != theirs:
            raise ValueError("the optimiser must hold exactly the parameters of U_List and Xi")
        steps = set()
        for q in mine:
            st = self.opt.state[q]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(q, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(q, memory_format=torch.preserve_format)
            steps.add(float(st["step"]))
        if len(steps) != 1:
            raise ValueError("parameters with different Adam step counts")
        self.step0 = steps.pop()
        self.step_dev = torch.full((1,), self.step0, dtype=torch.float32, device=dev)

        # ---- static device buffers ----
        S, N, cap, d, K = self.S, self.N, self.N + self.cap_t, self.d, self.K
        i64 = dict(dtype=torch.int64, device=dev)
        self.epoch_dev = torch.zeros(1, **i64)                               # epochs replayed so far
        self.seeds = torch.zeros((S, 1), **i64)                              # Philox keys (uint64 bit patterns)
        base = (int(Seed) * 1000003 + int(First_Epoch)) * 64                 # the keys Run_Epochs uses: (Seed*1000003 + e)*64 + i
        for i in range(S):
            self.seeds[i, 0] = ((base + i) & 0xFFFFFFFFFFFFFFFF) - (1 << 64 if ((base + i) & 0xFFFFFFFFFFFFFFFF) >= (1 << 63) else 0)
        self.coll = [torch.zeros((cap, d), dtype=torch.float32, device=dev) for _ in range(S)]
        self.count = [torch.full((1,), N, **i64) for _ in range(S)]          # rows in use: N random + targeted
        self.tbuf = [torch.zeros((self.cap_t, d), dtype=torch.float32, device=dev) for _ in range(S)]
        self.tcount = [torch.zeros(1, **i64) for _ in range(S)]
        self.tstats = [torch.zeros(4, dtype=torch.float64, device=dev) for _ in range(S)]
        self.tscratch = [torch.zeros((cap + 1023) // 1024 + 1, dtype=torch.int32, device=dev) for _ in range(S)]
        self.resid = [torch.zeros(cap, dtype=torch.float32, device=dev) for _ in range(S)]
        self.n_scal = [pl.n_param_scalars for pl in self.coll_plans]
        # all gradients of an epoch in ONE flat buffer [network 0 | network 1 | ... | xi]: the unit of the gradient all-reduce
        self.flat = torch.zeros(sum(self.n_scal) + max(K, 1), dtype=torch.float32, device=dev)
        offs = [0]
        for n in self.n_scal:
            offs.append(offs[-1] + n)
        self.gc = [self.flat[offs[i]:offs[i + 1]] for i in range(S)]
        self.red = torch.zeros((S, 5), dtype=torch.float64, device=dev)     # per data set: rows, sum r^2, sum |r|, sum r^2 (random rows), data partial
        self.mom = torch.zeros((S, 3), dtype=torch.float64, device=dev)     # global moments handed to pdl_select_targeted_dev
        self.n_data = [int(X.shape[0]) for X in self.inputs]
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor(self.n_data, dtype=torch.int64, device=dev)
            dist.all_reduce(t, group=self.group)
            self.n_data = [int(v) for v in t.tolist()]
        self.gd = [torch.zeros(n, dtype=torch.float32, device=dev) for n in self.n_scal]
        self.g2 = [torch.zeros(n, dtype=torch.float32, device=dev) for n in self.n_scal]
        self.gx = torch.zeros((S, max(K, 1)), dtype=torch.float32, device=dev)
        self.gxi = self.flat[offs[S]:offs[S] + max(K, 1)]
        self.g_lp = torch.zeros(max(K, 1), dtype=torch.float32, device=dev)
        self.lp = torch.zeros(1, dtype=torch.float32, device=dev)
        self.l2 = torch.zeros(S, dtype=torch.float32, device=dev)
        self.stats_c = torch.zeros((S, 4), dtype=torch.float64, device=dev)
        self.stats_d = torch.zeros((S, 4), dtype=torch.float64, device=dev)
        self.scal = torch.zeros(4 * S + 1, dtype=torch.float64, device=dev)   # pdl_closure_tail: coll_*, data_*, l2_*, total_*, lp
        self.total32 = torch.zeros(1, dtype=torch.float32, device=dev)
        self.hcap = int(History_Capacity)
        self.hist = torch.zeros((self.hcap, 5 * S + 1), dtype=torch.float64, device=dev)   # coll, data, l2, total, n_targeted | lp
        self._read = 0                                                       # epochs already returned by history()
        self._done = 0
        self._adam_tables()
        self.graph = None

    # ------------------------------------------------------------------------------------------
    def _adam_tables(self):
        th, g, m, v, n = [], [], [], [], []
        for i, prm in enumerate(self.params):
            off = 0
            for q in prm:
                st = self.opt.state[q]
                th.append(q.data_ptr()); g.append(self.gc[i].data_ptr() + 4 * off)
                m.append(st["exp_avg"].data_ptr()); v.append(st["exp_avg_sq"].data_ptr()); n.append(q.numel())
                off += q.numel()
            assert off == self.n_scal[i]
        st = self.opt.state[self.Xi]
        th.append(self.Xi.data_ptr()); g.append(self.gxi.data_ptr()); m.append(st["exp_avg"].data_ptr())
        v.append(st["exp_avg_sq"].data_ptr()); n.append(self.K)
        self._adam = []
        for a in range(0, len(th), 48):
            sl = slice(a, a + 48)
            k = len(th[sl])
            self._adam.append(((C.c_void_p * k)(*th[sl]), (C.c_void_p * k)(*g[sl]), (C.c_void_p * k)(*m[sl]), (C.c_void_p * k)(*v[sl]),
                               (C.c_int64 * k)(*n[sl]), k))

    def _epoch(self):
        """One epoch, enqueued on the current stream (captured once, replayed afterwards)."""
        S, N, K, W, d, world = self.S, self.N, self.K, self.W, self.d, self.world
        lib = L.lib()
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        if world > 1:
            import torch.distributed as dist
        with torch.no_grad():
            L.check(lib.pdl_lp_loss(self.Xi.data_ptr(), self.mask_dev.data_ptr(), K, self.p, self.lp.data_ptr(), self.g_lp.data_ptr(), stream))
            for i in range(S):
                L.check(lib.pdl_sample_points_dev(self.bounds[i].ctypes.data_as(C.POINTER(C.c_double)), d, N, self.seeds[i].data_ptr(),
                                                  self.rank * N, self.coll[i].data_ptr(), stream))
                prm = self.params[i]
                self.coll_plans[i].launch(self.coll[i], prm, self.Xi, self.mask_dev, None, 0.0, True, self.resid[i], None, self.gc[i],
                                          self.gx[i], self.stats_c[i], n_points_dev=self.count[i])
                self.data_plans[i].launch(self.inputs[i], prm, None, None, self.targets[i], 1.0 / max(self.n_data[i], 1),
                                          True, None, None, self.gd[i], None, self.stats_d[i])
                n = len(prm)
                ptrs = (C.c_void_p * n)(*[q.data_ptr() for q in prm])
                sizes = (C.c_int32 * n)(*[int(q.numel()) for q in prm])
                L.check(lib.pdl_l2_squared_loss(ptrs, sizes, n, self.l2[i:i + 1].data_ptr(), self.g2[i].data_ptr(), stream))
                if world > 1:                                   # this shard's share of the global sums
                    L.check(lib.pdl_abs_residual_moments(self.resid[i].data_ptr(), N, self.tstats[i].data_ptr(), stream))
                    self.red[i, 0:1].copy_(self.count[i])
                    self.red[i, 1:2].copy_(self.stats_c[i, 1:2])
                    self.red[i, 2:4].copy_(self.tstats[i][0:2])
                    self.red[i, 4:5].copy_(self.stats_d[i, 0:1])
            if world > 1:
                rows_local = torch.cat(self.count).to(torch.float64)
                dist.all_reduce(self.red, group=self.group)    # collective 1: 5 doubles per data set
                ratio = (rows_local / self.red[:, 0]).to(torch.float32)      # local mean -> share of the global mean
                coll, data = self.red[:, 1] / self.red[:, 0], self.red[:, 4]
                self.mom[:, 0:2].copy_(self.red[:, 2:4]); self.mom[:, 2].fill_(float(self.N_global))
                self.gx.mul_(ratio.unsqueeze(1))
            else:
                coll, data = self.stats_c[:, 0], self.stats_d[:, 0]
            for i in range(S):
                if world > 1:
                    self.gc[i].mul_(ratio[i])
                L.check(lib.pdl_axpby3(self.gc[i].data_ptr(), self.n_scal[i], self.gc[i].data_ptr(), W["Coll"], self.gd[i].data_ptr(),
                                       W["Data"], self.g2[i].data_ptr(), W["L2"] / world, stream))
            if world > 1:
                torch.add(self.gx.sum(0) * W["Coll"], self.g_lp, alpha=S * W["Lp"] / world, out=self.gxi)
                dist.all_reduce(self.flat, group=self.group)   # collective 2: the flat gradient buffer [networks | xi]
                # losses of this epoch (before the update, like Training's return value) -> history row
                l2d, lpd = self.l2.to(torch.float64), self.lp.to(torch.float64)
                totals = W["Data"] * data + W["Coll"] * coll + W["Lp"] * lpd + W["L2"] * l2d
            else:
                # one launch: weighted xi gradient + the epoch's loss vector [coll_*, data_*, l2_*, total_*, lp]
                L.check(lib.pdl_closure_tail(self.stats_c.data_ptr(), self.stats_d.data_ptr(), self.l2.data_ptr(), self.lp.data_ptr(),
                                             self.gx.data_ptr(), self.g_lp.data_ptr(), S, K, max(K, 1), W["Data"], W["Coll"], W["Lp"], W["L2"],
                                             S * W["Lp"], self.scal.data_ptr(), self.total32.data_ptr(),
                                             self.gxi.data_ptr() if K > 0 else None, stream))
                l2d, totals, lpd = self.scal[2 * S:3 * S], self.scal[3 * S:4 * S], self.scal[4 * S:4 * S + 1]
            # Adam (Code/Test_Train.py:193)
            self.step_dev += 1.0
            for th, g, m, v, n, k in self._adam:
                L.check(lib.pdl_adam_step_multi(th, g, m, v, n, k, self.lr, self.b1, self.b2, self.eps, self.step_dev.data_ptr(), stream))
            # targeted points for the next epoch (Code/main.py:365-384)
            for i in range(S):
                if world > 1:
                    L.check(lib.pdl_select_targeted_dev(self.coll[i].data_ptr(), self.resid[i].data_ptr(), N + self.cap_t,
                                                        self.count[i].data_ptr(), d, self.mom[i].data_ptr(), self.tbuf[i].data_ptr(),
                                                        self.cap_t, self.tcount[i].data_ptr(), self.tstats[i].data_ptr(),
                                                        self.tscratch[i].data_ptr(), stream))
                else:
                    L.check(lib.pdl_targeted_update_dev(self.coll[i].data_ptr(), self.resid[i].data_ptr(), N + self.cap_t,
                                                        self.count[i].data_ptr(), N, d, self.tbuf[i].data_ptr(), self.cap_t,
                                                        self.tcount[i].data_ptr(), self.tstats[i].data_ptr(), self.tscratch[i].data_ptr(), stream))
                self.coll[i][N:].copy_(self.tbuf[i])
                torch.add(self.tcount[i], N, out=self.count[i])
            # rows selected before clamping to the carry-over capacity (stats[3], int64 bit pattern): history() raises when rows were lost
            nt = torch.cat([st[3:4].view(torch.int64) for st in self.tstats]).to(torch.float64)
            row = torch.cat([coll, data, l2d, totals, nt, lpd]).unsqueeze(0)
            self.hist.index_copy_(0, self.epoch_dev % self.hcap, row)
            self.epoch_dev += 1
            self.seeds += 64

    # ------------------------------------------------------------------------------------------
    def capture(self):
        """Warm up (module load, function attributes, workspaces) on a side stream, restore the state, then capture."""
        if self.graph is not None:
            return self
        if getattr(self, "_closed", False):
            raise RuntimeError("this Captured_Epochs was closed")
        snap = self._snapshot()
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._epoch()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self._restore(snap)
        self.graph = torch.cuda.CUDAGraph()
        # with NCCL collectives inside, other threads of the process (the process group's watchdog) keep making CUDA calls while this
        # thread captures: only this thread's calls belong to the capture
        mode = {"capture_error_mode": "thread_local"} if self.world > 1 else {}
        with torch.cuda.graph(self.graph, **mode):
            self._epoch()
        self._restore(snap)          # capture does not execute, but keep the invariant explicit
        return self

    def _state_tensors(self):
        out = [q for prm in self.params for q in prm] + [self.Xi, self.step_dev, self.epoch_dev, self.seeds, self.hist]
        for q in [q for prm in self.params for q in prm] + [self.Xi]:
            st = self.opt.state[q]
            out += [st["exp_avg"], st["exp_avg_sq"]]
        out += self.coll + self.count + self.tcount
        return out

    def _snapshot(self):
        return [t.detach().clone() for t in self._state_tensors()]

    def _restore(self, snap):
        with torch.no_grad():
            for t, c in zip(self._state_tensors(), snap):
                t.copy_(c)

    def run(self, Num_Epochs: int) -> None:
        """Replays the captured epoch Num_Epochs times; no host synchronisation."""
        self.capture()
        if self._done + int(Num_Epochs) - self._read > self.hcap:
            raise RuntimeError("history ring buffer would overflow: call history() at least every %d epochs" % self.hcap)
        for _ in range(int(Num_Epochs)):
            self.graph.replay()
        self._done += int(Num_Epochs)
        for q in [q for prm in self.params for q in prm] + [self.Xi]:        # what torch.optim.Adam would hold
            self.opt.state[q]["step"] = torch.tensor(self.step0 + self._done, dtype=torch.float32)

    def close(self) -> None:
        """Drops the captured graph (state and history stay readable).  With a Process_Group call this -- or let the object go out
        of scope -- BEFORE torch.distributed.destroy_process_group(): a live graph that contains NCCL kernels keeps the communicator
        busy and the teardown waits for it forever (observed on B200 with NCCL 2.28)."""
        torch.cuda.synchronize(self.dev)
        self.graph = None
        self._closed = True

    def history(self) -> Dict[str, numpy.ndarray]:
        """Losses of the epochs replayed since the last call (one device->host copy; synchronises)."""
        S = self.S
        idx = torch.arange(self._read, self._done, device=self.dev) % self.hcap
        h = self.hist.index_select(0, idx).cpu().numpy()
        self._read = self._done
        out = {"Train Coll": h[:, 0:S], "Train Data": h[:, S:2 * S], "L2": h[:, 2 * S:3 * S], "Train Total": h[:, 3 * S:4 * S],
               "Num Targeted": h[:, 4 * S:5 * S].astype(numpy.int64), "Lp": h[:, 5 * S]}
        if out["Num Targeted"].size and int(out["Num Targeted"].max()) > self.cap_t:
            first = int(numpy.argmax(out["Num Targeted"].max(axis=1) > self.cap_t))
            raise RuntimeError("targeted-point capacity exceeded: epoch %d selected %d rows, %d can be carried over; the epochs after "
                               "it trained on a truncated set -- rebuild with a larger Max_Targeted"
                               % (self._read - h.shape[0] + first, int(out["Num Targeted"].max()), self.cap_t))
        return out

    def targeted_points(self, i: int = 0) -> torch.Tensor:
        """The targeted points carried into the next epoch (copy; synchronises)."""
        n = int(self.tcount[i].item())
        return self.coll[i][self.N:self.N + n].clone()
