"""
Epoch-level pieces around the hot path: the targeted-collocation update, the epoch loop of
Code/main.py:281-384 and the save format of Code/main.py:487-524, all with device-resident data.

    Update_Targeted_Points(Coll_Points, Residual, Num_Random)       main.py:365-384  (one 8-byte D2H per call)
    Run_Epochs(...)                                                 main.py:281-405  (sampling, Training, Testing)
    Threshold_Xi / Build_Mask                                       main.py:31,212-216,455-460
    Save_State / Load_State                                         main.py:51-57,487-524
"""
import ctypes as C
from typing import Callable, Dict, List, Optional

import numpy
import torch

from . import _lib as L
from .Network import Network
from .Points import Generate_Points
from .Term import Build_Term_From_State
from .Derivative import Derivative
from .Test_Train import Testing, Training

THRESHOLD = 0.0005           # Code/main.py:31


def Update_Targeted_Points(Coll_Points: torch.Tensor, Residual: torch.Tensor, Num_Random: int, Process_Group=None) -> torch.Tensor:
    """
    Rows of Coll_Points whose |residual| >= mean + 3*std, the moments taken over the first Num_Random rows
    (the freshly sampled points) and the selection over all rows, order preserved (Code/main.py:365-384).
    Everything runs on the device; only the number of kept rows is read back.
    With `Process_Group` the rows are this rank's shard: sum |r|, sum r^2 and the row count of the random rows are all-reduced
    (3 doubles) so that every rank applies the cutoff of the GLOBAL batch -- the targeted set is then the same for any number of
    ranks -- and each rank compacts its own shard.
    """
    P = Coll_Points.detach()
    r = Residual.detach()
    if not (P.is_cuda and r.is_cuda):
        raise RuntimeError("Update_Targeted_Points works on device tensors (Training(..., Residuals_Device='cuda'))")
    P = P.contiguous(); r = r.to(torch.float32).contiguous()
    n, d = int(P.shape[0]), int(P.shape[1])
    assert r.numel() == n and 0 <= Num_Random <= n
    out = torch.empty_like(P)
    cnt = torch.zeros(1, dtype=torch.int64, device=P.device)
    stats = torch.zeros(4, dtype=torch.float64, device=P.device)
    scratch = torch.empty((n + 1023) // 1024 + 1, dtype=torch.int32, device=P.device)
    stream = torch.cuda.current_stream(P.device).cuda_stream
    with torch.cuda.device(P.device):
        if Process_Group is None:
            L.check(L.lib().pdl_targeted_update(P.data_ptr(), r.data_ptr(), n, int(Num_Random), d, out.data_ptr(), cnt.data_ptr(),
                                                stats.data_ptr(), scratch.data_ptr(), stream))
        else:
            import torch.distributed as dist
            L.check(L.lib().pdl_abs_residual_moments(r.data_ptr(), int(Num_Random), stats.data_ptr(), stream))
            glob = torch.stack([stats[0], stats[1], torch.tensor(float(Num_Random), dtype=torch.float64, device=P.device)])
            dist.all_reduce(glob, group=Process_Group)
            s1, s2, m = (float(v) for v in glob.cpu())
            cutoff = targeted_cutoff(s1, s2, m)
            L.check(L.lib().pdl_select_targeted(P.data_ptr(), r.data_ptr(), n, d, C.c_float(cutoff), out.data_ptr(), cnt.data_ptr(),
                                                scratch.data_ptr(), stream))
    return out[:int(cnt.item())]


def targeted_cutoff(sum_abs: float, sum_sq: float, count: float) -> float:
    """mean + 3 * unbiased std of |r| from its sums (Code/main.py:369-380), rounded to float32 like the device cutoff kernel."""
    mean = sum_abs / count if count > 0 else 0.0
    var = (sum_sq - count * mean * mean) / (count - 1.0) if count > 1 else 0.0
    return float(numpy.float32(mean + 3.0 * numpy.sqrt(max(var, 0.0))))


def Build_Mask(Xi: torch.Tensor, Threshold: float = THRESHOLD) -> torch.Tensor:
    """Mask[k] = |Xi[k]| < Threshold: True EXCLUDES the term (Code/main.py:212-216, Loss.py:216-218)."""
    return (Xi.detach().abs() < Threshold).cpu()


def Threshold_Xi(Xi: torch.Tensor, Mask: torch.Tensor, Threshold: float = THRESHOLD) -> List[int]:
    """Indices of the RHS terms that survive: not masked and |xi_k| >= Threshold (Code/main.py:455-460)."""
    x = Xi.detach().cpu(); m = Mask.cpu() if isinstance(Mask, torch.Tensor) else torch.as_tensor(Mask)
    return [k for k in range(x.numel()) if (not bool(m[k])) and abs(float(x[k])) >= Threshold]


def Run_Epochs(U_List: List[Network], Xi: torch.Tensor, Mask: torch.Tensor, Inputs_List: List[torch.Tensor],
               Targets_List: List[torch.Tensor], Bounds_List: List[numpy.ndarray], Derivatives: List, LHS_Term, RHS_Terms: List,
               p: float, Weights: Dict[str, float], Optimizer: torch.optim.Optimizer, Num_Epochs: int,
               Num_Train_Coll_Points: int, Num_Test_Coll_Points: int = 0, Test_Inputs_List=None, Test_Targets_List=None,
               Point_Sampler: Optional[Callable] = None, Process_Group=None, Seed: int = 0,
               On_Epoch: Optional[Callable] = None, Use_CUDA_Graph: bool = False, Max_Targeted: Optional[int] = None) -> Dict:
    """
    The epoch loop of Code/main.py:281-405: per epoch and per data set sample Num_Train_Coll_Points random points,
    append last epoch's targeted points, run Training, optionally Testing on fresh points, update the targeted points.
    `Point_Sampler(data_set_index, epoch, n)` may supply the random points (default: on-device Philox sampler).
    Returns the loss histories (numpy arrays, one row per epoch).
    `Use_CUDA_Graph=True` (Adam, default sampler) replays the whole epoch as one captured CUDA graph
    (Epoch_Graph.Captured_Epochs): same numbers, no per-epoch host work; epochs are replayed in chunks of 10 unless Testing
    is requested every epoch.  Together with `Process_Group` the graph contains the two NCCL all-reduces of the data-parallel step.
    `Process_Group` (one process per GPU): the epoch's Num_Train_Coll_Points random rows and the data rows are SHARDED -- rank g
    draws global rows [g*n, (g+1)*n) of the same Philox stream (n = Num_Train_Coll_Points / world, which must divide), takes its
    slice of Inputs/Targets, keeps the targeted points of its own shard, and the targeted cutoff uses the all-reduced moments:
    the numbers equal the one-process run at 1/world of the work per rank.  A user `Point_Sampler` receives the per-rank count
    and must return this rank's rows.
    """
    S = len(U_List)
    dev = Xi.device
    d = int(Bounds_List[0].shape[0])
    rank, world = 0, 1
    if Process_Group is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(Process_Group), dist.get_world_size(Process_Group)
        if Num_Train_Coll_Points % world or (Num_Test_Coll_Points % world):
            raise ValueError("Num_Train_Coll_Points and Num_Test_Coll_Points must be multiples of the world size (%d)" % world)
        shard = lambda T: T[rank * (T.shape[0] // world):(rank + 1) * (T.shape[0] // world) if rank + 1 < world else T.shape[0]].contiguous()
        Inputs_List = [shard(X) for X in Inputs_List]
        Targets_List = [shard(Y) for Y in Targets_List]
    n_loc, n_test_loc = Num_Train_Coll_Points // world, Num_Test_Coll_Points // world
    targeted = [torch.empty((0, d), dtype=torch.float32, device=dev) for _ in range(S)]
    hist = {"Train Coll": numpy.zeros((Num_Epochs, S)), "Train Data": numpy.zeros((Num_Epochs, S)),
            "Train Total": numpy.zeros((Num_Epochs, S)), "L2": numpy.zeros((Num_Epochs, S)), "Lp": numpy.zeros(Num_Epochs),
            "Test Coll": numpy.zeros((Num_Epochs, S)), "Test Data": numpy.zeros((Num_Epochs, S)),
            "Num Targeted": numpy.zeros((Num_Epochs, S), dtype=numpy.int64)}
    sample = Point_Sampler or (lambda i, e, n: Generate_Points(Bounds_List[i], n, dev, Seed=(Seed * 1000003 + e) * 64 + i,
                                                               First_Index=rank * n))
    if Use_CUDA_Graph:
        if Point_Sampler is not None:
            raise ValueError("Use_CUDA_Graph works with the built-in sampler")
        from .Epoch_Graph import Captured_Epochs
        cap = Captured_Epochs(U_List, Xi, Mask, Inputs_List, Targets_List, Bounds_List, Derivatives, LHS_Term, RHS_Terms, p, Weights,
                              Optimizer, Num_Train_Coll_Points, Max_Targeted=Max_Targeted, Seed=Seed, Process_Group=Process_Group)
        chunk = 1 if Num_Test_Coll_Points > 0 else 10
        e = drained = 0
        while e < Num_Epochs:
            m = min(chunk, Num_Epochs - e)
            cap.run(m)
            if Num_Test_Coll_Points > 0:
                tp = [sample(i, Num_Epochs + e, n_test_loc).to(dev) for i in range(S)]
                te = Testing(U_List, Xi, Mask, tp, Test_Inputs_List or Inputs_List, Test_Targets_List or Targets_List, Derivatives,
                             LHS_Term, RHS_Terms, p, Weights)
                hist["Test Coll"][e] = te["Coll Losses"]; hist["Test Data"][e] = te["Data Losses"]
            if On_Epoch is not None or e + m >= Num_Epochs or (e + m - drained) + chunk > cap.hcap:
                h = cap.history()                       # drained at least once per History_Capacity epochs (ring buffer)
                drained = e + m
                a = e + m - h["Lp"].shape[0]
                for key in ("Train Coll", "Train Data", "Train Total", "L2", "Lp", "Num Targeted"):
                    hist[key][a:e + m] = h[key]
                if On_Epoch is not None:
                    for q in range(e, e + m):
                        On_Epoch(q, hist)
            e += m
        if Process_Group is not None:
            cap.close()                  # a live graph holding NCCL kernels would block a later destroy_process_group()
        return hist
    for e in range(Num_Epochs):
        coll = [torch.vstack((sample(i, e, n_loc).to(dev), targeted[i])) for i in range(S)]
        tr = Training(U_List, Xi, Mask, coll, Inputs_List, Targets_List, Derivatives, LHS_Term, RHS_Terms, p, Weights, Optimizer,
                      Process_Group=Process_Group, Residuals_Device="cuda")
        hist["Train Coll"][e] = tr["Coll Losses"]; hist["Train Data"][e] = tr["Data Losses"]
        hist["Train Total"][e] = tr["Total Losses"]; hist["L2"][e] = tr["L2 Losses"]; hist["Lp"][e] = tr["Lp Loss"]
        if Num_Test_Coll_Points > 0:
            tp = [sample(i, Num_Epochs + e, n_test_loc).to(dev) for i in range(S)]
            te = Testing(U_List, Xi, Mask, tp, Test_Inputs_List or Inputs_List, Test_Targets_List or Targets_List, Derivatives,
                         LHS_Term, RHS_Terms, p, Weights)          # per-rank means over this rank's test rows
            hist["Test Coll"][e] = te["Coll Losses"]; hist["Test Data"][e] = te["Data Losses"]
        for i in range(S):
            targeted[i] = Update_Targeted_Points(coll[i], tr["Residuals"][i], n_loc, Process_Group=Process_Group)
            hist["Num Targeted"][e, i] = targeted[i].shape[0]          # this rank's share
        if On_Epoch is not None:
            On_Epoch(e, hist)
    return hist


def Save_State(Path: str, U_List: List[Network], Xi: torch.Tensor, Optimizer, Derivatives: List, LHS_Term, RHS_Terms: List,
               DataSet_Names: List[str]) -> None:
    """Same dictionary as the reference's save file (Code/main.py:505-524), so either code base can resume it."""
    torch.save({"U States": [U.Get_State() for U in U_List],
                "Xi": Xi,
                "Optimizer": Optimizer.state_dict(),
                "Derivative Encodings": [D.Encoding for D in Derivatives],
                "LHS Term State": LHS_Term.Get_State(),
                "RHS Term States": [T.Get_State() for T in RHS_Terms],
                "DataSet Names": list(DataSet_Names)}, Path)


def Load_State(Path: str, Device: torch.device, Hidden_Activation: Optional[str] = None) -> Dict:
    """
    Loads a save written by Save_State or by the reference (Code/main.py:51-57; pickled numpy arrays inside, hence
    weights_only=False).  Returns {"U List", "Xi", "Optimizer State", "Derivatives", "LHS Term", "RHS Terms", "DataSet Names"}.
    """
    blob = torch.load(Path, map_location=Device, weights_only=False)
    U_List = []
    for st in blob["U States"]:
        act = Hidden_Activation or st["Activation Types"][0]
        U = Network(Widths=list(st["Widths"]), Hidden_Activation=act, Output_Activation=st["Activation Types"][-1], Device=Device)
        U.Set_State(st)
        U_List.append(U.to(Device))
    Xi = blob["Xi"].detach().to(Device).requires_grad_(True)
    return {"U List": U_List, "Xi": Xi, "Optimizer State": blob["Optimizer"],
            "Derivatives": [Derivative(Encoding=numpy.asarray(e)) for e in blob["Derivative Encodings"]],
            "LHS Term": Build_Term_From_State(blob["LHS Term State"]),
            "RHS Terms": [Build_Term_From_State(s) for s in blob["RHS Term States"]],
            "DataSet Names": blob["DataSet Names"]}
