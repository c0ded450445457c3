"""
Command-line driver with the reference's workflow (Code/main.py:35-539) on top of the B200 engine:

    python -m pde_learn_b200.main --settings Settings.txt [--data-dir Data/DataSets] [--saves-dir Saves]

Reads the reference's Settings.txt / Library.txt formats, loads the `.npz` data sets written by the reference's
Data/Create_Data_Set.py ("Train_Inputs", "Train_Targets", "Test_Inputs", "Test_Targets", "Input_Bounds"), builds one network
per data set and a shared Xi, optionally resumes from a save file (U / Xi+library / optimiser, masking |xi| < 5e-4), runs the
epochs (random + targeted collocation points, Training, Testing), restores masked coefficients, prints the identified PDE and
writes a save file in the reference's format.  Plots are out of scope.  The orchestration is thin on purpose: every number
comes from the fused CUDA kernels (Loss.py / Test_Train.py / Driver.py).
"""
import argparse
import os
import time
from typing import Dict, List

import numpy
import torch

from .Driver import Build_Mask, Load_State, Run_Epochs, Save_State, THRESHOLD
from .Library_Reader import Read_Library
from .Network import Network
from .Settings_Reader import Settings_Reader


def Data_Loader(DataSet_Name: str, Device: torch.device, Data_Dir: str) -> Dict:
    """Same dictionary as the reference's Data_Loader (Code/Data.py:6-52)."""
    D = numpy.load(os.path.join(Data_Dir, DataSet_Name + ".npz"))
    return {"Train Inputs": torch.from_numpy(D["Train_Inputs"]).to(device=Device, dtype=torch.float32),
            "Train Targets": torch.from_numpy(D["Train_Targets"]).to(device=Device),
            "Test Inputs": torch.from_numpy(D["Test_Inputs"]).to(device=Device, dtype=torch.float32),
            "Test Targets": torch.from_numpy(D["Test_Targets"]).to(device=Device),
            "Input Bounds": D["Input_Bounds"],
            "Number of Dimensions": D["Input_Bounds"].shape[0]}


def Format_PDE(LHS_Term, RHS_Terms: List, Xi: torch.Tensor, Mask: torch.Tensor, Threshold: float = THRESHOLD) -> str:
    """The identified PDE as the reference prints it (Code/main.py:454-471)."""
    out = str(LHS_Term) + " = "
    printed = 0
    x = Xi.detach().cpu()
    for k, T in enumerate(RHS_Terms):
        if bool(Mask[k]) or abs(float(x[k])) < Threshold:
            continue
        if printed != 0 and float(x[k]) > 0:
            out += " + "
        elif float(x[k]) < 0:
            out += " - "
        out += "%7.4f" % abs(float(x[k])) + str(T)
        printed += 1
    return out


def main(argv=None) -> Dict:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--settings", default="Settings.txt")
    ap.add_argument("--data-dir", default=None, help="directory of the .npz data sets (default: <settings dir>/Data/DataSets)")
    ap.add_argument("--saves-dir", default=None, help="where save files live (default: <settings dir>/Saves)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--cuda-graph", action="store_true", help="replay each epoch as one captured CUDA graph (Adam only)")
    args = ap.parse_args(argv)
    root = os.path.dirname(os.path.abspath(args.settings))
    data_dir = args.data_dir or os.path.join(root, "Data", "DataSets")
    saves_dir = args.saves_dir or os.path.join(root, "Saves")
    S = Settings_Reader(args.settings)
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(args.seed)

    if not (S["Load U"] or S["Load Xi, Library"]):
        # fail before any data is read when the architecture / library is outside the kernels' limits (d is at least the longest
        # derivative encoding; the exact d is checked again by the plan itself once the data bounds are known)
        from .engine import validate_configuration
        Dv, Lh, Rh = Read_Library(S["Library Path"])
        validate_configuration(max(2, max(len(D.Encoding) for D in Dv)), S["Hidden Layer Widths"], S["Hidden Activation Function"], Dv, Lh, Rh)

    saved = None
    if S["Load U"] or S["Load Xi, Library"] or S["Load Optimizer"]:
        saved = Load_State(os.path.join(saves_dir, S["Load File Name"]), dev)
    names = saved["DataSet Names"] if S["Load U"] else S["DataSet Names"]
    data = [Data_Loader(n, dev, data_dir) for n in names]
    d = int(data[0]["Number of Dimensions"])

    if S["Load U"]:
        U_List = saved["U List"]
        act = type(list(U_List[0].Activation_Functions)[0]).__name__
    else:
        act = S["Hidden Activation Function"]
        U_List = [Network(Widths=[d] + S["Hidden Layer Widths"] + [1], Hidden_Activation=act, Output_Activation="None", Device=dev)
                  for _ in names]
    if S["Load Xi, Library"]:
        Xi, Derivs, LHS, RHS = saved["Xi"], saved["Derivatives"], saved["LHS Term"], saved["RHS Terms"]
    else:
        Derivs, LHS, RHS = Read_Library(S["Library Path"])
        Xi = torch.zeros(len(RHS), dtype=torch.float32, device=dev, requires_grad=True)
    K = len(RHS)
    Mask = torch.zeros(K, dtype=torch.bool)
    if S["Load Xi, Library"] and S["Mask Small Xi Components"]:
        Mask = Build_Mask(Xi)                                          # main.py:212-216
    Initial_Xi = Xi.detach().clone()

    params = [q for U in U_List for q in U.parameters()] + [Xi]
    if S["Optimizer"] == "Adam":
        opt = torch.optim.Adam(params, lr=S["Learning Rate"])
    else:
        opt = torch.optim.LBFGS(params, lr=S["Learning Rate"])
    if S["Load Optimizer"]:
        opt.load_state_dict(saved["Optimizer State"])
        for g in opt.param_groups:
            g["lr"] = S["Learning Rate"]                               # main.py:244-246

    def report(e, h):
        if args.quiet or (e + 1) % 10 != 0:
            return
        for i in range(len(U_List)):
            print("Epoch #%-4d | Train: Data[%u] = %.7f  Coll[%u] = %.7f  Total[%u] = %.7f | Test: Data = %.7f Coll = %.7f | targeted %d"
                  % (e + 1, i, h["Train Data"][e, i], i, h["Train Coll"][e, i], i, h["Train Total"][e, i], h["Test Data"][e, i],
                     h["Test Coll"][e, i], h["Num Targeted"][e, i]))

    t0 = time.perf_counter()
    hist = Run_Epochs(U_List, Xi, Mask, [D["Train Inputs"] for D in data], [D["Train Targets"] for D in data],
                      [D["Input Bounds"] for D in data], Derivs, LHS, RHS, S["p"], S["Weights"], opt, S["Num Epochs"],
                      S["Num Train Coll Points"], S["Num Test Coll Points"], [D["Test Inputs"] for D in data],
                      [D["Test Targets"] for D in data], Seed=args.seed, On_Epoch=report,
                      Use_CUDA_Graph=bool(args.cuda_graph and S["Optimizer"] == "Adam"))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    with torch.no_grad():                                              # masked components keep their initial value (main.py:419-423)
        for k in range(K):
            if bool(Mask[k]):
                Xi[k] = Initial_Xi[k]
    pde = Format_PDE(LHS, RHS, Xi, Mask)
    if not args.quiet:
        print("Done! It took %7.2fs, an average of %7.4fs per epoch" % (dt, dt / max(S["Num Epochs"], 1)))
        print(pde)
    os.makedirs(saves_dir, exist_ok=True)
    base = "_".join(list(names) + [act, S["Optimizer"]])
    name, c = base, 0
    while os.path.isfile(os.path.join(saves_dir, name)):               # main.py:494-500
        c += 1
        name = base + "_%u" % c
    Save_State(os.path.join(saves_dir, name), U_List, Xi, opt, Derivs, LHS, RHS, list(names))
    if not args.quiet:
        print('Saved as "%s"' % name)
    return {"History": hist, "PDE": pde, "Xi": Xi.detach().cpu(), "Mask": Mask, "Save Name": name, "Seconds": dt}


if __name__ == "__main__":
    main()
