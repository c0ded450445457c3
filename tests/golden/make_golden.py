"""
Generates tests/golden/*.npz by importing the UNMODIFIED reference from /root/reference (read-only) and
running its own float64 autograd path.  Run once in the build container (the reference does not exist on
the GPU box); the outputs are committed.

    python tests/golden/make_golden.py

Every case stores the inputs (float32 parameters, points, xi, mask, library description) and the
reference's float64 outputs (derivative features, residual, losses, gradients of every parameter and
xi).  Reference entry points exercised: Code/Loss.py:66 Coll_Loss, :23 Data_Loss, :252 Lp_Loss,
:316 L2_Squared_Loss, Code/Evaluate_Derivatives.py:20 Derivative_From_Derivative,
Code/Classes/Network.py:63 Network, Code/Test_Train.py:22 Training, :205 Testing.
"""
import copy
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference/Code"
for p in (REF, os.path.join(REF, "Classes"), os.path.join(REF, "Readers")):
    sys.path.insert(0, p)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from Derivative import Derivative                                # noqa: E402  (reference)
from Term import Term                                            # noqa: E402
from Network import Network                                      # noqa: E402
from Loss import Coll_Loss, Data_Loss, Lp_Loss, L2_Squared_Loss  # noqa: E402
from Evaluate_Derivatives import Derivative_From_Derivative      # noqa: E402
from Test_Train import Training, Testing                         # noqa: E402
from Library_Reader import Read_Library                          # noqa: E402

from oracle.pde_oracle import config_library, distinct_encodings  # noqa: E402  (library *definitions* only)


def ref_terms(lhs, rhs):
    """TermSpec lists -> reference Derivative/Term objects + sorted derivative list."""
    mk = lambda t: Term(Derivatives=[Derivative(Encoding=np.array(e)) for e, _ in t],
                        Powers=[int(p) for _, p in t])
    encs = distinct_encodings(lhs, rhs)
    return [Derivative(Encoding=np.array(e)) for e in encs], mk(lhs), [mk(t) for t in rhs], encs


def points_in(bounds, n, gen):
    lo = torch.tensor([b[0] for b in bounds]); hi = torch.tensor([b[1] for b in bounds])
    return (lo + (hi - lo) * torch.rand(n, len(bounds), generator=gen)).to(torch.float32)


def spec_json(lhs, rhs):
    return json.dumps({"lhs": lhs, "rhs": rhs})


def run_case(name, d, lhs, rhs, bounds, hidden, act, n_pts, seed, masked=(), xi_scale=0.1,
             perturb_rational=False, bias_scale=0.0, U_trained=None, xi_trained=None):
    """`U_trained` / `xi_trained` (tests/golden/make_golden_trained.py): a reference network and coefficient vector taken from the end
    of a reference training run instead of the Xavier initialisation."""
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    U = U_trained if U_trained is not None else Network(Widths=[d] + list(hidden) + [1], Hidden_Activation=act, Output_Activation="None")
    if bias_scale > 0:
        with torch.no_grad():
            for L in U.Layers:
                L.bias.copy_(bias_scale * torch.randn(L.bias.shape, generator=gen))
    if perturb_rational:
        with torch.no_grad():
            for af in U.Activation_Functions[:-1]:
                af.a.add_(0.05 * torch.randn(4, generator=gen))
                af.b.add_(0.05 * torch.randn(3, generator=gen)); af.b[0] = af.b[0].abs() + 0.5
    Derivs, LHS, RHS, encs = ref_terms(lhs, rhs)
    K = len(rhs)
    xi32 = (xi_scale * torch.randn(K, generator=gen)).to(torch.float32)
    if xi_trained is not None:
        xi32 = xi_trained.detach().clone().to(torch.float32)
    mask = torch.zeros(K, dtype=torch.bool)
    for k in masked:
        mask[k] = True
    pts32 = points_in(bounds, n_pts, gen)
    n_data = 64
    x_data = points_in(bounds, n_data, gen)
    y_data = torch.randn(n_data, generator=gen, dtype=torch.float64)

    # float64 twin of the reference network.
    U64 = copy.deepcopy(U).double()
    xi64 = xi32.double().requires_grad_(True)
    P = pts32.double()
    loss, resid = Coll_Loss(U=U64, Xi=xi64, Mask=mask, Coll_Points=P, Derivatives=Derivs,
                            LHS_Term=LHS, RHS_Terms=RHS)
    params = list(U64.parameters())
    grads = torch.autograd.grad(loss, params + [xi64], retain_graph=True)
    # Derivative features through the reference's own evaluator.
    P2 = pts32.double().requires_grad_(True)
    u = U64(P2).view(-1)
    I = Derivative(Encoding=np.array([0, 0]))
    feats = [Derivative_From_Derivative(Da=D, Db=I, Db_U=u, Coords=P2).detach().numpy() for D in Derivs]

    # Data / Lp / L2 with the float32 network exactly as the training loop sees them.
    d_loss = Data_Loss(U=U64, Inputs=x_data.double(), Targets=y_data)
    d_grads = torch.autograd.grad(d_loss, params)
    p_exp = 0.1
    lp = Lp_Loss(Xi=xi64, Mask=mask, p=p_exp)
    lp_grad = torch.autograd.grad(lp, [xi64])[0]
    l2 = L2_Squared_Loss(U=U64)

    out = {
        "spec": np.array(spec_json(lhs, rhs)), "d": d, "widths": np.array([d] + list(hidden) + [1]),
        "activation": np.array(act.lower()), "bounds": np.array(bounds, dtype=np.float64),
        "encodings": np.array([list(e) + [0] * (4 - len(e)) for e in encs], dtype=np.int32),
        "xi": xi32.numpy(), "mask": mask.numpy(), "points": pts32.numpy(),
        "x_data": x_data.numpy(), "y_data": y_data.numpy(), "p": p_exp,
        "coll_loss": float(loss), "residual": resid.detach().numpy(), "features": np.stack(feats),
        "grad_xi": grads[-1].numpy(), "data_loss": float(d_loss), "lp_loss": float(lp),
        "lp_grad": lp_grad.numpy(), "l2_loss": float(l2),
    }
    for i, prm in enumerate(U.parameters()):
        out["param_%02d" % i] = prm.detach().numpy()
        out["grad_%02d" % i] = grads[i].numpy()
        out["dgrad_%02d" % i] = d_grads[i].numpy()
    out["n_params"] = len(params)
    if U_trained is not None:
        # conditioning of the case: the error of the reference's OWN float32 path (same functions, float32 network and points)
        # against its float64 path, as relative max-norms [loss, residual, dtheta, dxi, features]
        rel = lambda a, b: float(np.max(np.abs(np.asarray(a, dtype=np.float64) - b)) / max(np.max(np.abs(b)), 1e-300))
        xi_f = xi32.clone().requires_grad_(True)
        l32, r32 = Coll_Loss(U=U, Xi=xi_f, Mask=mask, Coll_Points=pts32.clone(), Derivatives=Derivs, LHS_Term=LHS, RHS_Terms=RHS)
        g32 = torch.autograd.grad(l32, list(U.parameters()) + [xi_f], retain_graph=True)
        P3 = pts32.clone().requires_grad_(True)
        u3 = U(P3).view(-1)
        f32 = [Derivative_From_Derivative(Da=D, Db=I, Db_U=u3, Coords=P3).detach().numpy() for D in Derivs]
        out["ref_fp32_err"] = np.array([
            abs(float(l32) - float(loss)) / float(loss), rel(r32.detach().numpy(), resid.detach().numpy()),
            rel(np.concatenate([x.numpy().ravel() for x in g32[:-1]]), np.concatenate([x.numpy().ravel() for x in grads[:-1]])),
            rel(g32[-1].numpy(), grads[-1].numpy()), max(rel(a, b) for a, b in zip(f32, feats))])
        print("  reference float32 vs float64 [loss, resid, dtheta, dxi, feats]:", ["%.1e" % v for v in out["ref_fp32_err"]])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-22s J=%d K=%d loss=%.6e data=%.6e" % (name, len(encs), K, float(loss), float(d_loss)))


def run_training_case(name, optimizer="Adam"):
    """One reference Training() step with Adam (or LBFGS: one step = up to 20 closure evaluations, Code/main.py:236, README.md:96)
    + one Testing() call, float32, two data sets sharing xi."""
    d, lhs, rhs, bounds = config_library("heat")
    torch.manual_seed(7)
    gen = torch.Generator().manual_seed(8)
    Derivs, LHS, RHS, encs = ref_terms(lhs, rhs)
    K = len(rhs)
    U_List = [Network(Widths=[d, 16, 16, 16, 1], Hidden_Activation="Tanh", Output_Activation="None")
              for _ in range(2)]
    Xi = (0.1 * torch.randn(K, generator=gen)).to(torch.float32).requires_grad_(True)
    Mask = torch.zeros(K, dtype=torch.bool); Mask[3] = True
    Coll = [points_in(bounds, 200, gen) for _ in range(2)]
    Inp = [points_in(bounds, 50, gen) for _ in range(2)]
    Tgt = [torch.randn(50, generator=gen, dtype=torch.float64) for _ in range(2)]
    W = {"Data": 1.0, "Coll": 1.0, "Lp": 2e-4, "L2": 5e-6}
    out = {"spec": np.array(spec_json(lhs, rhs)), "d": d, "widths": np.array([d, 16, 16, 16, 1]),
           "xi0": Xi.detach().numpy().copy(), "mask": Mask.numpy(), "p": 0.1,
           "weights": np.array(json.dumps(W)), "lr": 1e-3}
    for s in range(2):
        out["coll_%d" % s] = Coll[s].numpy(); out["inp_%d" % s] = Inp[s].numpy(); out["tgt_%d" % s] = Tgt[s].numpy()
        for i, prm in enumerate(U_List[s].parameters()):
            out["p0_%d_%02d" % (s, i)] = prm.detach().numpy().copy()
    params = [q for U in U_List for q in U.parameters()] + [Xi]
    Opt = torch.optim.Adam(params, lr=1e-3) if optimizer == "Adam" else torch.optim.LBFGS(params, lr=0.1)
    out["optimizer"] = np.array(optimizer); out["lr"] = 1e-3 if optimizer == "Adam" else 0.1
    test0 = Testing(U_List, Xi, Mask, [c.clone() for c in Coll], Inp, Tgt, Derivs, LHS, RHS, 0.1, W)
    tr = Training(U_List, Xi, Mask, [c.clone() for c in Coll], Inp, Tgt, Derivs, LHS, RHS, 0.1, W, Opt)
    for key in ("Coll Losses", "Data Losses", "L2 Losses", "Total Losses"):
        out["train_" + key.replace(" ", "_")] = np.array(tr[key], dtype=np.float64)
        out["test_" + key.replace(" ", "_")] = np.array(test0[key], dtype=np.float64)
    out["test_Lp_Loss"] = float(test0["Lp Loss"])
    for s in range(2):
        out["resid_%d" % s] = tr["Residuals"][s].numpy()
        for i, prm in enumerate(U_List[s].parameters()):
            out["p1_%d_%02d" % (s, i)] = prm.detach().numpy().copy()
    out["xi1"] = Xi.detach().numpy().copy()
    if optimizer != "Adam":                        # losses after the step (LBFGS returns the FIRST closure's losses)
        test1 = Testing(U_List, Xi, Mask, [c.clone() for c in Coll], Inp, Tgt, Derivs, LHS, RHS, 0.1, W)
        out["after_Total_Losses"] = np.array(test1["Total Losses"], dtype=np.float64)
        out["n_iter"] = int(Opt.state[Opt._params[0]]["n_iter"]); out["func_evals"] = int(Opt.state[Opt._params[0]]["func_evals"])
    out["n_params"] = len(list(U_List[0].parameters()))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-22s train total=%s" % (name, tr["Total Losses"]))


def run_library_reader_case():
    """What the reference's reader makes of the shipped Library.txt (pins our reader + config_library)."""
    Derivs, LHS, RHS = Read_Library("/root/reference/Library.txt")
    st = lambda T: [[[int(v) for v in D.Encoding], int(p)] for D, p in zip(T.Derivatives, T.Powers)]
    blob = {"derivatives": [[int(v) for v in D.Encoding] for D in Derivs], "lhs": st(LHS),
            "rhs": [st(T) for T in RHS]}
    with open(os.path.join(HERE, "library_txt_parsed.json"), "w") as f:
        json.dump(blob, f, indent=1)
    print("library_txt_parsed     J=%d K=%d" % (len(Derivs), len(RHS)))


if __name__ == "__main__":
    torch.set_num_threads(8)
    U, Ut, Ux, Uxx = (0, 0), (1, 0), (0, 1), (0, 2)
    d, lhs, rhs, b = config_library("burgers")
    run_case("burgers_tanh", d, lhs, rhs, b, [40] * 5, "Tanh", 192, 11)
    run_case("burgers_tanh_masked", d, lhs, rhs, b, [40] * 5, "Tanh", 100, 12, masked=(1, 4, 16), bias_scale=0.1)
    run_case("burgers_rational20", d, lhs, rhs, b, [20] * 5, "Rational", 128, 13, perturb_rational=True)
    d, lhs, rhs, b = config_library("heat")
    run_case("heat_tanh", d, lhs, rhs, b, [40] * 5, "Tanh", 160, 14, bias_scale=0.05)
    run_case("heat_rational", d, lhs, rhs, b, [40] * 5, "Rational", 96, 15)
    d, lhs, rhs, b = config_library("kdv")
    run_case("kdv_tanh", d, lhs, rhs, b, [40] * 5, "Tanh", 96, 16)
    d, lhs, rhs, b = config_library("ks")
    run_case("ks_tanh", d, lhs, rhs, b, [40] * 5, "Tanh", 96, 17)
    run_case("ks_rational", d, lhs, rhs, b, [40] * 5, "Rational", 64, 18, perturb_rational=True)
    d, lhs, rhs, b = config_library("rd2d")
    run_case("rd2d_tanh", d, lhs, rhs, b, [40] * 5, "Tanh", 96, 19)
    # Wave-type library with second-order time and mixed derivatives, ragged widths, 3 hidden layers.
    Utt, Utx = (2, 0), (1, 1)
    run_case("wave_mixed_tanh", 2, [(Utt, 1)], [[(Uxx, 1)], [(Utx, 1)], [(U, 1), (Ux, 1)], [(Ut, 2)], [(U, 3)]],
             [(0.0, 2.0), (-1.0, 1.0)], [24, 17, 30], "Tanh", 77, 20, bias_scale=0.1)
    # 3 spatial dimensions (d = 4), mixed y-z derivative.
    run_case("heat3d_tanh", 4, [((1, 0), 1)],
             [[((0, 2), 1)], [((0, 0, 2), 1)], [((0, 0, 0, 2), 1)], [((0, 0, 1, 1), 1)], [((0, 0), 2)]],
             [(0.0, 1.0), (-1.0, 1.0), (-1.0, 1.0), (-1.0, 1.0)], [16, 16], "Tanh", 50, 21)
    run_training_case("training_step_heat")
    run_training_case("training_lbfgs_heat", optimizer="LBFGS")
    run_library_reader_case()


def run_settings_reader_case():
    """What the reference's Settings_Reader makes of the shipped Settings.txt (needs cwd = reference/Code)."""
    import subprocess
    code = ("import sys, json; sys.path.insert(0, 'Readers'); sys.path.insert(0, 'Classes'); "
            "from Settings_Reader import Settings_Reader; S = Settings_Reader(); "
            "S = {k: (str(v) if k == 'Device' else v) for k, v in S.items()}; "
            "json.dump(S, open('%s', 'w'), indent=1)" % os.path.join(HERE, "settings_txt_parsed.json"))
    subprocess.run([sys.executable, "-c", code], cwd=REF, check=True)


if __name__ == "__main__":
    run_settings_reader_case()
