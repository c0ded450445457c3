"""
Epoch-level GPU tests: the reference's own seeded training run on real Burgers data (tests/golden/burgers_run.npz,
made by tests/golden/make_burgers_run.py with the unmodified reference on the CPU) replayed through pde_learn_b200,
the targeted-point update, and the epoch driver.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import pde_oracle as O
from tests.helpers import GOLDEN_DIR

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


def test_burgers_run_recovers_same_xi_and_support():
    """240 Adam epochs (random + targeted collocation points) from identical weights/data/points: same Xi trajectory
    within tolerance and the same thresholded support as the reference run (SURVEY 8d parity gate)."""
    import pde_learn_b200 as P
    from pde_learn_b200 import configs
    z = np.load(os.path.join(GOLDEN_DIR, "burgers_run.npz"))
    W = json.loads(str(z["weights"]))
    d, _, derivs, lhs, rhs = configs.config("burgers")                  # == Read_Library(Library.txt), tested on the CPU
    U = P.Network([int(w) for w in z["widths"]], "Tanh", Device=_dev())
    with torch.no_grad():
        for j, prm in enumerate(U.parameters()):
            prm.copy_(torch.from_numpy(z["p0_%02d" % j].copy()).to(_dev()))
    K = len(rhs)
    Xi = torch.zeros(K, dtype=torch.float32, device=_dev(), requires_grad=True)
    Mask = torch.zeros(K, dtype=torch.bool)
    opt = torch.optim.Adam(list(U.parameters()) + [Xi], lr=float(z["lr"]))
    inputs = torch.from_numpy(z["inputs"].copy()).to(_dev()); targets = torch.from_numpy(z["targets"].copy()).to(_dev())
    bounds = z["bounds"]; n_coll = int(z["n_coll"]); every = int(z["check_every"])

    def sampler(i, e, n):                                              # the fixture's seeded CPU generator
        g = torch.Generator().manual_seed(1000 + e)
        u = torch.rand(n, 2, generator=g, dtype=torch.float32)
        lo = torch.tensor(bounds[:, 0], dtype=torch.float32); hi = torch.tensor(bounds[:, 1], dtype=torch.float32)
        return lo + (hi - lo) * u

    xi_hist = []
    hist = P.Run_Epochs([U], Xi, Mask, [inputs], [targets], [bounds], derivs, lhs, rhs, float(z["p"]), W, opt,
                        Num_Epochs=int(z["epochs"]), Num_Train_Coll_Points=n_coll, Point_Sampler=sampler,
                        On_Epoch=lambda e, h: xi_hist.append(Xi.detach().cpu().numpy().copy()) if (e + 1) % every == 0 else None)
    ref_xi = z["xi_hist"]
    dev = [float(np.max(np.abs(a - b))) for a, b in zip(xi_hist, ref_xi)]
    print("max |xi - xi_ref| per checkpoint:", dev, " |xi_ref|max", float(np.abs(ref_xi[-1]).max()))
    assert len(xi_hist) == ref_xi.shape[0]
    # measured on B200: <= 7e-8 at every checkpoint (|xi| up to 0.10); gate at 1e-5 absolute
    assert max(dev) <= 1e-5
    sup = lambda x: [k for k in range(K) if abs(x[k]) >= 5e-4]
    assert sup(xi_hist[-1]) == sup(ref_xi[-1])
    assert P.Threshold_Xi(Xi, Mask) == sup(ref_xi[-1])
    # loss histories follow the reference's (same points every epoch, so they are comparable one by one)
    assert np.allclose(hist["Train Data"][:40, 0], z["data_hist"][:40], rtol=2e-3)
    assert np.allclose(hist["Train Coll"][:40, 0], z["coll_hist"][:40], rtol=2e-2, atol=1e-6)
    assert np.abs(hist["Num Targeted"][:20, 0] - z["n_targeted"][:20]).max() <= 3
    assert abs(hist["Train Data"][-1, 0] - z["data_hist"][-1]) <= 0.05 * z["data_hist"][-1]


def test_update_targeted_points_matches_reference_logic():
    import pde_learn_b200 as P
    torch.manual_seed(4)
    n, n_rand = 50_000, 42_000
    pts = torch.rand(n, 3, device=_dev()); r = torch.randn(n, device=_dev()); r[5] = 20.0; r[n - 2] = -30.0
    got = P.Update_Targeted_Points(pts, r, n_rand)
    cutoff, want = O.targeted_cutoff_select(pts.cpu(), r.cpu(), n_rand)
    assert got.shape == want.shape and torch.equal(got.cpu(), want)
    none = P.Update_Targeted_Points(pts[:100], torch.ones(100, device=_dev()), 100)    # std = 0: every row is >= cutoff
    assert none.shape[0] == 100


def test_run_epochs_two_networks_with_testing_and_device_sampler():
    import pde_learn_b200 as P
    from pde_learn_b200 import configs
    torch.manual_seed(0)
    d, bounds, derivs, lhs, rhs = configs.config("heat")
    U_List = [P.Network([2, 16, 16, 1], "Rational", Device=_dev()) for _ in range(2)]
    K = len(rhs)
    Xi = torch.zeros(K, device=_dev(), requires_grad=True)
    Mask = torch.zeros(K, dtype=torch.bool); Mask[1] = True
    X = [P.Generate_Points(bounds, 300, _dev(), Seed=s) for s in (1, 2)]
    Y = [torch.sin(x[:, 0]).double() * torch.cos(x[:, 1]).double() for x in X]
    opt = torch.optim.Adam([q for U in U_List for q in U.parameters()] + [Xi], lr=2e-3)
    W = {"Data": 1.0, "Coll": 1.0, "Lp": 1e-4, "L2": 1e-6}
    h = P.Run_Epochs(U_List, Xi, Mask, X, Y, [bounds, bounds], derivs, lhs, rhs, 0.1, W, opt, Num_Epochs=30,
                     Num_Train_Coll_Points=500, Num_Test_Coll_Points=200, Seed=3)
    assert h["Train Data"].shape == (30, 2) and np.isfinite(h["Train Total"]).all() and np.isfinite(h["Test Coll"]).all()
    assert h["Train Data"][-1].sum() < h["Train Data"][0].sum()          # it trains
    assert float(Xi.detach()[1]) == 0.0                                            # masked coefficient never moves
    assert (h["Num Targeted"] >= 0).all()


SETTINGS = """
Load U from Save [bool]:                         {load}
Load Xi, Library from Save [bool]:               {load}
Load Optimizer from Save [bool]:                 {load}
    Load File Name [str]:                        {fname}
Library File [str]:                              Library
Hidden Layer Widths [List of int]:               20, 20, 20, 20, 20
Hidden Activation Function [str]:                Tanh
Train on CPU or GPU [GPU, CPU]:                  GPU
p [float]:                                       .1
Weights [Dict of float]:                         {{"Data" : 1.0, "Coll" : 1.0, "Lp" : {lp}, "L2" : 0.000005}}
Number of Training Collocation Points [int]:     3000
Number of Testing Collocation Points [int]:      1000
Mask Small Xi Components [bool]:                 True
Optimizer [Adam, LBFGS]:                         Adam
Learning Rate [float]:                           .001
Number of Epochs [int]:                          {epochs}
DataSet Names [List of str]:                     [Burgers_Sine_N25_P2000]
"""


def test_main_driver_burn_in_then_resume(tmp_path):
    """The reference's 2-stage workflow through `python -m pde_learn_b200.main`: burn-in run, save, resume with masking."""
    import pde_learn_b200 as P
    from pde_learn_b200 import configs
    from pde_learn_b200.main import main
    z = np.load(os.path.join(GOLDEN_DIR, "burgers_run.npz"))
    (tmp_path / "Data" / "DataSets").mkdir(parents=True)
    np.savez(tmp_path / "Data" / "DataSets" / "Burgers_Sine_N25_P2000.npz", Train_Inputs=z["inputs"], Train_Targets=z["targets"],
             Test_Inputs=z["inputs"][:500], Test_Targets=z["targets"][:500], Input_Bounds=z["bounds"])
    (tmp_path / "Library.txt").write_text("# shipped library\n" + configs.SHIPPED_LIBRARY)
    st = tmp_path / "Settings.txt"
    st.write_text(SETTINGS.format(load="False", fname="", lp="0.000", epochs=60))
    out1 = main(["--settings", str(st), "--quiet"])
    assert out1["History"]["Train Data"][-1, 0] < out1["History"]["Train Data"][0, 0]
    assert out1["PDE"].startswith("(D_t U) = ") and os.path.isfile(tmp_path / "Saves" / out1["Save Name"])
    xi1 = out1["Xi"].clone()
    # stage 2: resume everything, mask the small coefficients, switch the Lp term on
    st.write_text(SETTINGS.format(load="True", fname=out1["Save Name"], lp="0.0002", epochs=20))
    out2 = main(["--settings", str(st), "--quiet"])
    mask = out2["Mask"]
    assert mask.tolist() == (xi1.abs() < 5e-4).tolist()
    for k in range(len(mask)):
        if mask[k]:
            assert float(out2["Xi"][k]) == float(xi1[k])             # masked coefficients are restored, never trained
    assert out2["Save Name"] != out1["Save Name"] and np.isfinite(out2["History"]["Train Total"]).all()
    # the save file can be read back (and has the reference's keys)
    blob = P.Load_State(str(tmp_path / "Saves" / out2["Save Name"]), torch.device("cuda", 0))
    assert torch.equal(blob["Xi"].detach().cpu(), out2["Xi"]) and len(blob["RHS Terms"]) == 17


def test_resume_reference_save_on_gpu():
    """(f3) A save written by the reference's main.py (tests/golden/make_reference_save.py) is resumed on the B200: network, Xi and
    Adam moments restored by Load_State, then three Training epochs on the fixture's points.  The reference resumed the same file on
    the CPU for the same three epochs (Code/main.py:51-57,140-160,232-246 + Test_Train.py:22); Xi, the parameters and the losses
    must agree."""
    import pde_learn_b200 as P
    z = np.load(os.path.join(GOLDEN_DIR, "reference_save_resume.npz"))
    st = P.Load_State(os.path.join(GOLDEN_DIR, "reference_save_burgers.pt"), _dev())
    U, Xi = st["U List"][0], st["Xi"]
    assert Xi.is_cuda and all(q.is_cuda for q in U.parameters())
    opt = torch.optim.Adam(list(U.parameters()) + [Xi], lr=1e-3)
    opt.load_state_dict(st["Optimizer State"])
    Mask = torch.from_numpy(z["mask"].copy())
    W = {"Data": 1.0, "Coll": 1.0, "Lp": 0.0002, "L2": 0.000005}
    inputs = torch.from_numpy(z["inputs"].copy()).to(_dev()); targets = torch.from_numpy(z["targets"].copy()).to(_dev())
    rel = lambda a, b: abs(a - b) / max(abs(b), 1e-30)
    for e in range(z["points"].shape[0]):
        pts = torch.from_numpy(z["points"][e].copy()).to(_dev())
        out = P.Training([U], Xi, Mask, [pts], [inputs], [targets], st["Derivatives"], st["LHS Term"], st["RHS Terms"], 0.1, W, opt)
        assert rel(out["Coll Losses"][0], float(z["coll"][e])) <= 1e-5 and rel(out["Data Losses"][0], float(z["data"][e])) <= 1e-5
        assert rel(out["Total Losses"][0], float(z["total"][e])) <= 1e-5
        dxi = float(np.abs(Xi.detach().cpu().numpy() - z["xi"][e]).max())
        assert dxi <= 2e-6, (e, dxi)                       # Adam step = 1e-3; identical moments -> identical steps up to fp32 gradient noise
    flat = torch.cat([q.detach().reshape(-1) for q in U.parameters()]).cpu().numpy()
    assert float(np.abs(flat - z["params_after"]).max()) <= 2e-5
    assert float(opt.state_dict()["state"][0]["step"]) == 12 + 3


@pytest.mark.parametrize("optimizer,graph", [("LBFGS", False), ("Adam", True)])
def test_settings_driver_end_to_end(optimizer, graph, tmp_path):
    """`python -m pde_learn_b200.main` (the Settings.txt workflow of Code/main.py:35-539) on the data-set file the reference's
    From_MATLAB wrote (tests/golden/Burgers_Sine_N25_P4000.npz): a short burn-in with the given optimiser, then a second run that
    loads U, Xi + library and the optimiser state from the first run's save file (README.md:90-94) -- finite decreasing losses, a
    save file in the reference's format after each run, the resumed run continues the optimiser's step count."""
    import shutil
    import pde_learn_b200 as P
    from pde_learn_b200 import configs
    from pde_learn_b200 import main as driver
    root = tmp_path
    os.makedirs(root / "Data" / "DataSets"); os.makedirs(root / "Saves")
    shutil.copy(os.path.join(GOLDEN_DIR, "Burgers_Sine_N25_P4000.npz"), root / "Data" / "DataSets")
    (root / "Library.txt").write_text(configs.SHIPPED_LIBRARY)
    ref = json.load(open(os.path.join(GOLDEN_DIR, "reference_workflow_burgers.json")))
    epochs = 6 if optimizer == "LBFGS" else 40
    lr = "0.05" if optimizer == "LBFGS" else ".001"

    def settings(load, name):
        t = ref["settings"]["burn_in"].replace("DEVICE", "gpu").replace("<previous save>", name)
        t = t.replace("Number of Epochs [int]:                          1000", "Number of Epochs [int]:                          %d" % epochs)
        t = t.replace("Optimizer [Adam, LBFGS]:                         Adam", "Optimizer [Adam, LBFGS]:                         " + optimizer)
        t = t.replace("Learning Rate [float]:                           .001", "Learning Rate [float]:                           " + lr)
        if load:
            t = "\n".join(ln.replace("False", "True") if ln.startswith("Load ") else ln for ln in t.splitlines()) + "\n"
        return t

    (root / "Settings.txt").write_text(settings(False, "None"))
    argv = ["--settings", str(root / "Settings.txt"), "--quiet"] + (["--cuda-graph"] if graph else [])
    r1 = driver.main(argv)
    h = r1["History"]
    assert np.isfinite(h["Train Total"]).all() and np.isfinite(h["Test Data"]).all()
    assert h["Train Data"][-1, 0] < h["Train Data"][0, 0]                        # the data loss goes down during burn-in
    st = P.Load_State(str(root / "Saves" / r1["Save Name"]), _dev())
    assert st["DataSet Names"] == ["Burgers_Sine_N25_P4000"] and len(st["RHS Terms"]) == 17
    text = settings(True, r1["Save Name"])
    assert text.count("True") >= 4                                               # the three load settings + the mask setting
    (root / "Settings.txt").write_text(text)
    r2 = driver.main(argv)
    assert r2["Save Name"] != r1["Save Name"] and np.isfinite(r2["History"]["Train Total"]).all()
    assert r2["History"]["Train Data"][0, 0] < h["Train Data"][0, 0]              # resumed from the trained state, not from scratch
    if optimizer == "Adam":
        blob = torch.load(str(root / "Saves" / r2["Save Name"]), weights_only=False)
        assert float(blob["Optimizer"]["state"][0]["step"]) == 2 * epochs


def test_three_stage_workflow_recovers_burgers_pde():
    """The deliverable end to end (README.md:90-94 of the reference): burn-in -> sparsification -> fine-tuning through the Settings.txt
    driver with captured epochs, on the data-set file and the three Settings files the unmodified reference's main.py ran
    (tests/golden/make_reference_workflow.py).  Same thresholded support as the reference's own run = the true PDE
    u_t = -u u_x + 0.1 u_xx (Burgers_Sine.m:20-21), coefficients within 15 % of the truth (the reference: -0.963, 0.1022)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out", "test_staged_workflow.json")
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "staged_workflow.py"), "--out", out], capture_output=True, text=True,
                       timeout=900, cwd=root)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    res = json.load(open(out))
    assert res["same_support_as_reference"] and res["recovered_true_support"]
    last = res["stages"][-1]
    assert last["support"] == [2, 5]
    assert abs(last["xi"][5] + 1.0) <= 0.15 and abs(last["xi"][2] - 0.1) <= 0.015
    assert last["seconds_per_epoch"] < 0.01                                     # the reference: 0.24 s per epoch on 8 host cores
