"""
Generates tests/golden/reference_save_burgers.pt (+ reference_save_burgers.json, reference_save_resume.npz): a save file WRITTEN BY THE REFERENCE'S OWN
`Code/main.py` (main.py:487-524), produced by running the reference end to end in a scratch copy as SURVEY Appendix D describes:

    1. copy /root/reference to a temporary directory (main.py writes ../Saves, From_MATLAB writes Data/DataSets);
    2. stub matplotlib / seaborn modules on PYTHONPATH (absent in this image; plotting is out of scope);
    3. `ln -s Matlab MATLAB` (From_MATLAB.py reads ../MATLAB/Data/, the directory on disk is Matlab/);
    4. Points.py:57 in the SCRATCH copy: float(...) around the numpy-2 float32 scalars;
    5. a Settings.txt for a short run (Rational, Adam, 12 epochs, 5 x 20) on Burgers_Sine with 75 % noise, 2000 data points.

The committed fixture is what `tests/test_host_logic.py::test_reference_written_save_*` loads with `Driver.Load_State`, and
what `tests/test_gpu_epochs.py::test_resume_reference_save_on_gpu` resumes on the B200.  The json next to it records the numbers
main.py printed/holds at save time (Xi, a forward value of U at fixed inputs) so the loader is checked against the reference's
own state, not against itself; the npz holds three further epochs the REFERENCE trained after loading its own save (network,
Xi and Adam moments restored, one term masked) on seeded points -- the numbers a resumed run here must reproduce.

    python tests/golden/make_reference_save.py
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import textwrap

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

SETTINGS = """\
Load U from Save [bool]:                         False
Load Xi, Library from Save [bool]:               False
Load Optimizer from Save [bool]:                 False
    Load File Name [str]:                        None
Library File [str]:                              Library
Hidden Layer Widths [List of int]:               20, 20, 20, 20, 20
Hidden Activation Function [str]:                Rat
Train on CPU or GPU [GPU, CPU]:                  cpu
p [float]:                                       .1
Weights [Dict of float]:                         {"Data" : 1.0, "Coll" : 1.0, "Lp" : 0.0002, "L2" : 0.000005}
Number of Training Collocation Points [int]:     600
Number of Testing Collocation Points [int]:      200
Mask Small Xi Components [bool]:                 False
Optimizer [Adam, LBFGS]:                         Adam
Learning Rate [float]:                           .001
Number of Epochs [int]:                          12
DataSet Names [List of str]:                     [Burgers_Sine_N75_P2000]
"""

PROBE = """\
import sys, json, numpy, torch
sys.path[:0] = ["Classes", "Readers", "."]
from Network import Network
blob = torch.load("../Saves/%s", weights_only=False)
st = blob["U States"][0]
U = Network(Widths=st["Widths"], Hidden_Activation=st["Activation Types"][0], Output_Activation=st["Activation Types"][-1])
U.Set_State(st)
X = torch.tensor([[0.5, -3.0], [2.0, 0.25], [7.5, 6.0], [9.0, -7.5]], dtype=torch.float32)
out = {"xi": blob["Xi"].detach().double().tolist(), "probe_inputs": X.tolist(), "probe_values": U(X).detach().double().view(-1).tolist(),
       "widths": list(st["Widths"]), "activation_types": st["Activation Types"], "dataset_names": blob["DataSet Names"],
       "derivative_encodings": [numpy.asarray(e).tolist() for e in blob["Derivative Encodings"]],
       "num_rhs_terms": len(blob["RHS Term States"]), "adam_step": float(blob["Optimizer"]["state"][0]["step"]),
       "num_param_tensors": len(blob["Optimizer"]["state"])}
print("PROBE " + json.dumps(out))
# ---- resume: the reference loads its own save (main.py:51-57,140-160,232-246) and trains 3 more epochs on seeded points ----
from Library_Reader import Read_Library
from Test_Train import Training
from Term import Build_Term_From_State
from Derivative import Derivative
Derivs = [Derivative(Encoding = e) for e in blob["Derivative Encodings"]]
LHS = Build_Term_From_State(blob["LHS Term State"]); RHS = [Build_Term_From_State(s) for s in blob["RHS Term States"]]
Xi = blob["Xi"].detach().clone().requires_grad_(True)
Opt = torch.optim.Adam(list(U.parameters()) + [Xi], lr = 0.001)
Opt.load_state_dict(blob["Optimizer"])
D = numpy.load("../Data/DataSets/Burgers_Sine_N75_P2000.npz")
inputs = torch.from_numpy(D["Train_Inputs"]); targets = torch.from_numpy(D["Train_Targets"])
bounds = D["Input_Bounds"]
Mask = torch.zeros(len(RHS), dtype = torch.bool); Mask[3] = True
W = {"Data": 1.0, "Coll": 1.0, "Lp": 0.0002, "L2": 0.000005}
g = torch.Generator().manual_seed(77)
lo = torch.tensor(bounds[:, 0], dtype = torch.float32); hi = torch.tensor(bounds[:, 1], dtype = torch.float32)
pts, xis, coll, data, tot = [], [], [], [], []
for e in range(3):
    P = lo + (hi - lo) * torch.rand(600, 2, generator = g, dtype = torch.float32)
    pts.append(P.numpy().copy())
    o = Training([U], Xi, Mask, [P], [inputs], [targets], Derivs, LHS, RHS, 0.1, W, Opt)
    xis.append(Xi.detach().numpy().copy()); coll.append(o["Coll Losses"][0]); data.append(o["Data Losses"][0]); tot.append(o["Total Losses"][0])
numpy.savez("../resume.npz", inputs = inputs.numpy(), targets = targets.numpy(), bounds = bounds, points = numpy.stack(pts),
            xi = numpy.stack(xis), coll = numpy.array(coll), data = numpy.array(data), total = numpy.array(tot), mask = Mask.numpy(),
            params_after = numpy.concatenate([q.detach().numpy().reshape(-1) for q in U.parameters()]))
"""


def scratch_reference(tmp):
    """Scratch copy of the reference with the Appendix D shims; returns (root, env)."""
    root = os.path.join(tmp, "ref")
    shutil.copytree(REF, root, symlinks=True)
    os.symlink("Matlab", os.path.join(root, "MATLAB"))
    os.makedirs(os.path.join(root, "Saves"), exist_ok=True)
    os.makedirs(os.path.join(root, "Data", "DataSets"), exist_ok=True)
    stubs = os.path.join(tmp, "stubs")
    os.makedirs(os.path.join(stubs, "matplotlib"))
    open(os.path.join(stubs, "matplotlib", "__init__.py"), "w").write("")
    open(os.path.join(stubs, "matplotlib", "pyplot.py"), "w").write(textwrap.dedent("""\
        def __getattr__(name):
            def f(*a, **k):
                return None
            return f
        """))
    open(os.path.join(stubs, "seaborn.py"), "w").write("def __getattr__(name):\n    return lambda *a, **k: None\n")
    # Appendix D item 4, scratch copy only
    pts = os.path.join(root, "Code", "Points.py")
    src = open(pts).read()
    fixed = src.replace("random.uniform(Lower_Bound, Upper_Bound)", "random.uniform(float(Lower_Bound), float(Upper_Bound))")
    assert fixed != src, "Points.py:57 not found"
    open(pts, "w").write(fixed)
    return root, dict(os.environ, PYTHONPATH=stubs, OMP_NUM_THREADS="8")


def main():
    tmp = tempfile.mkdtemp(prefix="pdl_refsave_")
    root, env = scratch_reference(tmp)
    open(os.path.join(root, "Settings.txt"), "w").write(SETTINGS)
    # data set (Data/From_MATLAB.py, Make_Plot off), then the reference's main.py
    mk = ("import numpy, random, torch; numpy.random.seed(0); import From_MATLAB as F; F.Make_Plot = False; "
          "F.From_MATLAB_1D('Burgers_Sine', 0.75, 2000, 500)")
    subprocess.run([sys.executable, "-c", mk], cwd=os.path.join(root, "Data"), env=env, check=True)
    run = ("import random, numpy, torch, runpy, sys; random.seed(0); numpy.random.seed(0); torch.manual_seed(0); "
           "sys.argv = ['main.py']; runpy.run_path('main.py', run_name='__main__')")
    r = subprocess.run([sys.executable, "-c", run], cwd=os.path.join(root, "Code"), env=env, capture_output=True, text=True)
    sys.stdout.write(r.stdout[-2500:])
    if r.returncode != 0:
        sys.stderr.write(r.stderr[-4000:])
        raise SystemExit("the reference's main.py failed")
    saves = sorted(f for f in os.listdir(os.path.join(root, "Saves")) if not f.startswith("."))
    assert len(saves) == 1, saves
    pr = subprocess.run([sys.executable, "-c", PROBE % saves[0]], cwd=os.path.join(root, "Code"), env=env, capture_output=True, text=True,
                        check=True)
    line = [ln for ln in pr.stdout.splitlines() if ln.startswith("PROBE ")][0][6:]
    meta = json.loads(line)
    meta["save_name"] = saves[0]
    meta["written_by"] = "Code/main.py of the unmodified reference (scratch copy with the Appendix D shims), 12 Adam epochs, seeds 0"
    shutil.copy(os.path.join(root, "Saves", saves[0]), os.path.join(HERE, "reference_save_burgers.pt"))
    shutil.copy(os.path.join(root, "resume.npz"), os.path.join(HERE, "reference_save_resume.npz"))
    json.dump(meta, open(os.path.join(HERE, "reference_save_burgers.json"), "w"), indent=1)
    print("wrote reference_save_burgers.pt (%d bytes) from %s" % (os.path.getsize(os.path.join(HERE, "reference_save_burgers.pt")), saves[0]))
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
