"""
Generates tests/golden/reference_workflow_burgers.json (+ the data set it ran on, Burgers_Sine_N25_P4000.npz): the reference's
three-stage workflow (README.md:90-94) run END TO END by the reference's own main.py in a scratch copy (SURVEY Appendix D shims,
same recipe as make_reference_save.py):

    burn-in        1000 epochs, Lp weight 0,       all "load" settings false
    sparsification 1000 epochs, Lp weight 0.0002,  load U / Xi+library / optimiser from the burn-in save
    fine-tuning     500 epochs, Lp weight 0,       load everything from the sparsification save, small Xi components masked

on Matlab/Data/Burgers_Sine.mat (u_t = -u u_x + 0.1 u_xx, Burgers_Sine.m:20-21) with 25 % noise and 4000 data points, the shipped
Library.txt (17 terms), 5 x 20 Rational networks, 3000 random collocation points per epoch, Adam 1e-3.  Recorded: Xi after each
stage, the PDE main.py prints, seconds per epoch.  scripts/staged_workflow.py runs the same three Settings files through
`python -m pde_learn_b200.main --cuda-graph` on the B200 and compares the thresholded support.

    python tests/golden/make_reference_workflow.py        (about 10 minutes on 8 cores)
"""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_save as M                       # noqa: E402  (scratch-copy helpers)

STAGES = [("burn_in", 1000, 0.0, False, True), ("sparsification", 1000, 0.0002, True, True), ("fine_tuning", 500, 0.0, True, True)]


def settings_text(epochs, lp, load, load_name):
    b = "True" if load else "False"
    return f"""\
Load U from Save [bool]:                         {b}
Load Xi, Library from Save [bool]:               {b}
Load Optimizer from Save [bool]:                 {b}
    Load File Name [str]:                        {load_name}
Library File [str]:                              Library
Hidden Layer Widths [List of int]:               20, 20, 20, 20, 20
Hidden Activation Function [str]:                Rat
Train on CPU or GPU [GPU, CPU]:                  DEVICE
p [float]:                                       .1
Weights [Dict of float]:                         {{"Data" : 1.0, "Coll" : 1.0, "Lp" : {lp}, "L2" : 0.000005}}
Number of Training Collocation Points [int]:     3000
Number of Testing Collocation Points [int]:      1000
Mask Small Xi Components [bool]:                 True
Optimizer [Adam, LBFGS]:                         Adam
Learning Rate [float]:                           .001
Number of Epochs [int]:                          {epochs}
DataSet Names [List of str]:                     [Burgers_Sine_N25_P4000]
"""


def main():
    tmp = tempfile.mkdtemp(prefix="pdl_refflow_")
    root, env = M.scratch_reference(tmp)
    mk = ("import numpy; numpy.random.seed(0); import From_MATLAB as F; F.Make_Plot = False; "
          "F.From_MATLAB_1D('Burgers_Sine', 0.25, 4000, 1000)")
    subprocess.run([sys.executable, "-c", mk], cwd=os.path.join(root, "Data"), env=env, check=True)
    shutil.copy(os.path.join(root, "Data", "DataSets", "Burgers_Sine_N25_P4000.npz"), os.path.join(HERE, "Burgers_Sine_N25_P4000.npz"))
    out = {"stages": [], "threads": 8}
    load_name = "None"
    for si, (name, epochs, lp, load, _m) in enumerate(STAGES):
        open(os.path.join(root, "Settings.txt"), "w").write(settings_text(epochs, lp, load, load_name).replace("DEVICE", "cpu"))
        run = ("import random, numpy, torch, runpy, sys; random.seed(%d); numpy.random.seed(%d); torch.manual_seed(%d); "
               "torch.serialization.add_safe_globals([]); _l = torch.load; torch.load = lambda *a, **k: _l(*a, **dict(k, weights_only=False)); "
               "sys.argv = ['main.py']; runpy.run_path('main.py', run_name='__main__')" % (si, si, si))
        t0 = time.perf_counter()
        r = subprocess.run([sys.executable, "-W", "ignore", "-c", run], cwd=os.path.join(root, "Code"), env=env, capture_output=True, text=True)
        dt = time.perf_counter() - t0
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-2000:] + r.stderr[-4000:])
            raise SystemExit("stage %s failed" % name)
        save = re.search(r'Saved as "([^"]+)"', r.stdout).group(1)
        per = re.search(r"an average of\s+([0-9.]+)s per epoch", r.stdout)
        pde = [ln for ln in r.stdout.splitlines() if ln.startswith("(D_t U) =")][-1]
        masked = re.search(r"Masking (\d+) RHS terms", r.stdout)
        probe = ("import torch, json; b = torch.load('../Saves/%s', weights_only=False); "
                 "print('XI ' + json.dumps(b['Xi'].detach().double().tolist()))" % save)
        pr = subprocess.run([sys.executable, "-c", probe], cwd=os.path.join(root, "Code"), env=env, capture_output=True, text=True, check=True)
        xi = json.loads([ln for ln in pr.stdout.splitlines() if ln.startswith("XI ")][0][3:])
        out["stages"].append({"stage": name, "epochs": epochs, "lp_weight": lp, "save": save, "seconds_per_epoch": float(per.group(1)),
                              "wall_seconds": dt, "pde": pde, "xi": xi, "masked_at_start": int(masked.group(1)) if masked else 0,
                              "support": [k for k, v in enumerate(xi) if abs(v) >= 0.0005]})
        print(name, "%.3f s/epoch" % float(per.group(1)), pde, flush=True)
        load_name = save
    out["settings"] = {n: settings_text(e, lp, ld, "<previous save>") for n, e, lp, ld, _ in STAGES}
    json.dump(out, open(os.path.join(HERE, "reference_workflow_burgers.json"), "w"), indent=1)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
