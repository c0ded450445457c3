"""
The reference's three-stage workflow (README.md:90-94: burn-in -> sparsification -> fine-tuning) on the B200 through the Settings.txt
driver (`python -m pde_learn_b200.main`, whole epochs replayed as captured CUDA graphs), on the SAME data set file and the SAME three
Settings files the unmodified reference ran (tests/golden/make_reference_workflow.py -> tests/golden/reference_workflow_burgers.json).
Prints / writes: the PDE identified after each stage, the thresholded support (|xi_k| >= 5e-4, Code/main.py:31,455-460) next to the
reference's, and seconds per epoch next to the reference's.

    python scripts/staged_workflow.py [--out gpurun_out/r02_staged_workflow.json] [--eager]
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_staged_workflow.json"))
    ap.add_argument("--eager", action="store_true", help="per-epoch Training()/Testing() calls instead of captured graphs")
    args = ap.parse_args()
    import torch
    from pde_learn_b200 import configs
    from pde_learn_b200 import main as driver
    ref = json.load(open(os.path.join(GOLD, "reference_workflow_burgers.json")))
    tmp = tempfile.mkdtemp(prefix="pdl_flow_")
    os.makedirs(os.path.join(tmp, "Data", "DataSets")); os.makedirs(os.path.join(tmp, "Saves"))
    shutil.copy(os.path.join(GOLD, "Burgers_Sine_N25_P4000.npz"), os.path.join(tmp, "Data", "DataSets"))
    open(os.path.join(tmp, "Library.txt"), "w").write(configs.SHIPPED_LIBRARY)
    out = {"device": torch.cuda.get_device_name(0), "mode": "eager" if args.eager else "cuda-graph", "stages": []}
    prev = "None"
    for st in ref["stages"]:
        text = ref["settings"][st["stage"]].replace("DEVICE", "gpu").replace("<previous save>", prev)
        open(os.path.join(tmp, "Settings.txt"), "w").write(text)
        argv = ["--settings", os.path.join(tmp, "Settings.txt"), "--quiet", "--seed", str(len(out["stages"]))]
        if not args.eager:
            argv.append("--cuda-graph")
        t0 = time.perf_counter()
        r = driver.main(argv)
        wall = time.perf_counter() - t0
        xi = [float(v) for v in r["Xi"]]
        sup = [k for k, v in enumerate(xi) if abs(v) >= 0.0005 and not bool(r["Mask"][k])]
        rec = {"stage": st["stage"], "epochs": st["epochs"], "pde": r["PDE"], "support": sup, "xi": xi,
               "seconds_per_epoch": r["Seconds"] / st["epochs"], "wall_seconds_incl_setup": wall,
               "reference": {"pde": st["pde"], "support": st["support"], "seconds_per_epoch": st["seconds_per_epoch"],
                             "xi_on_support": [st["xi"][k] for k in st["support"]]},
               "speedup_per_epoch": st["seconds_per_epoch"] / (r["Seconds"] / st["epochs"])}
        out["stages"].append(rec)
        print("%-15s ours: %s   [%.2f ms/epoch]" % (st["stage"], r["PDE"], 1e3 * rec["seconds_per_epoch"]))
        print("%-15s  ref: %s   [%.1f ms/epoch]" % ("", st["pde"], 1e3 * st["seconds_per_epoch"]), flush=True)
        prev = r["Save Name"]
    last = out["stages"][-1]
    out["same_support_as_reference"] = last["support"] == last["reference"]["support"]
    # Burgers_Sine.m:20-21: u_t = -u u_x + 0.1 u_xx  ->  terms 5 = (D_x U)*(U) and 2 = (D_x^2 U) of the shipped Library.txt
    out["true_pde"] = "(D_t U) = -1.0(D_x U)*(U) + 0.1(D_x^2 U)"
    out["recovered_true_support"] = last["support"] == [2, 5]
    print("same support as the reference:", out["same_support_as_reference"], "| true support recovered:", out["recovered_true_support"])
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    shutil.rmtree(tmp, ignore_errors=True)
    return 0 if out["same_support_as_reference"] else 1


if __name__ == "__main__":
    sys.exit(main())
