"""Host-side logic that needs no GPU: symbols, Library.txt reader, configs, network introspection, gloo all-reduce."""
import json
import os
import tempfile

import numpy as np
import pytest
import torch

import pde_learn_b200 as P
from pde_learn_b200 import configs, engine
from pde_learn_b200.Library_Reader import Parse_Term, Read_Library
from tests.helpers import GOLDEN_DIR

SHIPPED = configs.SHIPPED_LIBRARY


def test_read_library_matches_reference_reader():
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write("# comment line\n\n" + SHIPPED.replace("(U)^2\n", "(U)^2   # trailing comment\n", 1))
        path = f.name
    derivs, lhs, rhs = Read_Library(path)
    os.unlink(path)
    blob = json.load(open(os.path.join(GOLDEN_DIR, "library_txt_parsed.json")))
    st = lambda T: [[[int(v) for v in D.Encoding], int(p)] for D, p in zip(T.Derivatives, T.Powers)]
    assert [[int(v) for v in D.Encoding] for D in derivs] == blob["derivatives"]
    assert st(lhs) == blob["lhs"] and [st(t) for t in rhs] == blob["rhs"]


def test_parse_term_grammar():
    T = Parse_Term("(D_x^2 D_y U)^3*(D_t U)*U^2")
    assert [D.Encoding.tolist() for D in T.Derivatives] == [[0, 2, 1], [1, 0], [0, 0]]
    assert T.Powers == [3, 1, 2] and T.Num_Sub_Terms == 3
    assert str(P.Derivative(np.array([1, 2, 0, 1]))) == "D_t D_x^2 D_z "
    with pytest.raises(ValueError):
        Parse_Term("(D_q U)")


def test_derivative_and_term_semantics():
    a, b = P.Derivative(np.array([0, 1])), P.Derivative(np.array([0, 2]))
    assert a.Order == 1 and a.Is_Child_Of(b) and not b.Is_Child_Of(a)
    with pytest.raises(AssertionError):
        P.Derivative(np.array([1]))
    with pytest.raises(AssertionError):
        P.Term([a], [0])
    t = P.Term([a, b], [1, 2])
    t2 = P.Build_Term_From_State(t.Get_State())
    assert str(t2) == str(t) == "(D_x U)*(D_x^2 U)^2"


def test_config_sizes_match_survey():
    for name, J, K, d in [("burgers", 5, 17, 2), ("heat", 4, 9, 2), ("kdv", 5, 14, 2), ("ks", 6, 30, 2), ("rd2d", 6, 21, 3)]:
        dd, bounds, derivs, lhs, rhs = configs.config(name)
        assert (dd, len(derivs), len(rhs), bounds.shape) == (d, J, K, (d, 2))


def test_network_layout_and_signature():
    torch.manual_seed(0)
    U = P.Network([2, 40, 40, 40, 40, 40, 1], "Rational")
    names = [n for n, _ in U.named_parameters()]
    assert names[:4] == ["Layers.0.weight", "Layers.0.bias", "Layers.1.weight", "Layers.1.bias"]
    assert names[-2:] == ["Activation_Functions.4.a", "Activation_Functions.4.b"]
    assert sum(p.numel() for p in U.parameters()) == 6721 + 35          # SURVEY "Parameter order"
    assert [tuple(p.shape) for p in engine.network_param_tensors(U)] == [tuple(p.shape) for p in U.parameters()]
    assert engine.network_signature(U) == ((2, 40, 40, 40, 40, 40, 1), 1)
    assert float(U.Layers[0].bias.detach().abs().max()) == 0.0
    V = P.Network([2, 40, 40, 40, 40, 40, 1], "Rational")
    V.Set_State(U.Get_State())
    x = torch.rand(5, 2)
    assert torch.equal(U(x), V(x))
    with pytest.raises(ValueError):
        engine.network_signature(P.Network([2, 8, 1], "Sigmoid"))
    with pytest.raises(TypeError):
        engine.network_signature(lambda X: X)
    with pytest.raises(ValueError):
        P.Network([2, 8, 1], "Sin")                     # accepted by the reference's reader, not by its Network


def test_library_key_pads_and_dedups():
    d, _, derivs, lhs, rhs = configs.config("rd2d")
    encs, terms = engine.library_key(derivs, lhs, rhs)
    assert (0, 0, 1, 0) in encs and (0, 0, 0, 0) in encs and len(encs) == 6
    assert terms[0] == ((encs.index((1, 0, 0, 0)), 1),) and len(terms) == 22


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pde_learn_b200.Test_Train import _global_counts, allreduce_gradients
    a = torch.nn.Parameter(torch.zeros(3, 2)); b = torch.nn.Parameter(torch.zeros(5)); c = torch.nn.Parameter(torch.zeros(1))
    a.grad = torch.full((3, 2), float(rank + 1)); b.grad = torch.arange(5.0) * (rank + 1)     # c.grad stays None
    allreduce_gradients([a, b, c], dist.group.WORLD)
    counts = _global_counts([100 + rank, 7], dist.group.WORLD)
    q.put((rank, a.grad.clone(), b.grad.clone(), c.grad, counts))
    dist.destroy_process_group()


def test_gradient_allreduce_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, ga, gb, gc, counts in got:
        assert torch.equal(ga, torch.full((3, 2), 3.0))
        assert torch.equal(gb, torch.arange(5.0) * 3)
        assert gc is None and counts == [201, 14]


def test_settings_reader_matches_reference_reader(tmp_path):
    """Our reader on the reference's shipped Settings.txt (copied text below) == the reference reader's dictionary."""
    ref = json.load(open(os.path.join(GOLDEN_DIR, "settings_txt_parsed.json")))
    text = """
# Save, Load settings.
Load U from Save [bool]:                         False
Load Xi, Library from Save [bool]:               False
Load Optimizer from Save [bool]:                 False
    Load File Name [str]:
Library File [str]:                              Library
Hidden Layer Widths [List of int]:               20, 20, 20, 20, 20
Hidden Activation Function [str]:                Rat
Train on CPU or GPU [GPU, CPU]:                  gpu
p [float]:                                       .1
Weights [Dict of float]:                         {"Data" : 1.0, "Coll" : 1.0, "Lp" : 0.000, "L2" : 0.000005}
Number of Training Collocation Points [int]:     3000
Number of Testing Collocation Points [int]:      1000
Mask Small Xi Components [bool]:                 True     # trailing comment
Optimizer [Adam, LBFGS]:                         Adam
Learning Rate [float]:                           .001
Number of Epochs [int]:                          1000
DataSet Names [List of str]:                     [Burgers_Sine_N75_P2000]
"""
    f = tmp_path / "Settings.txt"
    f.write_text(text)
    S = P.Settings_Reader(str(f))
    for k, v in ref.items():
        if k in ("Device", "Library Path"):
            continue
        assert S[k] == v, k
    assert S["Device"].type == "cuda" and S["Library Path"].endswith("Library.txt")
    f.write_text(text.replace("gpu", "cpu"))
    with pytest.raises(P.Read_Error):
        P.Settings_Reader(str(f))                      # no CPU path: refuse instead of silently switching


def test_save_load_state_round_trip(tmp_path):
    torch.manual_seed(1)
    d, _, derivs, lhs, rhs = configs.config("heat")
    U = P.Network([2, 12, 12, 1], "Rational")
    Xi = torch.randn(len(rhs), requires_grad=True)
    opt = torch.optim.Adam(list(U.parameters()) + [Xi], lr=1e-3)
    path = str(tmp_path / "save")
    P.Save_State(path, [U], Xi, opt, derivs, lhs, rhs, ["Heat"])
    blob = torch.load(path, weights_only=False)
    assert sorted(blob.keys()) == sorted(["U States", "Xi", "Optimizer", "Derivative Encodings", "LHS Term State",
                                          "RHS Term States", "DataSet Names"])          # Code/main.py:517-523
    st = P.Load_State(path, torch.device("cpu"))
    V = st["U List"][0]
    x = torch.rand(4, 2)
    assert torch.equal(U(x), V(x)) and torch.equal(st["Xi"], Xi)
    assert [D.Encoding.tolist() for D in st["Derivatives"]] == [D.Encoding.tolist() for D in derivs]
    assert str(st["LHS Term"]) == str(lhs) and [str(t) for t in st["RHS Terms"]] == [str(t) for t in rhs]
    assert P.Threshold_Xi(torch.tensor([0.1, 1e-4, -0.2]), torch.tensor([False, False, True])) == [0]
    assert P.Build_Mask(torch.tensor([0.1, 1e-4, -0.2])).tolist() == [False, True, False]


@pytest.mark.parametrize("dim,args", [("1d", (0.25, 120, 40)), ("2d", (0.25, 90, 30))])
def test_dataset_preparation_matches_reference_files(dim, args, tmp_path):
    """Data_Sets.From_MATLAB_* writes, bit for bit, the file the reference's Data/From_MATLAB.py wrote from the same .mat with the
    same numpy seed (fixtures: tests/golden/make_golden_dataset.py)."""
    import os
    import numpy as np
    from pde_learn_b200 import Data_Sets
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    f = Data_Sets.From_MATLAB_1D if dim == "1d" else Data_Sets.From_MATLAB_2D
    path = f(os.path.join(gold, "synthetic_%s.mat" % dim), *args, Directory=str(tmp_path), Seed=11)
    assert os.path.basename(path) == "synthetic_%s_N25_P%d.npz" % (dim, args[1])
    got, want = np.load(path), np.load(os.path.join(gold, "dataset_%s.npz" % dim))
    assert sorted(got.files) == sorted(want.files)
    for k in want.files:
        assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape, k
        assert np.array_equal(got[k], want[k]), k
    assert Data_Sets.From_MATLAB(os.path.join(gold, "synthetic_%s.mat" % dim), *args, Directory=str(tmp_path), Seed=11) == path


def test_reference_written_save_loads():
    """A save file written by the reference's own main.py (tests/golden/make_reference_save.py; Code/main.py:487-524) is read by
    Load_State: network values, Xi, library and Adam moments equal what the reference itself sees in that file."""
    meta = json.load(open(os.path.join(GOLDEN_DIR, "reference_save_burgers.json")))
    st = P.Load_State(os.path.join(GOLDEN_DIR, "reference_save_burgers.pt"), torch.device("cpu"))
    U = st["U List"][0]
    assert list(U.Widths) == meta["widths"] and [U._Get_Activation_String(a) for a in U.Activation_Functions] == meta["activation_types"]
    X = torch.tensor(meta["probe_inputs"], dtype=torch.float32)
    assert np.allclose(U(X).detach().double().view(-1).numpy(), np.array(meta["probe_values"]), rtol=0, atol=1e-7)
    assert np.array_equal(st["Xi"].detach().double().numpy(), np.array(meta["xi"])) and st["Xi"].requires_grad
    assert [D.Encoding.tolist() for D in st["Derivatives"]] == meta["derivative_encodings"]
    assert len(st["RHS Terms"]) == meta["num_rhs_terms"] and st["DataSet Names"] == meta["dataset_names"]
    lhs, rhs = configs.config("burgers")[3:]
    assert str(st["LHS Term"]) == str(lhs) and [str(t) for t in st["RHS Terms"]] == [str(t) for t in rhs]   # = the shipped Library.txt
    opt = torch.optim.Adam(list(U.parameters()) + [st["Xi"]], lr=1e-3)
    opt.load_state_dict(st["Optimizer State"])                               # main.py:232-246
    assert len(opt.state_dict()["state"]) == meta["num_param_tensors"]
    assert float(opt.state_dict()["state"][0]["step"]) == meta["adam_step"]


def test_repo_written_save_loads_in_reference(tmp_path):
    """The other direction: a file written by Save_State is opened by the REFERENCE's classes (Network.Set_State, Network.py:258-283;
    Build_Term_From_State, Term.py:149-177; Derivative) and its optimiser state loads into a torch Adam over the reference network's
    parameters, as main.py:51-57,140-160,232-246 does.  Needs /root/reference (this container); skipped on the GPU box."""
    import subprocess
    import sys
    if not os.path.isdir("/root/reference/Code"):
        pytest.skip("the reference tree is not on this machine")
    torch.manual_seed(5)
    d, _, derivs, lhs, rhs = configs.config("kdv")
    U = P.Network([2, 16, 12, 1], "Rational")
    with torch.no_grad():
        U.Activation_Functions[1].a.add_(0.01)
    Xi = torch.randn(len(rhs), requires_grad=True)
    opt = torch.optim.Adam(list(U.parameters()) + [Xi], lr=1e-3)
    (U(torch.rand(8, 2)).sum() + Xi.sum()).backward(); opt.step()
    X = torch.rand(5, 2)
    path = str(tmp_path / "save")
    P.Save_State(path, [U], Xi, opt, derivs, lhs, rhs, ["KdV_Sine_N10_P500"])
    code = """
import sys, json, numpy, torch
for p in ("/root/reference/Code", "/root/reference/Code/Classes", "/root/reference/Code/Readers"): sys.path.insert(0, p)
from Network import Network
from Term import Build_Term_From_State
from Derivative import Derivative
blob = torch.load(sys.argv[1], weights_only=False)
st = blob["U States"][0]
U = Network(Widths=st["Widths"], Hidden_Activation=st["Activation Types"][0], Output_Activation=st["Activation Types"][-1])
U.Set_State(st)
Xi = blob["Xi"].detach().clone().requires_grad_(True)
opt = torch.optim.Adam(list(U.parameters()) + [Xi], lr=1e-3); opt.load_state_dict(blob["Optimizer"])
D = [Derivative(Encoding=e) for e in blob["Derivative Encodings"]]
T = [Build_Term_From_State(s) for s in blob["RHS Term States"]]; L = Build_Term_From_State(blob["LHS Term State"])
X = torch.tensor(json.loads(sys.argv[2]), dtype=torch.float32)
print("OUT " + json.dumps({"u": U(X).detach().double().view(-1).tolist(), "xi": Xi.detach().double().tolist(), "lhs": str(L),
      "rhs": [str(t) for t in T], "enc": [[int(v) for v in d.Encoding] for d in D], "names": blob["DataSet Names"],
      "step": float(opt.state_dict()["state"][0]["step"])}))
"""
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", code, path, json.dumps(X.tolist())], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    got = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("OUT ")][0][4:])
    assert np.allclose(np.array(got["u"]), U(X).detach().double().view(-1).numpy(), rtol=0, atol=1e-7)
    assert np.array_equal(np.array(got["xi"]), Xi.detach().double().numpy())
    assert got["lhs"] == str(lhs) and got["rhs"] == [str(t) for t in rhs]
    assert got["enc"] == [[int(v) for v in D.Encoding] for D in derivs] and got["names"] == ["KdV_Sine_N10_P500"] and got["step"] == 1.0


def test_library_key_sees_in_place_term_changes():
    """Term.Append (public API kept from the reference, Code/Classes/Term.py) changes a term in place: the memoised library key must
    not keep pointing at the old plan."""
    d, _, derivs, lhs, rhs = configs.config("heat")
    rhs = list(rhs)
    k1 = engine.library_key(derivs, lhs, rhs)
    assert engine.library_key(derivs, lhs, rhs) is k1                      # memo hit on the same objects
    rhs[0].Append(derivs[-1], 2)
    k2 = engine.library_key(derivs, lhs, rhs)
    assert k2 != k1 and len(k2[1][1]) == len(k1[1][1]) + 1                  # the first RHS term gained a sub-term


def test_validate_configuration_reports_kernel_limits():
    d, _, derivs, lhs, rhs = configs.config("ks")
    engine.validate_configuration(2, [40] * 5, "Rat", derivs, lhs, rhs)
    for widths, act in (([80] * 3, "Tanh"), ([20] * 3, "Sigmoid")):
        with pytest.raises(ValueError):
            engine.validate_configuration(2, widths, act, derivs, lhs, rhs)
