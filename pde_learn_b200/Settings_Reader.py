"""
Settings_Reader(Path) -> Dict: reader for the reference's Settings.txt (format: `Phrase [type]: value` lines,
'#' comments; Code/Readers/Settings_Reader.py:9-144, File_Reader.py).  Produces the same keys as the reference:
"Load U", "Load Xi, Library", "Load Optimizer", ["Load File Name"], ["Library Path"], ["Hidden Layer Widths",
"Hidden Activation Function"], "Device", "p", "Weights", "Num Train Coll Points", "Num Test Coll Points",
"Mask Small Xi Components", "Optimizer", "Learning Rate", "Num Epochs", ["DataSet Names"].

Differences, on purpose: the path is an argument (the reference hard-codes "../Settings.txt"); phrases are looked
up anywhere in the file instead of strictly in file order; "Device" is always a CUDA device here -- a request for
the CPU raises, because this engine has no CPU path.
"""
import os
from typing import Dict

import torch


class Read_Error(Exception):
    pass


def _lines(path):
    with open(path, "r") as f:
        for raw in f:
            line = raw.split("#", 1)[0].strip()
            if line:
                yield line


def _table(path) -> Dict[str, str]:
    out = {}
    for line in _lines(path):
        if "]:" in line:
            key, val = line.split("]:", 1)
            out[(key + "]:").strip().lower()] = val.strip()
    return out


def _get(tab, phrase):
    k = phrase.strip().lower()
    if k not in tab:
        raise Read_Error("Settings file has no line starting with %r" % phrase)
    return tab[k]


def _bool(tab, phrase):
    v = _get(tab, phrase)
    if v[:1] in ("T", "t"):
        return True
    if v[:1] in ("F", "f"):
        return False
    raise Read_Error("%s should be True or False, got %r" % (phrase, v))


def _list(v):
    return [s.strip() for s in v.strip().strip("[]").split(",") if s.strip()]


def _dict(v):
    out = {}
    for item in v.strip().strip("{}").split(","):
        if ":" in item:
            k, x = item.split(":", 1)
            out[k.strip().strip('"').strip("'")] = float(x)
    return out


def Settings_Reader(Path: str = "../Settings.txt") -> Dict:
    tab = _table(Path)
    S: Dict = {}
    S["Load U"] = _bool(tab, "Load U from Save [bool]:")
    S["Load Xi, Library"] = _bool(tab, "Load Xi, Library from Save [bool]:")
    S["Load Optimizer"] = _bool(tab, "Load Optimizer from Save [bool]:")
    if S["Load U"] or S["Load Xi, Library"] or S["Load Optimizer"]:
        S["Load File Name"] = _get(tab, "Load File Name [str]:").strip()
    if not S["Load Xi, Library"]:
        S["Library Path"] = os.path.join(os.path.dirname(os.path.abspath(Path)), _get(tab, "Library File [str]:") + ".txt")
    if not S["Load U"]:
        S["Hidden Layer Widths"] = [int(v) for v in _list(_get(tab, "Hidden Layer Widths [List of int]:"))]
        a = _get(tab, "Hidden Activation Function [str]:")
        if a[:1] in ("R", "r"):
            S["Hidden Activation Function"] = "Rational"
        elif a[:1] in ("T", "t"):
            S["Hidden Activation Function"] = "Tanh"
        else:
            raise Read_Error("Hidden Activation Function should be Tanh or Rational (Sin is accepted by the reference's "
                             "reader but not implemented by its Network), got %r" % a)
    dev = _get(tab, "Train on CPU or GPU [GPU, CPU]:")
    if dev[:1] in ("G", "g"):
        S["Device"] = torch.device("cuda")
    elif dev[:1] in ("C", "c"):
        raise Read_Error("Settings ask for the CPU; the B200 engine has no CPU path. Set 'Train on CPU or GPU' to GPU.")
    else:
        raise Read_Error("\"Train on CPU or GPU\" should be \"CPU\" or \"GPU\". Got " + dev)
    S["p"] = float(_get(tab, "p [float]:"))
    S["Weights"] = _dict(_get(tab, "Weights [Dict of float]:"))
    S["Num Train Coll Points"] = int(_get(tab, "Number of Training Collocation Points [int]:"))
    S["Num Test Coll Points"] = int(_get(tab, "Number of Testing Collocation Points [int]:"))
    S["Mask Small Xi Components"] = _bool(tab, "Mask Small Xi Components [bool]:")
    o = _get(tab, "Optimizer [Adam, LBFGS]:")
    if o[:1] in ("A", "a"):
        S["Optimizer"] = "Adam"
    elif o[:1] in ("L", "l"):
        S["Optimizer"] = "LBFGS"
    else:
        raise Read_Error("\"Optimizer [Adam, LBFGS]:\" should be \"Adam\" or \"LBFGS\". Got " + o)
    S["Learning Rate"] = float(_get(tab, "Learning Rate [float]:"))
    S["Num Epochs"] = int(_get(tab, "Number of Epochs [int]:"))
    if not S["Load U"]:
        S["DataSet Names"] = _list(_get(tab, "DataSet Names [List of str]:"))
    return S
