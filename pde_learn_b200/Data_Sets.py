"""
Data-set preparation: MATLAB solution file -> the `.npz` data set the driver reads.

Same files, field names, dtypes and random draws as the reference (Data/From_MATLAB.py:37-263, Data/Create_Data_Set.py:5-56):
coordinates float32, targets float64 (float32 solution + float64 noise), bounds float32 [d][2]; noise =
Noise_Proportion * std(usol) * randn, then two `numpy.random.choice(..., replace=False)` draws (train, test) from the global
numpy generator, in that order -- so seeding numpy the same way reproduces the reference's file bit for bit.  Host-side numpy
only; the coordinate grids are built by broadcasting instead of a triple Python loop.  Plotting is out of scope.

    python -m pde_learn_b200.Data_Sets Burgers_Sine.mat --noise 0.25 --train 4000 --test 1000 --out Data/DataSets
"""
import argparse
import os
from typing import Optional

import numpy


def Create_Data_Set(Name: str, Train_Inputs: numpy.ndarray, Train_Targets: numpy.ndarray, Test_Inputs: numpy.ndarray,
                    Test_Targets: numpy.ndarray, Input_Bounds: numpy.ndarray, Directory: str = "./DataSets") -> str:
    """Writes <Directory>/<Name>.npz with the reference's five fields (Data/Create_Data_Set.py:43-52); returns the path."""
    os.makedirs(Directory, exist_ok=True)
    path = os.path.join(Directory, Name + ".npz")
    with open(path, "wb") as f:
        numpy.savez(file=f, Train_Inputs=Train_Inputs, Train_Targets=Train_Targets, Test_Inputs=Test_Inputs,
                    Test_Targets=Test_Targets, Input_Bounds=Input_Bounds)
    return path


def _load(path: str):
    import scipy.io
    return scipy.io.loadmat(path)


def _finish(Data_File_Path, coords, values, bounds, Noise_Proportion, Num_Train_Examples, Num_Test_Examples, Directory):
    n = coords.shape[0]
    train = numpy.random.choice(n, Num_Train_Examples, replace=False)        # From_MATLAB.py:121 / :247
    test = numpy.random.choice(n, Num_Test_Examples, replace=False)          # From_MATLAB.py:122 / :248
    base = os.path.splitext(os.path.basename(Data_File_Path))[0]
    name = base + "_N" + str(int(100 * Noise_Proportion)) + "_P" + str(Num_Train_Examples)   # From_MATLAB.py:132-134
    return Create_Data_Set(name, coords[train, :], values[train], coords[test, :], values[test], bounds, Directory)


def From_MATLAB_1D(Data_File_Path: str, Noise_Proportion: float, Num_Train_Examples: int, Num_Test_Examples: int,
                   Directory: str = "./DataSets", Seed: Optional[int] = None) -> str:
    """One space dimension: fields t [1,Nt], x [1,Nx], usol [Nx,Nt] (Data/From_MATLAB.py:37-141).  Rows are (t, x)."""
    if Seed is not None:
        numpy.random.seed(Seed)
    D = _load(Data_File_Path)
    t = D["t"].reshape(-1).astype(numpy.float32)
    x = D["x"].reshape(-1).astype(numpy.float32)
    u = numpy.real(D["usol"]).astype(numpy.float32)
    assert u.shape == (x.size, t.size), "usol must be [Nx, Nt]"
    bounds = numpy.array([[t[0], t[-1]], [x[0], x[-1]]], dtype=numpy.float32)
    noisy = u + Noise_Proportion * numpy.std(u) * numpy.random.randn(*u.shape)          # float64, From_MATLAB.py:91
    tt, xx = numpy.meshgrid(t, x)                                                      # [Nx, Nt]: row = x index
    coords = numpy.hstack((tt.reshape(-1, 1), xx.reshape(-1, 1)))
    return _finish(Data_File_Path, coords, noisy.reshape(-1), bounds, Noise_Proportion, Num_Train_Examples, Num_Test_Examples, Directory)


def From_MATLAB_2D(Data_File_Path: str, Noise_Proportion: float, Num_Train_Examples: int, Num_Test_Examples: int,
                   Directory: str = "./DataSets", Seed: Optional[int] = None) -> str:
    """Two space dimensions: fields t, x, y and usol [Nt,Nx,Ny] (Data/From_MATLAB.py:152-263).  Rows are (t, x, y)."""
    if Seed is not None:
        numpy.random.seed(Seed)
    D = _load(Data_File_Path)
    t = D["t"].reshape(-1).astype(numpy.float32)
    x = D["x"].reshape(-1).astype(numpy.float32)
    y = D["y"].reshape(-1).astype(numpy.float32)
    u = numpy.real(D["usol"]).astype(numpy.float32)
    assert u.shape == (t.size, x.size, y.size), "usol must be [Nt, Nx, Ny]"
    bounds = numpy.array([[t[0], t[-1]], [x[0], x[-1]], [y[0], y[-1]]], dtype=numpy.float32)
    noisy = u + Noise_Proportion * numpy.std(u) * numpy.random.randn(*u.shape)
    shape = (t.size, x.size, y.size)
    tt = numpy.broadcast_to(t.reshape(-1, 1, 1), shape); xx = numpy.broadcast_to(x.reshape(1, -1, 1), shape)
    yy = numpy.broadcast_to(y.reshape(1, 1, -1), shape)
    coords = numpy.hstack((tt.reshape(-1, 1), xx.reshape(-1, 1), yy.reshape(-1, 1))).astype(numpy.float32)
    return _finish(Data_File_Path, coords, noisy.reshape(-1), bounds, Noise_Proportion, Num_Train_Examples, Num_Test_Examples, Directory)


def From_MATLAB(Data_File_Path: str, Noise_Proportion: float, Num_Train_Examples: int, Num_Test_Examples: int,
                Directory: str = "./DataSets", Seed: Optional[int] = None) -> str:
    """Dispatches on the presence of a `y` field (the reference asks for Num_Spatial_Dimensions, From_MATLAB.py:14-32)."""
    f = From_MATLAB_2D if "y" in _load(Data_File_Path) else From_MATLAB_1D
    return f(Data_File_Path, Noise_Proportion, Num_Train_Examples, Num_Test_Examples, Directory, Seed)


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("mat"); ap.add_argument("--noise", type=float, default=0.0); ap.add_argument("--train", type=int, default=10000)
    ap.add_argument("--test", type=int, default=1000); ap.add_argument("--out", default="./DataSets"); ap.add_argument("--seed", type=int, default=None)
    a = ap.parse_args()
    print(From_MATLAB(a.mat, a.noise, a.train, a.test, a.out, a.seed))
