"""
pde_learn_b200 -- B200-native (sm_100a) collocation-loss engine with the function-level API of
punkduckable/PDE-LEARN's hot path (Loss.py, Evaluate_Derivatives.py, Points.py, Test_Train.py).

    from pde_learn_b200 import Network, Derivative, Term
    from pde_learn_b200 import Coll_Loss, Data_Loss, Lp_Loss, L2_Squared_Loss
    from pde_learn_b200 import Generate_Points, Training, Testing, Evaluate_Derivatives

All numerical work runs in hand-written CUDA kernels behind the C ABI in include/pdelearn_b200.h.
No CPU fallback exists: without the built shared library or without a CUDA device the calls raise.
"""
from ._lib import NativeLibraryError, lib                                   # noqa: F401
from .Derivative import Derivative, Get_Order                               # noqa: F401
from .Term import Term, Build_Term_From_State                               # noqa: F401
from .Network import Network, Rational                                      # noqa: F401
from .Loss import Coll_Loss, Data_Loss, Lp_Loss, L2_Squared_Loss            # noqa: F401
from .Evaluate_Derivatives import (Evaluate_Derivatives, Derivative_From_Derivative, Evaluate_Residual,
                                   Evaluate_Network)                       # noqa: F401
from .Points import Generate_Points                                         # noqa: F401
from .Test_Train import Training, Testing, allreduce_gradients              # noqa: F401
from .engine import closure_total as Closure_Loss                          # noqa: F401
from .Library_Reader import Read_Library, Parse_Term                        # noqa: F401
from .Settings_Reader import Settings_Reader, Read_Error                     # noqa: F401
from .Driver import (Update_Targeted_Points, Run_Epochs, Build_Mask, Threshold_Xi, Save_State, Load_State)   # noqa: F401
from .Epoch_Graph import Captured_Epochs                                   # noqa: F401
