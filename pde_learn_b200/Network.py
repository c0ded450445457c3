"""
System-response network U(t, x[, y[, z]]).  Same module tree and parameter order as the reference
(Code/Classes/Network.py:63-129): `Layers` = ModuleList of nn.Linear (float32, Xavier-uniform W, zero b),
`Activation_Functions` = one module per layer (a separate trainable `Rational` per hidden layer, Identity
on the output), so state dicts, `Get_State`/`Set_State` and optimiser state are interchangeable.

`forward` here is the plain value-only evaluation used by callers that just want U(X) (plots, error
maps).  The loss path never calls it: losses run in the fused CUDA kernels (see Loss.py).
"""
from typing import Dict, List

import torch


class Rational(torch.nn.Module):
    """R(x) = (a0 + a1 x + a2 x^2 + a3 x^3) / (b0 + b1 x + b2 x^2), trainable (Network.py:7-59)."""

    def __init__(self, Device=torch.device("cpu")):
        super().__init__()
        self.a = torch.nn.Parameter(torch.tensor((0.0218, 0.5, 1.5957, 1.1915), dtype=torch.float32, device=Device))
        self.b = torch.nn.Parameter(torch.tensor((1.0, 0.0, 2.3830), dtype=torch.float32, device=Device))

    def forward(self, X):
        a, b = self.a, self.b
        return (a[0] + X * (a[1] + X * (a[2] + a[3] * X))) / (b[0] + X * (b[1] + b[2] * X))


_ACTIVATIONS = {"none": torch.nn.Identity, "tanh": torch.nn.Tanh, "sigmoid": torch.nn.Sigmoid, "elu": torch.nn.ELU,
                "softmax": torch.nn.Softmax}


class Network(torch.nn.Module):
    def __init__(self, Widths: List[int], Hidden_Activation: str, Output_Activation: str = "None",
                 Device: torch.device = torch.device("cpu")):
        for i, w in enumerate(Widths):
            assert w > 0, "Layer widths must be positive got Layer[%u] = %d" % (i, w)
        super().__init__()
        self.Widths = list(Widths)
        self.Num_Layers = len(Widths) - 1
        self.Num_Hidden_Layers = self.Num_Layers - 1
        self.Layers = torch.nn.ModuleList()
        for i in range(self.Num_Layers):
            self.Layers.append(torch.nn.Linear(Widths[i], Widths[i + 1], bias=True).to(dtype=torch.float32, device=Device))
        for L in self.Layers:
            torch.nn.init.xavier_uniform_(L.weight)
            torch.nn.init.zeros_(L.bias)
        self.Activation_Functions = torch.nn.ModuleList()
        for _ in range(self.Num_Hidden_Layers):
            self.Activation_Functions.append(self._Get_Activation_Function(Hidden_Activation, Device))
        self.Activation_Functions.append(self._Get_Activation_Function(Output_Activation, Device))

    def _Get_Activation_Function(self, Encoding: str, Device) -> torch.nn.Module:
        key = Encoding.strip().lower()
        if key == "rational":
            return Rational(Device=Device)
        if key not in _ACTIVATIONS:
            raise ValueError("Unknown Activation Function. Got %s" % Encoding)   # the reference calls exit()
        return _ACTIVATIONS[key]()

    def _Get_Activation_String(self, Activation: torch.nn.Module) -> str:
        if isinstance(Activation, Rational):
            return "Rational"
        for name, cls in (("None", torch.nn.Identity), ("Tanh", torch.nn.Tanh), ("Sigmoid", torch.nn.Sigmoid),
                          ("Elu", torch.nn.ELU), ("Softmax", torch.nn.Softmax)):
            if isinstance(Activation, cls):
                return name
        raise ValueError("Unknown Activation Function. Got %s" % str(type(Activation)))

    def Get_State(self) -> Dict:
        return {"Widths": self.Widths,
                "Layers": [L.state_dict() for L in self.Layers],
                "Activation Types": [self._Get_Activation_String(A) for A in self.Activation_Functions],
                "Activation States": [A.state_dict() for A in self.Activation_Functions]}

    def Set_State(self, State: Dict) -> None:
        assert list(self.Widths) == list(State["Widths"])
        for A, name in zip(self.Activation_Functions, State["Activation Types"]):
            assert self._Get_Activation_String(A) == name
        for L, sd in zip(self.Layers, State["Layers"]):
            L.load_state_dict(sd)
        for A, sd in zip(self.Activation_Functions, State["Activation States"]):
            A.load_state_dict(sd)

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        for L, A in zip(self.Layers, self.Activation_Functions):
            X = A(L(X))
        return X
