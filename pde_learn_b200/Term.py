"""
Library term  prod_i (D_i U)^{p_i}.  Same public surface as the reference class
(Code/Classes/Term.py:9-177): `.Derivatives`, `.Powers`, `.Num_Sub_Terms`, `Append`, `Get_State`,
`Build_Term_From_State`, `__str__`.
"""
from typing import Dict, List

from .Derivative import Derivative


class Term:
    def __init__(self, Derivatives: List[Derivative], Powers: List[int]):
        assert len(Derivatives) == len(Powers)
        assert all(int(p) >= 1 for p in Powers)          # Term.py:57-58
        self.Derivatives = list(Derivatives)
        self.Powers = [int(p) for p in Powers]
        self.Num_Sub_Terms = len(self.Powers)

    def Append(self, Derivative: Derivative, Power: int) -> None:
        assert Power >= 0
        self.Derivatives.append(Derivative)
        self.Powers.append(int(Power))
        self.Num_Sub_Terms += 1

    def Get_State(self) -> Dict:
        return {"Powers": self.Powers, "Derivative Encodings": [D.Encoding for D in self.Derivatives]}

    def __str__(self) -> str:
        out = []
        for D, p in zip(self.Derivatives, self.Powers):
            s = "(" + str(D) + "U)"
            if p > 1:
                s += "^" + str(p)
            out.append(s)
        return "*".join(out)


def Build_Term_From_State(State: Dict) -> Term:
    return Term(Derivatives=[Derivative(Encoding=e) for e in State["Derivative Encodings"]],
                Powers=list(State["Powers"]))
