"""
Symbolic partial-derivative operator.  Same public surface as the reference class
(Code/Classes/Derivative.py:5-150: `.Encoding` int32 array of length 2..4, `.Order`, `Is_Child_Of`,
`__str__`), so objects of either class can be handed to Coll_Loss / Training.
"""
import numpy

_AXES = ("t", "x", "y", "z")


class Derivative:
    def __init__(self, Encoding):
        enc = numpy.asarray(Encoding).astype(numpy.int32)
        assert enc.ndim == 1
        assert enc.size in (2, 3, 4)                     # Derivative.py:46
        assert bool((enc >= 0).all())
        self.Encoding = enc
        self.Order = int(enc.sum())

    def __str__(self):
        parts = []
        for axis, n in enumerate(self.Encoding.tolist()):
            if n == 1:
                parts.append("D_%s " % _AXES[axis])
            elif n > 1:
                parts.append("D_%s^%u " % (_AXES[axis], n))
        return "".join(parts)

    def Is_Child_Of(self, D):
        """True when D = (some derivative operator) applied after self (Derivative.py:101-140)."""
        mine, other = self.Encoding.tolist(), D.Encoding.tolist()
        common = min(len(mine), len(other))
        if any(mine[k] > other[k] for k in range(common)):
            return False
        # the reference skips index `common` itself here (range(l+1, n)); kept for identical answers
        return all(v == 0 for v in mine[common + 1:])


def Get_Order(D):
    return D.Order
