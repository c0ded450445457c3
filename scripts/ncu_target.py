"""Small target for ncu: a few fused fwd+adjoint launches of one workload (default heat 2^20 rows, tanh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.perf_probe import probe

if __name__ == "__main__":
    w = sys.argv[1] if len(sys.argv) > 1 else "heat"
    act = sys.argv[2] if len(sys.argv) > 2 else "Tanh"
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
    print(probe(w, act, B, reps=1))
