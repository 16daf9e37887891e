"""Shared parity helpers for the GPU tests (test infrastructure)."""
import torch

ID_MARGIN_TOL = 2e-2   # = 2 x the north star's per-logit tolerance (1e-2 of max|logit|)


def rel(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-6)).item()


def check_greedy_ids(gen, ref, margins, what=""):
    """Greedy ids must equal the oracle's wherever the oracle's (top-1 - top-2) / max|logit| margin exceeds
    ID_MARGIN_TOL: fp16 rounding may only flip a near-tie.  After such a flip the two runs see different inputs and
    are no longer comparable, so the comparison stops there.  Returns the number of leading ids that agree."""
    n = min(len(gen), len(ref))
    i = next((j for j in range(n) if gen[j] != ref[j]), n)
    if i < n:
        assert i < len(margins) and margins[i] < ID_MARGIN_TOL, \
            f"{what}: id {i} differs ({gen[i]} vs oracle {ref[i]}) where the oracle margin is {margins[i]:.4f} >= {ID_MARGIN_TOL}"
    return i
