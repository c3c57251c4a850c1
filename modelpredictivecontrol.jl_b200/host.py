"""Host-side helpers that mirror the reference's constructors (one-off, not on the hot path)."""
import numpy as np


def move_blocking(Hp, Hc):
    """move_blocking (reference src/controller/construct.jl:629-660)."""
    if np.isscalar(Hc):
        Hc = int(Hc)
        if Hc < 1:
            raise ValueError("Control horizon Hc should be >= 1")
        nb = [1] * Hc
        nb[-1] = Hp - Hc + 1
        if nb[-1] < 1:
            raise ValueError("Control horizon Hc should be <= prediction horizon Hp")
        return nb
    nb = [int(v) for v in Hc]
    if not all(v > 0 for v in nb):
        raise ValueError("Move blocking vector must be strictly positive integers.")
    if sum(nb) < Hp:
        nb = nb + [Hp - sum(nb)]
    elif sum(nb) > Hp:
        cs = np.cumsum(nb)
        nb = nb[: int(np.flatnonzero(cs >= Hp)[0]) + 1]
        if sum(nb) > Hp:
            nb[-1] = Hp - sum(nb[:-1])
    return nb
