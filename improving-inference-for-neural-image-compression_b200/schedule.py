"""Temperature schedule of the SGA loop (host scalar per step).

Mirrors utils.annealed_temperature (/root/reference/utils.py:151-180) as called at sga.py:211 with
scheme='exp0', ub=0.5: tau = ub*exp(-r*(t-t0)) clipped to [lb, ub]."""
import math


def annealed_temperature(t, r, ub, lb=1e-8, scheme="exp", t0=700):
    if scheme == "exp":
        tau = math.exp(-r * t)
    elif scheme == "exp0":
        tau = ub * math.exp(-r * (t - t0))
    elif scheme == "linear":
        tau = -r * (t - t0) + ub
    else:
        raise NotImplementedError
    return min(max(tau, lb), ub)
