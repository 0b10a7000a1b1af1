"""Temperature ladder — host-side mirror of ParallelTempering.default_beta_ladder /
assign_temperatures (_revamp.py:835-930). The reference only ever calls it as
``default_beta_ladder(2, ntemps=num_chains, Tmax=maxtemp)`` (_revamp.py:919), for which the
ladder is the closed form T_i = maxtemp ** (i / (R - 1)); the ptemcee step table only matters
when Tmax is None (all 100 entries are kept; beyond ndim = 100 the same source's large-dimension
approximation applies)."""
from __future__ import annotations

import numpy as np

_TSTEP = (  # ptemcee's tuned temperature step per dimension, ndim = 1..100 (the numeric contract of RV:854-874)
    25.2741, 7.0, 4.47502, 3.5236, 3.0232, 2.71225, 2.49879, 2.34226, 2.22198, 2.12628,
    2.04807, 1.98276, 1.92728, 1.87946, 1.83774, 1.80096, 1.76826, 1.73895, 1.7125, 1.68849,
    1.66657, 1.64647, 1.62795, 1.61083, 1.59494, 1.58014, 1.56632, 1.55338, 1.54123, 1.5298,
    1.51901, 1.50881, 1.49916, 1.49, 1.4813, 1.47302, 1.46512, 1.45759, 1.45039, 1.4435,
    1.4369, 1.43056, 1.42448, 1.41864, 1.41302, 1.40761, 1.40239, 1.39736, 1.3925, 1.38781,
    1.38327, 1.37888, 1.37463, 1.37051, 1.36652, 1.36265, 1.35889, 1.35524, 1.3517, 1.34825,
    1.3449, 1.34164, 1.33847, 1.33538, 1.33236, 1.32943, 1.32656, 1.32377, 1.32104, 1.31838,
    1.31578, 1.31325, 1.31076, 1.30834, 1.30596, 1.30364, 1.30137, 1.29915, 1.29697, 1.29484,
    1.29275, 1.29071, 1.2887, 1.28673, 1.2848, 1.28291, 1.28106, 1.27923, 1.27745, 1.27569,
    1.27397, 1.27227, 1.27061, 1.26898, 1.26737, 1.26579, 1.26424, 1.26271, 1.26121, 1.25973,
)


def default_beta_ladder(ndim, ntemps=None, Tmax=None):
    if type(ndim) != int or ndim < 1:
        raise ValueError('Invalid number of dimensions specified.')
    if ntemps is None and Tmax is None:
        raise ValueError('Must specify one of ``ntemps`` and ``Tmax``.')
    if Tmax is not None and Tmax <= 1:
        raise ValueError('``Tmax`` must be greater than 1.')
    if ntemps is not None and (type(ntemps) != int or ntemps < 1):
        raise ValueError('Invalid number of temperatures specified.')
    tstep = _TSTEP[ndim - 1] if ndim <= len(_TSTEP) else 1.0 + 2.0 * np.sqrt(np.log(4.0)) / np.sqrt(ndim)
    append_inf = False
    if Tmax == np.inf:
        append_inf, Tmax, ntemps = True, None, ntemps - 1
    if ntemps is not None:
        if Tmax is None:
            Tmax = tstep ** (ntemps - 1)
    else:
        if Tmax is None:
            raise ValueError('Must specify at least one of ``ntemps`` and finite ``Tmax``.')
        ntemps = int(np.log(Tmax) / np.log(tstep) + 2)
    betas = np.logspace(0, -np.log10(Tmax), ntemps)
    if append_inf:
        betas = np.concatenate((betas, [0]))
    return betas


def assign_temperatures(num_chains: int, maxtemp: float, geometric: bool = True) -> np.ndarray:
    """_revamp.py:909-930. geometric: 1/beta_i; linear: 1 + i*maxtemp/num_chains."""
    if geometric:
        betas = default_beta_ladder(2, ntemps=int(num_chains), Tmax=maxtemp)
        return np.array([np.inf if b == 0 else 1.0 / b for b in betas], dtype=np.float64)
    rate = maxtemp / num_chains
    return np.array([1 + i * rate for i in range(num_chains)], dtype=np.float64)
