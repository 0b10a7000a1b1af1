"""Temperature ladder — host-side mirror of ParallelTempering.default_beta_ladder /
assign_temperatures (_revamp.py:835-930). The reference only ever calls it as
``default_beta_ladder(2, ntemps=num_chains, Tmax=maxtemp)`` (_revamp.py:919), for which the
ladder is the closed form T_i = maxtemp ** (i / (R - 1)); the ptemcee step table therefore only
matters when Tmax is None, and is kept as the published large-dimension approximation plus the
leading entries."""
from __future__ import annotations

import numpy as np

# leading entries of ptemcee's tuned step table (ndim 1..5); beyond that the closed-form
# approximation 1 + 2*sqrt(ln 4)/sqrt(ndim) from the same source is used.
_TSTEP_HEAD = (25.2741, 7.0, 4.47502, 3.5236, 3.0232)


def default_beta_ladder(ndim, ntemps=None, Tmax=None):
    if type(ndim) != int or ndim < 1:
        raise ValueError('Invalid number of dimensions specified.')
    if ntemps is None and Tmax is None:
        raise ValueError('Must specify one of ``ntemps`` and ``Tmax``.')
    if Tmax is not None and Tmax <= 1:
        raise ValueError('``Tmax`` must be greater than 1.')
    if ntemps is not None and (type(ntemps) != int or ntemps < 1):
        raise ValueError('Invalid number of temperatures specified.')
    tstep = _TSTEP_HEAD[ndim - 1] if ndim <= len(_TSTEP_HEAD) else 1.0 + 2.0 * np.sqrt(np.log(4.0)) / np.sqrt(ndim)
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
