"""TEST INFRASTRUCTURE ONLY — fp64 CPU restatement of the reference's hot path.

This module is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it. The shipped
engine (``sapt_b200``) never routes through it and has no CPU fallback.

Parity status: PINNED. Every function here is checked by ``tests/test_oracle_golden.py``
against fixtures under ``tests/golden/`` that were produced by running the unmodified
reference classes (``oracle/make_golden.py`` imports them from /root/reference with stubbed
keras/matplotlib), and — when /root/reference is present — against the live reference.
Exception: the surrogate MLP (Keras, third-party, un-pinned version, not installed): its
forward pass is restated from the architecture at _revamp.py:217-233 and is "parity
unpinned" against Keras itself; everything around it (de-normalisation, blend, bookkeeping)
IS pinned because the reference's own ``surrogate.predict`` arithmetic runs in the fixture
generator on top of a fake ``load_model``.

Abbreviations: RV=_revamp.py, RW=misc_codeapril2020/surrogate_rw_common_interval.py,
LG=misc_codeapril2020/surrogate_pt_classifier_langevingrad.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# a1  encode / decode                                                   RV:106-127
# --------------------------------------------------------------------------------------


def num_param(topology: Sequence[int]) -> int:
    d, h, o = topology
    return d * h + h * o + h + o


def decode(w: np.ndarray, topology: Sequence[int]):
    """flat theta -> (W1[D,H], W2[H,O], B1[1,H], B2[1,O]); order W1,W2,B1,B2 (RV:106-116)."""
    d, h, o = topology
    w = np.asarray(w, dtype=np.float64)
    a = d * h
    b = a + h * o
    W1 = w[0:a].reshape(d, h)
    W2 = w[a:b].reshape(h, o)
    B1 = w[b:b + h].reshape(1, h)
    B2 = w[b + h:b + h + o].reshape(1, o)
    return W1, W2, B1, B2


def encode(W1, W2, B1, B2) -> np.ndarray:
    """(RV:120-127) inverse of decode."""
    return np.concatenate([np.ravel(W1), np.ravel(W2), np.ravel(B1), np.ravel(B2)]).astype(np.float64)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))  # RV:62-63


# --------------------------------------------------------------------------------------
# a2-a4  forward pass / evaluate_proposal                               RV:70-76,129-131,155-174
# --------------------------------------------------------------------------------------


def forward(X: np.ndarray, w: np.ndarray, topology: Sequence[int]):
    """Vectorised over rows. z = X.W - B (biases SUBTRACTED), sigmoid twice (RV:70-76).
    Returns (hid[N,H], out[N,O])."""
    W1, W2, B1, B2 = decode(w, topology)
    hid = sigmoid(X @ W1 - B1)
    out = sigmoid(hid @ W2 - B2)
    return hid, out


def evaluate_proposal(data: np.ndarray, w: np.ndarray, topology: Sequence[int]):
    """(RV:155-174) -> fx[N] (argmax class as float, first max on ties), prob[N,O] =
    softmax over the *sigmoid outputs* (RV:129-131)."""
    d = topology[0]
    _, out = forward(np.asarray(data, dtype=np.float64)[:, :d], w, topology)
    fx = np.argmax(out, axis=1).astype(np.float64)
    e = np.exp(out)
    prob = e / np.sum(e, axis=1, keepdims=True)
    return fx, prob


def evaluate_proposal_rowloop(data: np.ndarray, w: np.ndarray, topology: Sequence[int]):
    """Per-row port with the reference's loop structure (one tiny dot per row, RV:165-169).
    Used (a) to show the vectorised form is the same arithmetic and (b) as the CPU
    baseline ("port") in bench.py, because this is how the reference spends its time."""
    d, h, o = topology
    W1, W2, B1, B2 = decode(w, topology)
    n = data.shape[0]
    fx = np.zeros(n)
    prob = np.zeros((n, o))
    for i in range(n):
        x = data[i, 0:d]
        z1 = x.dot(W1) - B1
        hid = 1 / (1 + np.exp(-z1))
        z2 = hid.dot(W2) - B2
        out = 1 / (1 + np.exp(-z2))
        fx[i] = np.argmax(out)
        prob[i] = np.exp(out) / np.sum(np.exp(out))
    return fx, prob


# --------------------------------------------------------------------------------------
# a5-a7  likelihood / prior / metrics                                   RV:380-410
# --------------------------------------------------------------------------------------


def rmse(pred, actual) -> float:
    return float(np.sqrt(((np.asarray(pred) - np.asarray(actual)) ** 2).mean()))  # RV:380-382


def accuracy(pred, actual) -> float:
    pred = np.asarray(pred)
    actual = np.asarray(actual)
    return 100.0 * (int(np.sum(pred == actual)) / pred.shape[0])  # RV:384-389


def likelihood_func(data, w, topology, temperature: float = 1.0, rowloop: bool = False):
    """(RV:391-402) -> [L/T, fx, rmse(fx, y), L] with L = sum_i log prob[i, y_i]."""
    d = topology[0]
    data = np.asarray(data, dtype=np.float64)
    y = data[:, d]
    fx, prob = (evaluate_proposal_rowloop if rowloop else evaluate_proposal)(data, w, topology)
    yi = y.astype(np.int64)
    if rowloop:
        lhood = 0.0
        for i in range(data.shape[0]):  # RV:397-401 (the j loop collapses to the y_i term)
            lhood += np.log(prob[i, yi[i]])
    else:
        lhood = float(np.sum(np.log(prob[np.arange(data.shape[0]), yi])))
    return [lhood / temperature, fx, rmse(fx, y), lhood]


def loglik_stats(data, w, topology):
    """What the device kernel returns for one (replica, data set): (L, n_correct, sum (fx-y)^2)."""
    d = topology[0]
    y = np.asarray(data)[:, d]
    _, fx, _, L = likelihood_func(data, w, topology)
    return L, int(np.sum(fx == y)), float(np.sum((fx - y) ** 2))


def prior_likelihood(sigma_squared: float, nu_1, nu_2, w, topology) -> float:
    """(RV:404-410) -(P/2) ln sigma^2 - sum(w^2)/(2 sigma^2); nu_* unused by the reference."""
    d, h, o = topology
    part1 = -1 * ((d * h + h + o + h * o) / 2) * np.log(sigma_squared)
    part2 = 1 / (2 * sigma_squared) * (np.sum(np.square(np.asarray(w, dtype=np.float64))))
    return float(part1 - part2)


# --------------------------------------------------------------------------------------
# a9  Langevin "gradient" = one epoch of per-row online SGD              LG:96-106,138-156
# --------------------------------------------------------------------------------------


def langevin_gradient(data, w, topology, lrate: float, depth: int = 1) -> np.ndarray:
    """Sequential per-row SGD exactly as LG:138-156: squared-error delta through the output
    sigmoid; hid_delta uses W2 BEFORE its update; biases move by -lr*delta (LG:100-106)."""
    d, h, o = topology
    W1, W2, B1, B2 = (a.copy() for a in decode(w, topology))
    data = np.asarray(data, dtype=np.float64)
    for _ in range(depth):
        for i in range(data.shape[0]):
            x = data[i, 0:d].reshape(1, d)
            y = int(data[i, d])
            hid = sigmoid(x @ W1 - B1)
            out = sigmoid(hid @ W2 - B2)
            desired = np.zeros((1, o))
            desired[0, y] = 1.0
            out_delta = (desired - out) * (out * (1 - out))
            hid_delta = (out_delta @ W2.T) * (hid * (1 - hid))
            W2 += hid.T @ (out_delta * lrate)
            B2 += -1 * lrate * out_delta
            W1 += x.T @ (hid_delta * lrate)
            B1 += -1 * lrate * hid_delta
    return encode(W1, W2, B1, B2)


def langevin_batch_gradient(data, w, topology) -> np.ndarray:
    """Batch restatement at FIXED w of the reference's per-row update direction:
    g = sum_i g_i(w) with g_i the LG:100-106 increments for lr=1 (sign convention: the
    vector to ADD to theta). SURVEY.md H1: this is what the "X^T.delta grouped contraction"
    computes; it equals langevin_gradient only to first order in lr."""
    d, h, o = topology
    W1, W2, B1, B2 = decode(w, topology)
    data = np.asarray(data, dtype=np.float64)
    X = data[:, :d]
    y = data[:, d].astype(np.int64)
    hid = sigmoid(X @ W1 - B1)
    out = sigmoid(hid @ W2 - B2)
    desired = np.zeros_like(out)
    desired[np.arange(X.shape[0]), y] = 1.0
    out_delta = (desired - out) * (out * (1 - out))
    hid_delta = (out_delta @ W2.T) * (hid * (1 - hid))
    gW2 = hid.T @ out_delta
    gB2 = -np.sum(out_delta, axis=0, keepdims=True)
    gW1 = X.T @ hid_delta
    gB1 = -np.sum(hid_delta, axis=0, keepdims=True)
    return encode(gW1, gW2, gB1, gB2)


# --------------------------------------------------------------------------------------
# a15  temperature ladder                                               RV:835-930
# --------------------------------------------------------------------------------------


PTEMCEE_TSTEP = np.array("""
        25.2741 7. 4.47502 3.5236 3.0232 2.71225 2.49879 2.34226 2.22198 2.12628 2.04807 1.98276
        1.92728 1.87946 1.83774 1.80096 1.76826 1.73895 1.7125 1.68849 1.66657 1.64647 1.62795 1.61083
        1.59494 1.58014 1.56632 1.55338 1.54123 1.5298 1.51901 1.50881 1.49916 1.49 1.4813 1.47302
        1.46512 1.45759 1.45039 1.4435 1.4369 1.43056 1.42448 1.41864 1.41302 1.40761 1.40239 1.39736
        1.3925 1.38781 1.38327 1.37888 1.37463 1.37051 1.36652 1.36265 1.35889 1.35524 1.3517 1.34825
        1.3449 1.34164 1.33847 1.33538 1.33236 1.32943 1.32656 1.32377 1.32104 1.31838 1.31578 1.31325
        1.31076 1.30834 1.30596 1.30364 1.30137 1.29915 1.29697 1.29484 1.29275 1.29071 1.2887 1.28673
        1.2848 1.28291 1.28106 1.27923 1.27745 1.27569 1.27397 1.27227 1.27061 1.26898 1.26737 1.26579
        1.26424 1.26271 1.26121 1.25973
""".split(), dtype=np.float64)   # RV:854-874, one entry per ndim = 1..100


def default_beta_ladder(ndim, ntemps=None, Tmax=None) -> np.ndarray:
    """Geometric beta ladder (ptemcee), restated from RV:835-907. Only the branch the
    reference reaches (both ntemps and Tmax given, ndim=2) matters for the engine; the
    argument validation and the Tmax-from-ntemps branch are kept for API parity."""
    if type(ndim) != int or ndim < 1:
        raise ValueError('Invalid number of dimensions specified.')
    if ntemps is None and Tmax is None:
        raise ValueError('Must specify one of ``ntemps`` and ``Tmax``.')
    if Tmax is not None and Tmax <= 1:
        raise ValueError('``Tmax`` must be greater than 1.')
    if ntemps is not None and (type(ntemps) != int or ntemps < 1):
        raise ValueError('Invalid number of temperatures specified.')
    # ptemcee's tuned step per dimension; the large-dimension approximation covers ndim > 100 (RV:876-881)
    tstep = PTEMCEE_TSTEP[ndim - 1] if ndim <= PTEMCEE_TSTEP.shape[0] else 1.0 + 2.0 * np.sqrt(np.log(4.0)) / np.sqrt(ndim)
    append_inf = False
    if Tmax == np.inf:
        append_inf = True
        Tmax = None
        ntemps = ntemps - 1
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
    """(RV:909-930) geometric: T_i = 1/beta_i = maxtemp^(i/(R-1)); linear: 1 + i*maxtemp/R."""
    if geometric:
        betas = default_beta_ladder(2, ntemps=num_chains, Tmax=maxtemp)
        return np.array([np.inf if b == 0 else 1.0 / b for b in betas], dtype=np.float64)
    rate = maxtemp / num_chains
    return np.array([1 + i * rate for i in range(num_chains)], dtype=np.float64)


# --------------------------------------------------------------------------------------
# a14  replica exchange sweep                                           RW:869-898, RW:1095-1099
# --------------------------------------------------------------------------------------


def swap_probability(l1: float, l2: float) -> float:
    return min(1.0, 0.5 * math.exp(min(709.0, l2 - l1)))  # RW:883


def swap_sweep(post_lik: np.ndarray, u: np.ndarray):
    """Sequential sweep k = 0..R-2 over the posted messages. ``post_lik[k]`` is the value
    chain k put at index P+1 of its message (RW:623 untempered / LG:805 tempered).
    Returns (src, swapped): after the sweep chain k holds the message that chain src[k]
    posted (whole message moves, incl. its likelihood), swapped[k] = decision of pair k.
    Because pair k's outputs are put back before pair k+1 is read (RW:1097-1098), a message
    can bubble up the whole ladder in one sweep."""
    R = len(post_lik)
    src = np.arange(R)
    lik = np.array(post_lik, dtype=np.float64).copy()
    swapped = np.zeros(max(R - 1, 0), dtype=bool)
    for k in range(R - 1):
        a = swap_probability(lik[k], lik[k + 1])
        if u[k] < a:
            src[k], src[k + 1] = src[k + 1], src[k]
            lik[k], lik[k + 1] = lik[k + 1], lik[k]
            swapped[k] = True
    return src, swapped


# --------------------------------------------------------------------------------------
# a12  surrogate MLP forward + de-normalisation                         RV:217-233,297-322 ; LG:302-327
# --------------------------------------------------------------------------------------

SURROGATE_HIDDEN = {1: (64, 16), 2: (120, 40), 3: (200, 50)}  # RV:221-229


def surrogate_forward(theta: np.ndarray, weights: Dict[str, np.ndarray]) -> np.ndarray:
    """P -> h1 (relu) -> h2 (relu) -> 1 (sigmoid), Keras Dense convention y = x.W + b,
    computed in float32 like Keras (RV:217-233). theta[n,P] -> s[n] float32."""
    x = np.asarray(theta, dtype=np.float32)
    a1 = np.maximum(x @ weights["W1"].astype(np.float32) + weights["b1"].astype(np.float32), 0)
    a2 = np.maximum(a1 @ weights["W2"].astype(np.float32) + weights["b2"].astype(np.float32), 0)
    z = a2 @ weights["W3"].astype(np.float32) + weights["b3"].astype(np.float32)
    return (1.0 / (1.0 + np.exp(-z.astype(np.float32)))).astype(np.float32).reshape(-1)


def surrogate_denorm(s, min_y: float, max_y: float, mode: str):
    """'minmax': s*(maxY-minY)+minY (RV:320, RW) ; 'half_min': s*minY/2 (LG:325)."""
    # the reference multiplies the float32 Keras output by numpy float64 scalars
    # (self.max_Y[0,0]), so the result is float64 under NEP-50 promotion.
    min_y = np.float64(min_y)
    max_y = np.float64(max_y)
    if mode == "minmax":
        return s * (max_y - min_y) + min_y
    if mode == "half_min":
        return s * min_y / 2
    raise ValueError(mode)


def make_surrogate_weights(P: int, topology_id: int, seed: int, scale: float = 0.05):
    h1, h2 = SURROGATE_HIDDEN[topology_id]
    rs = np.random.RandomState(seed)
    return {
        "W1": (rs.uniform(-scale, scale, (P, h1))).astype(np.float32),
        "b1": (rs.uniform(-scale, scale, (h1,))).astype(np.float32),
        "W2": (rs.uniform(-scale, scale, (h1, h2))).astype(np.float32),
        "b2": (rs.uniform(-scale, scale, (h2,))).astype(np.float32),
        "W3": (rs.uniform(-scale, scale, (h2, 1))).astype(np.float32),
        "b3": (rs.uniform(-scale, scale, (1,))).astype(np.float32),
    }


# --------------------------------------------------------------------------------------
# Philox4x32-10 (Random123; Salmon et al. SC'11) — integer-exact restatement for the
# device RNG; pinned by the Random123 known-answer vectors in tests/test_philox.py.
# --------------------------------------------------------------------------------------

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """counter[...,4] uint32, key[...,2] uint32 -> [...,4] uint32."""
    c = np.array(counter, dtype=np.uint32).copy()
    k = np.array(key, dtype=np.uint32).copy()
    c, k = np.broadcast_arrays(c, k[..., [0, 1, 0, 1]])
    c = c.copy()
    k0 = k[..., 0].copy()
    k1 = k[..., 1].copy()
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c[..., 0].astype(np.uint64) * PHILOX_M0
            p1 = c[..., 2].astype(np.uint64) * PHILOX_M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & mask).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & mask).astype(np.uint32)
            n0 = hi1 ^ c[..., 1] ^ k0
            n1 = lo1
            n2 = hi0 ^ c[..., 3] ^ k1
            n3 = lo0
            c = np.stack([n0, n1, n2, n3], axis=-1)
            k0 = (k0 + PHILOX_W0).astype(np.uint32)
            k1 = (k1 + PHILOX_W1).astype(np.uint32)
    return c


def u32_to_uniform(x: np.ndarray) -> np.ndarray:
    """(x + 0.5) * 2^-32 in fp32 arithmetic? No: the device computes in fp32
    ``fmaf(x, 2^-32, 2^-33)`` which is exact-rounded; restated here identically so the
    comparison is bit-exact: result in (0,1]."""
    return (np.asarray(x, dtype=np.uint32).astype(np.float32) * np.float32(2.0 ** -32)
            + np.float32(2.0 ** -33)).astype(np.float32)


# stream "kinds" — must match enum SaptStream in include/sapt.h
KIND_Z1, KIND_Z2, KIND_LX1, KIND_KU, KIND_LX2, KIND_U, KIND_SWAP = range(7)


def philox_uniform(seed: int, kind: int, step: int, chain: np.ndarray, index: np.ndarray = 0):
    """Uniform for (seed, kind, step, chain[, index]) — lane 0 of the Philox block
    counter=(index>>2... see device: counter = (index/4, step, kind, 0), key = (seed, chain)."""
    chain = np.asarray(chain, dtype=np.uint32)
    index = np.asarray(index, dtype=np.uint32)
    chain, index = np.broadcast_arrays(chain, index)
    ctr = np.stack([index // 4, np.full_like(index, step), np.full_like(index, kind), np.zeros_like(index)], axis=-1)
    key = np.stack([np.full_like(chain, seed & 0xFFFFFFFF), chain], axis=-1)
    blk = philox4x32_10(ctr, key)
    lane = (index % 4).astype(np.int64)
    return u32_to_uniform(np.take_along_axis(blk, lane[..., None], axis=-1)[..., 0])


def philox_normal(seed: int, kind: int, step: int, chain: int, n: int) -> np.ndarray:
    """n standard normals for (seed, kind, step, chain): block b = philox(counter=(b,step,kind,0),
    key=(seed,chain)) gives 4 uniforms -> 2 Box-Muller pairs -> elements 4b..4b+3 in the order
    (r0 cos, r0 sin, r1 cos, r1 sin). fp64 restatement of the fp32 device transform — compare
    with tolerance, the integer stage is bit-exact."""
    nb = (n + 3) // 4
    b = np.arange(nb, dtype=np.uint32)
    ctr = np.stack([b, np.full_like(b, step), np.full_like(b, kind), np.zeros_like(b)], axis=-1)
    key = np.stack([np.full_like(b, seed & 0xFFFFFFFF), np.full_like(b, chain)], axis=-1)
    blk = philox4x32_10(ctr, key)
    u = u32_to_uniform(blk).astype(np.float64)
    r0 = np.sqrt(-2.0 * np.log(u[:, 0]))
    r1 = np.sqrt(-2.0 * np.log(u[:, 2]))
    t0 = 2.0 * np.pi * u[:, 1]
    t1 = 2.0 * np.pi * u[:, 3]
    z = np.stack([r0 * np.cos(t0), r0 * np.sin(t0), r1 * np.cos(t1), r1 * np.sin(t1)], axis=-1).reshape(-1)
    return z[:n]


# --------------------------------------------------------------------------------------
# a8,a10-a13,a16  step oracle: ptReplica.run + coordinator as a pure function of streams
# --------------------------------------------------------------------------------------


@dataclass
class VariantSpec:
    """Behavioural switches that differ between the reference's forks."""
    name: str
    swap: bool                 # replica exchange active
    swap_at_zero: bool         # LG swaps at i % swap_interval == 0 including i == 0 (LG:804)
    post_tempered: bool        # message carries tempered L (LG:805) or L*T (RW:623)
    take_post_lik: bool        # LG: `likelihood = result[P+1]` after the swap (LG:813-815)
    gate_on_swap_interval: bool  # RW gates surrogate use on swap_interval (RW:507)
    common_interval: bool      # RW: swap + surrogate rendezvous share i % swap_interval == 0, i != 0
    denorm: str                # 'minmax' (RV:320) | 'half_min' (LG:325)
    compare_surrogate: bool    # RW:371/RV:370 True ; LG:381 False
    double_eval: bool          # LG evaluates the true likelihood twice per true step (LG:622,676)


VARIANTS = {
    "rw": VariantSpec("rw", True, False, False, False, True, True, "minmax", True, False),
    "lg": VariantSpec("lg", True, True, True, True, False, False, "half_min", False, True),
    "rv": VariantSpec("rv", False, False, False, False, False, False, "minmax", True, False),
}

STEP_W = 0.025        # RV:449 (std of the random-walk proposal)
SIGMA_SQUARED = 25.0  # RV:456
L_PROB = 0.5          # LG:385


@dataclass
class Streams:
    """Injected random streams, indexed [step, chain] (+[P] for normals).
    z1  : first proposal noise (np.random.normal, LG:551)
    z2  : second draw — surrogate-branch redraw (LG:560) or Langevin proposal noise (LG:651)
    lx2 : np.random.uniform before the Langevin coin (LG:647)
    ku  : python random.uniform surrogate coin (LG:553)
    u   : python random.uniform MH coin (LG:704)
    swap_u : [sweep, R-1] coordinator np.random.uniform (RW:888)"""
    z1: np.ndarray
    z2: Optional[np.ndarray]
    lx2: Optional[np.ndarray]
    ku: np.ndarray
    u: np.ndarray
    swap_u: Optional[np.ndarray]


@dataclass
class SurrogateState:
    """What the workers know about the global surrogate after a rendezvous."""
    weights: Optional[Dict[str, np.ndarray]] = None
    min_y: float = 0.0
    max_y: float = 0.0
    signature: int = 0


def mh_probability(diff_likelihood: float, diff_prior: float) -> float:
    try:
        return min(1.0, math.exp(diff_likelihood + diff_prior))  # RV:571-577
    except OverflowError:
        return 1.0


def run_pt(variant: str, traindata, testdata, topology, temperatures, w0, samples: int,
           swap_interval: int, surrogate_interval: int, surrogate_prob: float,
           streams: Streams, use_langevin: bool = False, learn_rate: float = 0.5,
           surrogate_weights_per_signature: Optional[List[Dict[str, np.ndarray]]] = None,
           use_surrogate: bool = True):
    """Restatement of R x ptReplica.run (RW:413-704 / LG:428-908 / RV:412-780) in lock-step
    with the coordinator's barrier (RW:1068-1129), as a pure function of the injected
    streams. ``samples`` is per chain (= int(NumSample/num_chains), RV:800).

    ``surrogate_weights_per_signature[s-1]`` = MLP weights the trainer produced at model
    signature s (training itself is off the hot path; the fixture generator fakes it).

    Returns a dict of per-chain traces with the reference's array names."""
    spec = VARIANTS[variant]
    traindata = np.asarray(traindata, dtype=np.float64)
    testdata = np.asarray(testdata, dtype=np.float64)
    R = len(temperatures)
    P = num_param(topology)
    d = topology[0]
    y_train = traindata[:, d]
    y_test = testdata[:, d]
    T = np.asarray(temperatures, dtype=np.float64)

    w = np.tile(np.asarray(w0, dtype=np.float64), (R, 1))  # same start for all chains, RV:940-943
    pos_w = np.ones((R, samples, P))
    lhood_list = np.zeros((R, samples))
    rmse_train = np.zeros((R, samples))
    rmse_test = np.zeros((R, samples))
    acc_train = np.zeros((R, samples))
    acc_test = np.zeros((R, samples))
    prop_list = np.zeros((R, samples, P))
    likeh_list = np.zeros((R, samples, 2))
    likeh_list[:, 0, :] = -100
    surg = np.zeros((R, samples, 3))
    accept_list = np.zeros((R, samples))
    is_surrogate_step = np.zeros((R, samples), dtype=bool)
    accepted = np.zeros((R, samples), dtype=bool)
    langevin_step = np.zeros((R, samples), dtype=bool)

    naccept = np.zeros(R, dtype=np.int64)
    likelihood = np.zeros(R)       # the `likelihood` variable (what LG posts)
    likelihood_copy = np.zeros(R)  # what MH reads (RV:586-587)
    prior_current = np.zeros(R)
    rmsetrain = np.zeros(R)
    rmsetest = np.zeros(R)
    pred_train = [None] * R
    pred_test = [None] * R
    likl_without_temp = np.zeros(R)
    for r in range(R):
        prior_current[r] = prior_likelihood(SIGMA_SQUARED, 0, 0, w[r], topology)
        lik, pred_train[r], rmsetrain[r], _ = likelihood_func(traindata, w[r], topology, T[r])
        _, pred_test[r], rmsetest[r], lwt = likelihood_func(testdata, w[r], topology, T[r])
        likelihood[r] = lik
        likelihood_copy[r] = lik
        likl_without_temp[r] = lwt  # the TEST-set value survives initialisation (RW:466-467)

    surrogate = SurrogateState()
    surr_train_sets: List[List[np.ndarray]] = [[] for _ in range(R)]
    surr_batches: List[np.ndarray] = []  # what the trainer received at each rendezvous
    swap_decisions: List[np.ndarray] = []
    swap_sources: List[np.ndarray] = []
    num_swap = 0
    total_swap_proposals = 0
    sweep = 0
    gate = swap_interval if spec.gate_on_swap_interval else surrogate_interval

    for i in range(samples - 1):
        for r in range(R):
            w_proposal = w[r] + STEP_W * streams.z1[i, r]
            ku = streams.ku[i, r]
            if use_surrogate and ku < surrogate_prob and i >= gate + 1:
                is_true = False
                if spec.name == "lg":
                    w_proposal = w[r] + STEP_W * streams.z2[i, r]  # redraw, LG:560
                s = surrogate_forward(w_proposal[None, :], surrogate.weights)
                # float32 prediction array * float64 scalars, as in the reference (RV:319-320)
                pred = surrogate_denorm(s, surrogate.min_y, surrogate.max_y, spec.denorm)
                surrogate_likelihood = pred * (1.0 / T[r])
                mov = (surg[r, i, 2] + surg[r, i - 1, 2] + surg[r, i - 2, 2]) / 3
                likelihood_proposal = (surrogate_likelihood[0] * 0.5) + (mov * 0.5)
                if spec.compare_surrogate:
                    lt, pred_train[r], rmsetrain[r], likl_without_temp[r] = likelihood_func(
                        traindata, w_proposal, topology, T[r])
                else:
                    lt = 0
                surg[r, i + 1, 0] = lt
                surg[r, i + 1, 1] = likelihood_proposal
                surg[r, i + 1, 2] = mov
            else:
                is_true = True
                surg[r, i + 1, 1] = np.nan
                likelihood_proposal, pred_train[r], rmsetrain[r], likl_without_temp[r] = likelihood_func(
                    traindata, w_proposal, topology, T[r])
                _, pred_test[r], rmsetest[r], _ = likelihood_func(testdata, w_proposal, topology, T[r])
                surr_train_sets[r].append(np.concatenate([w_proposal, [likl_without_temp[r]]]))
                surg[r, i + 1, 0] = likelihood_proposal
                surg[r, i + 1, 2] = likelihood_proposal
                if spec.double_eval:
                    lx = streams.lx2[i, r]
                    if use_langevin and lx < L_PROB:
                        langevin_step[r, i + 1] = True
                        w_gd = langevin_gradient(traindata, w[r].copy(), topology, learn_rate, 1)
                        w_proposal = w_gd + STEP_W * streams.z2[i, r]
                        # w_prop_gd / diff_prop are computed by the reference (LG:652-668) but
                        # never enter the MH ratio (LG:687-691); omitted here on purpose.
                    likelihood_proposal, pred_train[r], rmsetrain[r], likl_without_temp[r] = likelihood_func(
                        traindata, w_proposal, topology, T[r])
                    _, pred_test[r], rmsetest[r], _ = likelihood_func(testdata, w_proposal, topology, T[r])
                    surg[r, i + 1, 1] = np.nan
                    surg[r, i + 1, 0] = likelihood_proposal
            is_surrogate_step[r, i + 1] = not is_true
            prior_prop = prior_likelihood(SIGMA_SQUARED, 0, 0, w_proposal, topology)
            mh = mh_probability(likelihood_proposal - likelihood_copy[r], prior_prop - prior_current[r])
            accept_list[r, i + 1] = naccept[r]
            u = streams.u[i, r]
            prop_list[r, i + 1] = w_proposal
            likeh_list[r, i + 1, 0] = likl_without_temp[r]
            if u < mh:
                accepted[r, i + 1] = True
                naccept[r] += 1
                likelihood[r] = likelihood_proposal
                likelihood_copy[r] = likelihood_proposal
                prior_current[r] = prior_prop
                w[r] = w_proposal
                pos_w[r, i + 1] = w_proposal
                if is_true:
                    lhood_list[r, i + 1] = likelihood[r] * T[r]
                    rmse_train[r, i + 1] = rmsetrain[r]
                    rmse_test[r, i + 1] = rmsetest[r]
                    acc_train[r, i + 1] = accuracy(pred_train[r], y_train)
                    acc_test[r, i + 1] = accuracy(pred_test[r], y_test)
                else:
                    lhood_list[r, i + 1] = np.inf
                    rmse_train[r, i + 1] = rmse_test[r, i + 1] = np.inf
                    acc_train[r, i + 1] = acc_test[r, i + 1] = np.inf
            else:
                pos_w[r, i + 1] = pos_w[r, i]
                if is_true:
                    lhood_list[r, i + 1] = likelihood_proposal * T[r]
                    rmse_train[r, i + 1] = rmse_train[r, i]
                    rmse_test[r, i + 1] = rmse_test[r, i]
                    acc_train[r, i + 1] = acc_train[r, i]
                    acc_test[r, i + 1] = acc_test[r, i]
                else:
                    lhood_list[r, i + 1] = np.inf
                    rmse_train[r, i + 1] = rmse_test[r, i + 1] = np.inf
                    acc_train[r, i + 1] = acc_test[r, i + 1] = np.inf

        # ---- rendezvous (barrier over all chains) ----
        do_swap = spec.swap and (i % swap_interval == 0) and (spec.swap_at_zero or i != 0)
        if spec.common_interval:
            do_train = do_swap
        else:
            do_train = use_surrogate and (i % surrogate_interval == 0) and (i != 0)
        if do_swap and R > 1:
            post = likelihood.copy() if spec.post_tempered else likelihood * T
            src, sw = swap_sweep(post, streams.swap_u[sweep])
            sweep += 1
            num_swap += int(sw.sum())
            total_swap_proposals += R - 1
            w = w[src].copy()
            if spec.take_post_lik:
                likelihood = post[src].copy()
            swap_decisions.append(sw)
            swap_sources.append(src)
        if do_train:
            rows = [np.asarray(surr_train_sets[r]).reshape(-1, P + 1) for r in range(R)]
            if spec.name in ("rw", "lg"):
                # RW/LG ship surr_train_set INCLUDING its all-zero first row (RW:496,626; LG:543,822)
                rows = [np.vstack([np.zeros((1, P + 1)), x]) for x in rows]
            else:
                # RV both vstacks and overwrites row count_real, then ships surr_train_set[1:count_real]
                # (RV:560-567,683): the interval's FIRST true proposal of every chain is dropped
                rows = [x[1:] for x in rows]
            batch = np.concatenate(rows, axis=0)
            surr_batches.append(batch)
            if batch.shape[0] > 0:
                Y = batch[:, P]
                surrogate.signature += 1
                if surrogate.signature == 1:  # minmax.txt written once (RV:996-997, RV:707-710)
                    surrogate.min_y = float(np.min(Y) * 2)
                    surrogate.max_y = -1.0
                if surrogate_weights_per_signature is not None:
                    surrogate.weights = surrogate_weights_per_signature[surrogate.signature - 1]
            surr_train_sets = [[] for _ in range(R)]

    return dict(pos_w=pos_w, lhood_list=lhood_list, rmse_train=rmse_train, rmse_test=rmse_test,
                acc_train=acc_train, acc_test=acc_test, prop_list=prop_list, likeh_list=likeh_list,
                surg_likeh_list=surg, accept_list=accept_list, naccept=naccept, w_final=w,
                is_surrogate_step=is_surrogate_step, accepted=accepted, langevin_step=langevin_step,
                swap_decisions=swap_decisions, swap_sources=swap_sources, num_swap=num_swap,
                total_swap_proposals=total_swap_proposals, surr_batches=surr_batches,
                likelihood=likelihood, likelihood_copy=likelihood_copy, prior_current=prior_current,
                min_y=surrogate.min_y, max_y=surrogate.max_y, signature=surrogate.signature)


# --------------------------------------------------------------------------------------
# N4  local-surrogate fork (LS = localsurrogates/surrogate_pt_classifier_rw_localsurrogate.py)
# --------------------------------------------------------------------------------------
# Every replica owns a 4 -> 10 -> 1 sigmoid MLP that predicts the next tempered likelihood from the last four
# entries of surg_likeh_list[:, 2] (lag-2 windows of five, "Takens embedding"), retrained online for 200 epochs
# at every surrogate interval. Quirks kept on purpose (they define the reference's traces):
#   * BackwardPass one-hot encodes int(desired) (LS:90-93): the regression target y/limit lies in [0, 0.5], so the
#     network is always trained towards 1.0 and a target >= 1 would raise IndexError in the reference;
#   * limit = 2 * min(y) is fixed at the first training (LS:791) and later windows are scaled by the same value;
#   * the window fed to predict() is rows i-4 .. i-1, i.e. it skips the newest entry (LS:592);
#   * the prediction is used as the tempered likelihood directly: no 1/T, no moving-average blend (LS:598-602);
#   * an ACCEPTED surrogate proposal leaves `likelihood` a 1-element array and the reference then dies in
#     np.concatenate at the next exchange (LS:757); the restatement keeps the scalar and goes on.

LS_BLOCK = 5        # LS:548
LS_LAG = 2          # LS:779
LS_HIDDEN = 10      # LS:243
LS_LRATE = 0.1      # LS:262
LS_EPOCHS = 200     # LS:253


def takens_windows(y: np.ndarray, blocksize: int = LS_BLOCK, lag: int = LS_LAG):
    """LS:773-787. Returns (windows [nb, blocksize], the trimmed series)."""
    y = np.asarray(y, dtype=np.float64)
    n = y.shape[0] % blocksize
    y = y[n:]
    nb = int((y.shape[0] - blocksize) / lag)
    v = np.zeros((nb, blocksize))
    for b in range(nb):
        v[b, :] = y[b * lag: b * lag + blocksize]
    return v, y


def local_net_init(z: np.ndarray, n_in: int = LS_BLOCK - 1, hidden: int = LS_HIDDEN) -> Dict[str, np.ndarray]:
    """create_network (LS:57-70) from n_in*hidden + hidden + hidden + 1 standard normals in the order the
    reference draws them: W1, B1, W2, B2."""
    z = np.asarray(z, dtype=np.float64).reshape(-1)
    a, b, c = n_in * hidden, n_in * hidden + hidden, n_in * hidden + 2 * hidden
    return dict(W1=z[:a].reshape(n_in, hidden) / np.sqrt(n_in), B1=z[a:b].reshape(1, hidden) / np.sqrt(hidden),
                W2=z[b:c].reshape(hidden, 1) / np.sqrt(hidden), B2=z[c:c + 1].reshape(1, 1) / np.sqrt(hidden))


def local_net_forward(x: np.ndarray, net: Dict[str, np.ndarray]):
    """ForwardPass (LS:80-87) -> (hidout [1,h], out [1,1])."""
    hid = sigmoid(np.asarray(x, dtype=np.float64).reshape(1, -1) @ net["W1"] - net["B1"])
    out = sigmoid(hid @ net["W2"] - net["B2"])
    return hid, out


def local_net_train(data: np.ndarray, net: Dict[str, np.ndarray], lrate: float = LS_LRATE, epochs: int = LS_EPOCHS) -> float:
    """backprop_train (LS:152-177) + BackwardPass (LS:90-101), in place. The desired output is the one-hot of
    int(target) over ONE output neuron, i.e. 1.0 (IndexError in the reference if target >= 1)."""
    data = np.asarray(data, dtype=np.float64)
    n_in = net["W1"].shape[0]
    sse = 0.0
    for _ in range(epochs):
        for row in data:
            x = row[:n_in].reshape(1, n_in)
            hid, out = local_net_forward(x, net)
            out_delta = (1.0 - out) * (out * (1 - out))
            hid_delta = (out_delta @ net["W2"].T) * (hid * (1 - hid))
            net["W2"] += hid.T @ (out_delta * lrate)
            net["B2"] += -1 * lrate * out_delta
            net["W1"] += x.T @ (hid_delta * lrate)
            net["B1"] += -1 * lrate * hid_delta
            sse += float(np.sum(np.square(out - row[n_in:])) / 1.0)   # sampleEr against the raw target (LS:75-78)
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.sqrt(np.float64(sse) / len(data) * 1))


def run_pt_local(traindata, testdata, topology, temperatures, w0, samples: int, swap_interval: int,
                 surrogate_interval: int, surrogate_prob: float, streams: Streams, net_init: np.ndarray):
    """R x ptReplica.run of the local-surrogate fork (LS:447-836) in lock-step with a barrier coordinator that
    sweeps swap_procedure (LS:1056-1084) over k = 0..R-2 at every i % swap_interval == 0 (the fork's own
    run_chains is racy in the same way as LG's, LS:1108). `net_init[r]` = the 61 standard normals chain r draws
    in create_network at its first training."""
    traindata = np.asarray(traindata, dtype=np.float64)
    testdata = np.asarray(testdata, dtype=np.float64)
    R = len(temperatures)
    P = num_param(topology)
    d = topology[0]
    y_train = traindata[:, d]
    y_test = testdata[:, d]
    T = np.asarray(temperatures, dtype=np.float64)
    w = np.tile(np.asarray(w0, dtype=np.float64), (R, 1))
    lhood_list = np.zeros((R, samples))
    rmse_train = np.zeros((R, samples)); rmse_test = np.zeros((R, samples))
    acc_train = np.zeros((R, samples)); acc_test = np.zeros((R, samples))
    likeh_list = np.zeros((R, samples, 2)); likeh_list[:, 0, :] = -100
    surg = np.zeros((R, samples, 3))
    accept_list = np.zeros((R, samples))
    is_surrogate_step = np.zeros((R, samples), dtype=bool)
    accepted = np.zeros((R, samples), dtype=bool)
    naccept = np.zeros(R, dtype=np.int64)
    likelihood = np.zeros(R); likelihood_copy = np.zeros(R); prior_current = np.zeros(R)
    rmsetrain = np.zeros(R); rmsetest = np.zeros(R)
    pred_train = [None] * R; pred_test = [None] * R
    likl_without_temp = np.zeros(R)
    for r in range(R):
        prior_current[r] = prior_likelihood(SIGMA_SQUARED, 0, 0, w[r], topology)
        lik, pred_train[r], rmsetrain[r], _ = likelihood_func(traindata, w[r], topology, T[r])
        _, pred_test[r], rmsetest[r], lwt = likelihood_func(testdata, w[r], topology, T[r])
        likelihood[r] = lik; likelihood_copy[r] = lik; likl_without_temp[r] = lwt   # LS:500-501
    nets: List[Optional[Dict[str, np.ndarray]]] = [None] * R
    limit = np.zeros(R)
    train_rmse: List[np.ndarray] = []
    swap_decisions: List[np.ndarray] = []
    num_swap = 0; total_swap_proposals = 0; sweep = 0
    for i in range(samples - 1):
        for r in range(R):
            w_proposal = w[r] + STEP_W * streams.z1[i, r]
            ku = streams.ku[i, r]
            if ku < surrogate_prob and i >= surrogate_interval + 1:                      # LS:588
                is_true = False
                x = surg[r, i - (LS_BLOCK - 1):i, 2] / limit[r]                           # LS:592
                _, out = local_net_forward(x, nets[r])
                likelihood_proposal = float(out[0, 0] * limit[r])                          # LS:598
                lt, pred_train[r], rmsetrain[r], likl_without_temp[r] = likelihood_func(traindata, w_proposal, topology, T[r])
                surg[r, i + 1, 0] = lt
                surg[r, i + 1, 1] = likelihood_proposal
                surg[r, i + 1, 2] = np.mean(x) * limit[r]                                 # LS:617
            else:
                is_true = True
                surg[r, i + 1, 1] = np.nan
                likelihood_proposal, pred_train[r], rmsetrain[r], likl_without_temp[r] = likelihood_func(
                    traindata, w_proposal, topology, T[r])
                _, pred_test[r], rmsetest[r], _ = likelihood_func(testdata, w_proposal, topology, T[r])   # LS:630
                surg[r, i + 1, 0] = likelihood_proposal
                surg[r, i + 1, 2] = likelihood_proposal
            is_surrogate_step[r, i + 1] = not is_true
            prior_prop = prior_likelihood(SIGMA_SQUARED, 0, 0, w_proposal, topology)
            mh = mh_probability(likelihood_proposal - likelihood_copy[r], prior_prop - prior_current[r])
            accept_list[r, i + 1] = naccept[r]
            likeh_list[r, i + 1, 0] = likl_without_temp[r]
            if streams.u[i, r] < mh:
                accepted[r, i + 1] = True
                naccept[r] += 1
                likelihood[r] = likelihood_proposal; likelihood_copy[r] = likelihood_proposal
                prior_current[r] = prior_prop
                w[r] = w_proposal
                if is_true:
                    lhood_list[r, i + 1] = likelihood[r] * T[r]
                    rmse_train[r, i + 1] = rmsetrain[r]; rmse_test[r, i + 1] = rmsetest[r]
                    acc_train[r, i + 1] = accuracy(pred_train[r], y_train)
                    acc_test[r, i + 1] = accuracy(pred_test[r], y_test)
                else:
                    lhood_list[r, i + 1] = rmse_train[r, i + 1] = rmse_test[r, i + 1] = np.inf
                    acc_train[r, i + 1] = acc_test[r, i + 1] = np.inf
            else:
                if is_true:
                    lhood_list[r, i + 1] = likelihood_proposal * T[r]
                    rmse_train[r, i + 1] = rmse_train[r, i]; rmse_test[r, i + 1] = rmse_test[r, i]
                    acc_train[r, i + 1] = acc_train[r, i]; acc_test[r, i + 1] = acc_test[r, i]
                else:
                    lhood_list[r, i + 1] = rmse_train[r, i + 1] = rmse_test[r, i + 1] = np.inf
                    acc_train[r, i + 1] = acc_test[r, i + 1] = np.inf
        if i % swap_interval == 0 and R > 1:                                              # LS:756-768: tempered L posted
            src, sw = swap_sweep(likelihood.copy(), streams.swap_u[sweep])
            sweep += 1
            num_swap += int(sw.sum()); total_swap_proposals += R - 1
            w = w[src].copy()                                                             # only w is taken back (LS:765)
            swap_decisions.append(sw)
        if i % surrogate_interval == 0 and i != 0:                                        # LS:771
            rm = np.zeros(R)
            for r in range(R):
                if nets[r] is None:
                    v, ytrim = takens_windows(surg[r, 0:i, 2])
                    limit[r] = np.amin(ytrim) * 2                                         # LS:791
                    nets[r] = local_net_init(net_init[r])
                else:
                    v, _ = takens_windows(surg[r, i - surrogate_interval:i, 2])
                rm[r] = local_net_train(v / limit[r], nets[r])
            train_rmse.append(rm)
    return dict(lhood_list=lhood_list, rmse_train=rmse_train, rmse_test=rmse_test, acc_train=acc_train, acc_test=acc_test,
                likeh_list=likeh_list, surg_likeh_list=surg, accept_list=accept_list, naccept=naccept, w_final=w,
                is_surrogate_step=is_surrogate_step, accepted=accepted, swap_decisions=swap_decisions, num_swap=num_swap,
                total_swap_proposals=total_swap_proposals, likelihood=likelihood, likelihood_copy=likelihood_copy,
                prior_current=prior_current, limit=limit, nets=nets, train_rmse=train_rmse)


# --------------------------------------------------------------------------------------
# a14 (device algorithm): the exchange sweep WITHOUT a serial walk, restated in numpy so the algorithm behind
# swap_sweep_parallel_kernel (csrc/sapt_mcmc.cu) can be checked on the CPU against swap_sweep above.
# --------------------------------------------------------------------------------------

def _order_key(x: np.ndarray) -> np.ndarray:
    """Order-preserving int64 image of float64 (non-NaN)."""
    b = np.asarray(x, dtype=np.float64).view(np.int64)
    return b ^ ((b >> 63) & np.int64(0x7FFFFFFFFFFFFFFF))


def swap_sweep_parallel(post_lik: np.ndarray, u: np.ndarray):
    """Same result as swap_sweep, computed with data-parallel steps only:
    (1) u < min(1, 0.5 exp(min(709, p - c)))  <=>  u < 1 and c < p - ln(2u): thresholds t[j] = p[j+1] - ln(2 u[j]);
    (2) from the state "reset at m" (pair m next, carrying chain m's own message) every pair swaps until the first
        j >= m with t[j] <= p[m]: nxt[m] = that j + 1, found for all m with a range-minimum sparse table;
    (3) the states the sweep visits are the list 0 -> nxt[0] -> ..., marked by pointer doubling;
    (4) pair j did not swap <=> state j+1 is on the list; position k keeps chain k+1's message if pair k swapped,
        else the message carried then = the last list state <= k (a prefix maximum).
    Requires u > 0 (the device falls back to the serial walk otherwise). NaN: like Python's min(1, nan) == 1."""
    p = np.asarray(post_lik, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    R = len(p)
    NP = R - 1
    if NP <= 0:
        return np.arange(R), np.zeros(0, dtype=bool)
    assert np.all(u > 0)
    KMIN, KMAX = np.iinfo(np.int64).min, np.iinfo(np.int64).max
    kp = np.where(np.isnan(p), KMIN, _order_key(np.nan_to_num(p, nan=0.0)))
    with np.errstate(invalid="ignore", divide="ignore"):
        t = p[1:] - np.log(2.0 * u)
    kt = np.where(np.isnan(t), KMAX, _order_key(np.nan_to_num(t, nan=0.0)))
    kt = np.where(u < 1.0, kt, KMIN)
    levels = 1
    while (1 << levels) <= NP:
        levels += 1
    st = [kt]
    for l in range(1, levels):
        a = st[-1]
        half, span = 1 << (l - 1), 1 << l
        b = np.full(NP, KMAX, dtype=np.int64)
        n = NP - span + 1
        if n > 0:
            b[:n] = np.minimum(a[:n], a[half:half + n])
        st.append(b)
    j = np.arange(NP)
    c = kp[:NP]
    for l in range(levels - 1, -1, -1):
        ok = (j + (1 << l) <= NP)
        go = ok & (st[l][np.minimum(j, NP - 1)] > c)
        j = np.where(go, j + (1 << l), j)
    nxt = np.full(R + 1, R, dtype=np.int64)
    nxt[:NP] = np.where(j < NP, j + 1, R)
    mark = np.zeros(R + 1, dtype=bool)
    mark[0] = True
    J = nxt.copy()
    for _ in range(levels + 1):
        tgt = J[:R][mark[:R]]
        mark[tgt[tgt < R]] = True
        J = np.where(J < R, J[np.minimum(J, R)], R)
    swapped = ~mark[1:R]
    last = np.maximum.accumulate(np.where(mark[:R], np.arange(R), -1))
    src = np.empty(R, dtype=np.int64)
    src[:NP] = np.where(swapped, np.arange(1, R), last[:NP])
    src[NP] = last[NP]
    return src, swapped
