"""Helpers shared by the oracle-pinning tests and the GPU parity tests."""
import os

import numpy as np

from oracle import sapt_oracle as so

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RUNS = {"run_rw_cancer": "rw", "run_rw_nosurr": "rw", "run_lg_cancer": "lg", "run_lg_ions": "lg", "run_rv_cancer": "rv"}


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def run_inputs(g):
    """(kwargs for oracle.run_pt / the device engine) from a run_* fixture."""
    nsig = int(g["n_signatures"])
    sw = [{k: g[f"sur{s}_{k}"] for k in ("W1", "b1", "W2", "b2", "W3", "b3")} for s in range(nsig)]
    streams = so.Streams(z1=g["z1"], z2=g["z2"], lx2=g["lx2"], ku=g["ku"], u=g["u"], swap_u=g["swap_u"])
    return dict(traindata=g["train"], testdata=g["test"], topology=[int(v) for v in g["topology"]],
                temperatures=g["temperatures"], w0=g["w0"], samples=int(g["samples"]),
                swap_interval=int(g["swap_interval"]), surrogate_interval=int(g["surrogate_interval"]),
                surrogate_prob=float(g["surrogate_prob"]), streams=streams,
                use_langevin=bool(g["use_langevin"]), learn_rate=float(g["learn_rate"]),
                surrogate_weights_per_signature=sw)


TRACE_KEYS = ("pos_w", "rmse_train", "rmse_test", "acc_train", "acc_test", "surg_likeh_list", "likeh_list", "accept_list")


def compare_traces(g, out, rtol, atol, keys=TRACE_KEYS):
    """Assert `out` (oracle/device traces) reproduces the reference traces in fixture g."""
    for k in keys:
        a = g["trace_" + k]
        b = np.asarray(out[k])
        if k == "likeh_list":
            a, b = a[:, :, 0], b[:, :, 0]
        assert a.shape == b.shape, (k, a.shape, b.shape)
        assert np.array_equal(np.isnan(a), np.isnan(b)), k
        assert np.array_equal(np.isinf(a), np.isinf(b)), k
        m = np.isfinite(a)
        np.testing.assert_allclose(b[m], a[m], rtol=rtol, atol=atol, err_msg=k)


LS_TRACE_KEYS = ("rmse_train", "rmse_test", "acc_train", "acc_test", "surg_likeh_list", "likeh_list", "accept_list")


def run_local_inputs(g):
    """kwargs for oracle.run_pt_local / the device engine (variant 'ls') from a run_ls_* fixture."""
    streams = so.Streams(z1=g["z1"], z2=None, lx2=None, ku=g["ku"], u=g["u"], swap_u=g["swap_u"])
    return dict(traindata=g["train"], testdata=g["test"], topology=[int(v) for v in g["topology"]],
                temperatures=g["temperatures"], w0=g["w0"], samples=int(g["samples"]),
                swap_interval=int(g["swap_interval"]), surrogate_interval=int(g["surrogate_interval"]),
                surrogate_prob=float(g["surrogate_prob"]), streams=streams, net_init=g["net_init"])
