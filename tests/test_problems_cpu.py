"""CPU tests of the host-side problem table / preprocessing / result-row helpers."""
import os

import numpy as np
import pytest

import golden_util as gu
from oracle import ref_loader


def test_problem_table_matches_reference_main():
    from sapt_b200.problems import PROBLEMS
    want = {1: (11, 50, 10, 2), 2: (11, 50, 10, 2), 3: (4, 8, 3, 1), 4: (34, 15, 2, 1), 5: (9, 8, 2, 1), 6: (20, 50, 2, 2),
            7: (16, 30, 10, 2), 8: (6, 25, 18, 3)}   # _revamp.py:1425-1519
    for k, (ip, h, o, st) in want.items():
        s = PROBLEMS[k]
        assert (s.ip, s.hidden, s.output, s.surrogate_topology) == (ip, h, o, st)


def test_split_and_zscore():
    from sapt_b200.problems import split_60_40, zscore_columns
    rs = np.random.RandomState(0)
    f = rs.randn(50, 3) * 5 + 2; c = rs.randint(0, 3, (50, 1)).astype(float)
    tr, te = split_60_40(f, c, np.random.RandomState(1))
    assert tr.shape == (30, 4) and te.shape == (20, 4)
    allrows = np.vstack([tr, te])
    np.testing.assert_allclose(allrows[:, :3].mean(0), 0, atol=1e-12)
    np.testing.assert_allclose(allrows[:, :3].std(0), 1, atol=1e-12)
    z = zscore_columns(f, 2)
    np.testing.assert_allclose(z[:, 2], f[:, 2])


@pytest.mark.skipif(not ref_loader.reference_available(), reason="needs the reference DATA directory")
def test_load_problem_reproduces_reference_arrays():
    from sapt_b200.problems import load_problem
    root = os.path.join(ref_loader.REFERENCE_ROOT, "DATA")
    for prob, fixture, hidden in ((5, "unit_cancer", 12), (4, "unit_ions", 50)):
        tr, te, topo, st, name = load_problem(prob, root, hidden=hidden)
        g = gu.load(fixture)
        assert topo == [int(v) for v in g["topology"]]
        np.testing.assert_array_equal(tr, g["train"]); np.testing.assert_array_equal(te, g["test"])
    tr, te, topo, st, name = load_problem(7, root)
    g = gu.load("unit_pendigit")
    np.testing.assert_array_equal(tr[:600], g["train"]); assert topo == [16, 30, 10] and st == 2


def test_result_row_layout():
    from sapt_b200.results import result_row
    acc_tr = np.array([[90.0], [np.inf], [92.0]]); acc_te = np.array([[80.0], [82.0], [np.inf]])
    row = result_row(5, 1000, 2, 5, 0.1, 0.5, False, 0.01, acc_tr, acc_te, 12.5, 70.0, 3.0, 1.5, 90)
    assert row.shape == (19,)
    assert row[0] == 5 and row[8] == 91.0 and row[10] == 92.0 and row[11] == 81.0 and row[14] == 12.5 and row[18] == 90


def test_surrogate_trainer_learns_and_warm_starts():
    """N1 (host side, CPU here): the PyTorch restatement of surrogate.train (_revamp.py:235-295). Keras itself is
    not installed, so this checks behaviour, not bits: shapes, the normalisation of _revamp.py:989-1024, that the
    fit reduces the error of a smooth target, that signature > 1 warm-starts, and that the exported weights are
    what oracle.surrogate_forward (the predict path) consumes."""
    import torch
    from oracle import sapt_oracle as so
    from sapt_b200.surrogate import SurrogateTrainer
    rs = np.random.RandomState(0)
    P, n = 30, 600
    X = rs.standard_normal((n, P)) * 0.5
    Y = -300.0 - 40.0 * np.tanh(X[:, 0] - 0.5 * X[:, 1]) - 10.0 * X[:, 2] ** 2      # a smooth "log-likelihood"
    batch = np.concatenate([X, Y[:, None]], axis=1)
    tr = SurrogateTrainer(P, 1, "cpu", seed=3, epochs=20)
    w1 = tr(batch, 1)
    assert {k: v.shape for k, v in w1.items()} == {"W1": (P, 64), "b1": (64,), "W2": (64, 16), "b2": (16,), "W3": (16, 1), "b3": (1,)}
    min_y, max_y = 2.0 * Y.min(), -1.0
    yn = (Y - min_y) / (max_y - min_y)
    pred = so.surrogate_forward(X.astype(np.float32), w1).reshape(-1)
    mse1 = float(np.mean((pred - yn) ** 2))
    np.testing.assert_allclose(mse1, tr.train_mse[-1], rtol=1e-3)
    assert mse1 < 0.5 * float(np.var(yn))                                            # better than predicting the mean
    w2 = tr(batch, 2)                                                                # warm start: keeps improving
    assert tr.train_mse[-1] < tr.train_mse[0]
    tr2 = SurrogateTrainer(P, 1, "cpu", seed=3, epochs=20)
    assert np.array_equal(tr2(batch, 1)["W1"], w1["W1"])                             # deterministic given the seed
    w3 = tr(batch, 1)                                                                # signature 1 again: fresh model
    assert not np.array_equal(w3["W1"], w2["W1"])


def _res_from_traces(g):
    S = int(g["samples"])
    res = {k: g["trace_" + k] for k in ("pos_w", "rmse_train", "rmse_test", "acc_train", "acc_test", "surg_likeh_list", "likeh_list", "accept_list")}
    res["naccept"] = np.asarray(g["trace_accept_ratio"]).reshape(-1) * S / 100.0
    return res


@pytest.mark.parametrize("name", ["run_rv_cancer", "run_rw_cancer", "run_lg_cancer", "run_lg_ions"])
def test_show_results_values_match_reference(name, tmp_path):
    """N2: the VALUES of the reference's own ParallelTempering.show_results (_revamp.py:1231-1392 and the forks'
    copies; recorded by oracle/make_golden.py from the unmodified reference fed with the per-chain files its chains
    wrote): burn-in slicing, the reshape order of pos_w / likelihood_vec / surr_list, rmse_surr, mean acceptance and
    the aggregated files. Same per-chain traces in -> identical 12-tuple out, through the file round trip
    (dump_chains -> show_results, the reference's text formats) and, up to those formats, in memory (aggregate)."""
    from sapt_b200 import results as R
    g = gu.load(name)
    S = int(g["samples"]); P = g["trace_pos_w"].shape[2]; burn = float(g["show_burn_in"])
    temps = list(g["temperatures"])
    ntr, nte = g["train"].shape[0], g["test"].shape[0]
    res = _res_from_traces(g)
    R.dump_chains(str(tmp_path), temps, res, S)
    out = R.show_results(str(tmp_path), temps, S, P, burn, True, ntr, nte)
    names = ("posterior", "fx_train", "fx_test", "rmse_train", "rmse_test", "acc_train", "acc_test", "likelihood_vec", "accept_list",
             "rmse_surr", "surr_list", "accept")
    for n, v in zip(names, out):
        if n.startswith("fx_"):
            continue
        ref = g["show_" + n]
        assert np.shape(v) == ref.shape, (n, np.shape(v), ref.shape)
        np.testing.assert_array_equal(np.asarray(v, dtype=np.float64), ref, err_msg=n)   # NaN == NaN here
    assert [tuple(s) for s in g["show_fx_shapes"]] == [out[1].shape, out[2].shape]
    for fn in ("pos_param", "likelihood", "acc_train", "acc_test"):
        np.testing.assert_array_equal(np.loadtxt(os.path.join(str(tmp_path), fn + ".txt")), g["showfile_" + fn], err_msg=fn)
    # in memory (no text round trip): equal up to the formats of the per-chain files (%1.2f, %1.4f)
    mem = R.aggregate(res, S, len(temps), P, burn, True, ntr, nte)
    np.testing.assert_array_equal(mem[0], g["show_posterior"])
    for idx, n, tol in ((3, "rmse_train", 0.0051), (4, "rmse_test", 0.0051), (5, "acc_train", 0.0051), (6, "acc_test", 0.0051),
                        (7, "likelihood_vec", 5.1e-5), (10, "surr_list", 5.1e-5), (8, "accept_list", 5.1e-5)):
        a, b = np.asarray(mem[idx], dtype=np.float64), g["show_" + n]
        assert a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.isinf(a), np.isinf(b)), n
        m = np.isfinite(b)
        np.testing.assert_allclose(a[m], b[m], rtol=0, atol=tol, err_msg=n)
    np.testing.assert_allclose(mem[9], float(g["show_rmse_surr"]), rtol=0, atol=2e-4)
    np.testing.assert_allclose(mem[11], float(g["show_accept"]), rtol=0, atol=1e-4)


def test_ladder_step_table_matches_reference():
    """default_beta_ladder with Tmax / ntemps derived from ptemcee's step table (RV:854-881): all 100 tabulated
    dimensions and the large-dimension formula beyond, against the reference's own output; product and oracle."""
    from oracle import sapt_oracle as so
    from sapt_b200 import default_beta_ladder
    g = gu.load("swap_ladder")
    for fn in (default_beta_ladder, so.default_beta_ladder):
        for nd, betas, nt in zip(g["ladder_ndim"], g["ladder_betas_ntemps5"], g["ladder_ntemps_Tmax50"]):
            np.testing.assert_allclose(fn(int(nd), 5, None), betas, rtol=1e-14)
            assert len(fn(int(nd), None, 50.0)) == int(nt)
