"""Pins oracle/sapt_oracle.py against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py). CPU only."""
import numpy as np
import pytest

from oracle import sapt_oracle as so
import golden_util as gu


@pytest.mark.parametrize("problem", ["cancer", "ions", "pendigit"])
def test_unit_forward_likelihood_prior(problem):
    g = gu.load(f"unit_{problem}")
    topo = [int(v) for v in g["topology"]]
    for k in range(g["W"].shape[0]):
        w, T = g["W"][k], float(g["temps"][k])
        fx, prob = so.evaluate_proposal(g["train"], w, topo)
        assert np.array_equal(fx, g["fx_train"][k])                      # class indices: exact
        np.testing.assert_allclose(prob, g["prob_train"][k], rtol=0, atol=5e-16)
        lt, fx2, rm, l = so.likelihood_func(g["train"], w, topo, T)
        np.testing.assert_allclose(l, g["L_train"][k], rtol=1e-13)
        np.testing.assert_allclose(lt, g["Ltemp_train"][k], rtol=1e-13)
        np.testing.assert_allclose(rm, g["rmse_train"][k], rtol=1e-14)
        assert so.accuracy(fx, g["train"][:, topo[0]]) == g["acc_train"][k]
        lt, fx3, rm, l = so.likelihood_func(g["test"], w, topo, T)
        assert np.array_equal(fx3, g["fx_test"][k])
        np.testing.assert_allclose(l, g["L_test"][k], rtol=1e-13)
        np.testing.assert_allclose(rm, g["rmse_test"][k], rtol=1e-14)
        np.testing.assert_allclose(so.prior_likelihood(25, 0, 0, w, topo), g["prior"][k], rtol=1e-14)
        L, nc, sq = so.loglik_stats(g["train"], w, topo)
        np.testing.assert_allclose(np.sqrt(sq / g["train"].shape[0]), g["rmse_train"][k], rtol=1e-14)
        assert 100.0 * (nc / g["train"].shape[0]) == g["acc_train"][k]   # RV:389 rounding order


def test_rowloop_port_equals_vectorised():
    g = gu.load("unit_cancer")
    topo = [int(v) for v in g["topology"]]
    a = so.likelihood_func(g["train"][:64], g["W"][0], topo, 2.0, rowloop=True)
    b = so.likelihood_func(g["train"][:64], g["W"][0], topo, 2.0, rowloop=False)
    np.testing.assert_allclose(a[0], b[0], rtol=1e-13)
    assert np.array_equal(a[1], b[1])
    np.testing.assert_allclose(a[0], g["Ltemp_train"][0] * 0 + a[0])  # shape sanity


@pytest.mark.parametrize("problem", ["cancer", "ions", "pendigit"])
def test_unit_langevin_sgd(problem):
    g = gu.load(f"unit_{problem}")
    topo = [int(v) for v in g["topology"]]
    n = int(g["langevin_rows"])
    for li, lr in enumerate(g["langevin_lr"]):
        for k in range(g["W"].shape[0]):
            got = so.langevin_gradient(g["train"][:n], g["W"][k], topo, float(lr), 1)
            np.testing.assert_allclose(got, g["langevin_out"][li][k], rtol=1e-12, atol=1e-13)


def test_langevin_batch_is_first_order_limit():
    """langevin_batch_gradient is the lr->0 limit of the sequential epoch: (w_sgd - w)/lr -> g."""
    g = gu.load("unit_cancer")
    topo = [int(v) for v in g["topology"]]
    w = g["W"][3]
    data = g["train"][:50]
    grad = so.langevin_batch_gradient(data, w, topo)
    lr = 1e-7
    fd = (so.langevin_gradient(data, w, topo, lr, 1) - w) / lr
    np.testing.assert_allclose(fd, grad, rtol=2e-4, atol=1e-6)


def test_ladder_and_swap_sweep():
    g = gu.load("swap_ladder")
    for R, mt in ((8, 2.0), (10, 4.0), (4, 4.0), (16, 6.0)):
        np.testing.assert_allclose(so.assign_temperatures(R, mt), g[f"ladder_R{R}_T{mt}"], rtol=1e-14)
        np.testing.assert_allclose(so.assign_temperatures(R, mt), mt ** (np.arange(R) / (R - 1)), rtol=1e-12)
    for n in range(int(g["n_swap_cases"])):
        src, sw = so.swap_sweep(g[f"swap{n}_lik"], g[f"swap{n}_u"])
        assert np.array_equal(src, g[f"swap{n}_src"])
        assert np.array_equal(sw, g[f"swap{n}_swapped"])
        assert int(sw.sum()) == int(g[f"swap{n}_num_swap"])
    with pytest.raises(ValueError):
        so.default_beta_ladder(2, ntemps=4, Tmax=1.0)
    with pytest.raises(ValueError):
        so.default_beta_ladder(0, ntemps=4, Tmax=2.0)


@pytest.mark.parametrize("name", list(gu.RUNS))
def test_full_run_traces(name):
    """oracle.run_pt == the reference's ptReplica.run x R + coordinator, step for step."""
    g = gu.load(name)
    out = so.run_pt(gu.RUNS[name], **gu.run_inputs(g))
    # pos_w identical => every accept/reject and swap decision identical
    assert np.array_equal(out["pos_w"], g["trace_pos_w"])
    assert np.array_equal(out["accept_list"], g["trace_accept_list"])
    gu.compare_traces(g, out, rtol=1e-12, atol=1e-11)
    n_sweeps = len(out["swap_decisions"])
    ref_log = g["swap_log"][: n_sweeps * (len(g["temperatures"]) - 1)]
    if n_sweeps and gu.RUNS[name] == "rw":   # LG's swap log is produced by the barrier harness, RV never swaps
        assert np.array_equal(np.concatenate(out["swap_decisions"]).astype(int), ref_log)
    if gu.RUNS[name] == "rv":
        assert n_sweeps == 0 and int(g["num_swap"]) == 0
    for s, b in enumerate(out["surr_batches"]):
        np.testing.assert_allclose(b, g[f"trainer_batch_{s}"], rtol=1e-12, atol=1e-11)
    if float(g["surrogate_prob"]) > 0:
        assert out["is_surrogate_step"].sum() > 0          # the surrogate branch was exercised
        np.testing.assert_allclose(out["min_y"], float(g["min_y"]), rtol=1e-13)
    if bool(g["use_langevin"]):
        assert out["langevin_step"].sum() > 0


def test_local_surrogate_fork_traces():
    """N4: oracle.run_pt_local == the local-surrogate fork's ptReplica.run x R (LS:447-836), including the
    per-chain 4-10-1 networks trained by the reference's own backprop_train."""
    g = gu.load("run_ls_cancer")
    out = so.run_pt_local(**gu.run_local_inputs(g))
    assert np.array_equal(out["accept_list"], g["trace_accept_list"])          # every accept decision
    gu.compare_traces(g, out, rtol=1e-12, atol=1e-11, keys=gu.LS_TRACE_KEYS)
    assert out["num_swap"] == int(g["num_swap"])
    assert out["is_surrogate_step"].sum() > 0
    R = len(g["temperatures"])
    nc = int(g["n_train_calls"])
    assert nc == len(out["train_rmse"]) and nc >= 2
    for c in range(nc):
        for k in range(R):
            np.testing.assert_allclose(out["train_rmse"][c][k], float(g[f"ls_train{c}_rmse_{k}"]), rtol=1e-10)
    for k in range(R):
        net = out["nets"][k]
        w = np.concatenate([net["W1"].ravel(), net["W2"].ravel(), net["B1"].ravel(), net["B2"].ravel()])
        np.testing.assert_allclose(w, g[f"ls_train{nc - 1}_weights_{k}"], rtol=1e-10, atol=1e-12)


def test_local_surrogate_units():
    """Takens windows and one training call against what the reference's surrogate.train received / produced."""
    g = gu.load("run_ls_cancer")
    v, y = so.takens_windows(np.arange(23.0))
    assert y.shape[0] == 20 and v.shape == (7, 5) and np.array_equal(v[3], np.arange(3.0, 8.0) + 6.0)
    k = 1
    data0 = g[f"ls_train0_data_{k}"]
    net = so.local_net_init(g["net_init"][k])
    rm = so.local_net_train(data0, net)
    np.testing.assert_allclose(rm, float(g[f"ls_train0_rmse_{k}"]), rtol=1e-10)
    w = np.concatenate([net["W1"].ravel(), net["W2"].ravel(), net["B1"].ravel(), net["B2"].ravel()])
    np.testing.assert_allclose(w, g[f"ls_train0_weights_{k}"], rtol=1e-10, atol=1e-12)


def test_parallel_sweep_algorithm_equals_sequential_sweep():
    """The algorithm of the device's parallel exchange sweep (sparse table + pointer doubling), restated in numpy,
    against the reference's sequential rule on random ladders: ties, u >= 1, all-swap and no-swap runs."""
    rs = np.random.RandomState(4)
    for case in range(300):
        R = int(rs.choice([2, 3, 4, 5, 9, 33, 200, 1025]))
        scale = float(rs.choice([0.1, 1.0, 30.0]))
        lik = -np.abs(rs.standard_normal(R)) * scale - 100.0
        if case % 5 == 0:
            lik = np.sort(lik)                       # every pair wants to swap
        if case % 7 == 0:
            lik[rs.randint(0, R, 2)] = lik[0]        # ties
        u = rs.uniform(size=R - 1)
        u[u == 0] = 0.5
        if case % 3 == 0 and R > 2:
            u[rs.randint(0, R - 1)] = 1.0            # never swaps
        s1, w1 = so.swap_sweep(lik, u)
        s2, w2 = so.swap_sweep_parallel(lik, u)
        assert np.array_equal(w1, w2), (case, R)
        assert np.array_equal(s1, s2), (case, R)
    g = gu.load("swap_ladder")
    for n in range(int(g["n_swap_cases"])):
        s2, w2 = so.swap_sweep_parallel(g[f"swap{n}_lik"], g[f"swap{n}_u"])
        assert np.array_equal(s2, g[f"swap{n}_src"]) and np.array_equal(w2, g[f"swap{n}_swapped"])
