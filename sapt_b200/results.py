"""On-disk result layout of the reference (§8f N2): per-chain dumps of ptReplica.run (_revamp.py:746-777), the
aggregated files of show_results (_revamp.py:1383-1389) and the result.txt / master_result_file.txt row
(_revamp.py:1705-1710). Host side only; consumes PTEngine.results()."""
from __future__ import annotations

import os

import numpy as np

DIRS = ("predictions", "posterior", "results", "surrogate", "surrogate/learnsurrogate_data", "posterior/pos_w",
        "posterior/pos_likelihood", "posterior/surg_likelihood", "posterior/accept_list")


def make_directories(path: str):
    for d in DIRS:
        os.makedirs(os.path.join(path, d), exist_ok=True)


def dump_chains(path: str, temperatures, res: dict, samples: int):
    """One set of files per chain, named by its temperature exactly like the reference."""
    make_directories(path)
    for k, T in enumerate(temperatures):
        t = str(T)
        if "pos_w" in res:
            np.savetxt(os.path.join(path, "posterior/pos_w", f"chain_{t}.txt"), res["pos_w"][k])
        for key, fn in (("rmse_test", "rmse_test"), ("rmse_train", "rmse_train"), ("acc_test", "acc_test"), ("acc_train", "acc_train")):
            np.savetxt(os.path.join(path, "predictions", f"{fn}_chain_{t}.txt"), res[key][k], fmt="%1.2f")
        np.savetxt(os.path.join(path, "posterior/surg_likelihood", f"chain_{t}.txt"), res["surg_likeh_list"][k], fmt="%1.4f")
        np.savetxt(os.path.join(path, "posterior/pos_likelihood", f"chain_{t}.txt"), res["likeh_list"][k], fmt="%1.4f")
        np.savetxt(os.path.join(path, "posterior/accept_list", f"chain_{t}_accept.txt"), [res["naccept"][k] / (samples * 1.0) * 100], fmt="%1.4f")
        np.savetxt(os.path.join(path, "posterior/accept_list", f"chain_{t}.txt"), res["accept_list"][k], fmt="%1.4f")


def dump_aggregate(path: str, posterior, likelihood_vec, acc_train, acc_test):
    np.savetxt(os.path.join(path, "pos_param.txt"), posterior.T)
    np.savetxt(os.path.join(path, "likelihood.txt"), likelihood_vec, fmt="%1.5f")
    np.savetxt(os.path.join(path, "acc_train.txt"), acc_train, fmt="%1.2f")
    np.savetxt(os.path.join(path, "acc_test.txt"), acc_test, fmt="%1.2f")


def result_row(problem, NumSample, maxtemp, swap_interval, interval_ratio, surrogate_prob, use_langevin, learn_rate, acc_train, acc_test,
               swap_perc, accept, rmse_surr, minutes, seconds):
    """The 19 numbers of `allres` (_revamp.py:1705); inf entries (surrogate steps) are dropped first (_revamp.py:1650-1660)."""
    tr = acc_train[np.isfinite(acc_train)]; te = acc_test[np.isfinite(acc_test)]
    stat = lambda a: (float(np.mean(a)), float(np.std(a)), float(np.amax(a))) if a.size else (float("nan"),) * 3
    return np.asarray([problem, NumSample, maxtemp, swap_interval, interval_ratio, surrogate_prob, float(use_langevin), learn_rate,
                       *stat(tr), *stat(te), swap_perc, accept, rmse_surr, minutes, seconds], dtype=np.float64)


def aggregate(res: dict, samples: int, num_chains: int, num_param: int, burn_in: float, use_surrogate: bool, n_train: int, n_test: int,
              accept_percent=None, keep_fx: bool = True):
    """`ParallelTempering.show_results` (_revamp.py:1231-1392) on in-memory per-chain histories: burn-in slicing,
    the (param | likelihood column, chain, sample) reshape order of `pos_w` / `likelihood_vec` / `surr_list`, `rmse_surr` over
    the rows without NaN, mean acceptance. `res` holds arrays indexed [chain, sample, ...] (`PTEngine.results()`).
    -> the reference's 12-tuple (posterior, fx_train_all, fx_test_all, rmse_train, rmse_test, acc_train, acc_test,
    likelihood_vec.T, accept_list, rmse_surr, surr_list, accept)."""
    S, R, P = int(samples), int(num_chains), int(num_param)
    burnin = int(S * burn_in)
    keep = S - burnin
    pos_w = np.asarray(res["pos_w"])[:, burnin:, :] if "pos_w" in res else np.zeros((R, keep, P))
    posterior = pos_w.transpose(2, 0, 1).reshape(P, -1)
    # the reference returns all-zero fx arrays (their loading code is commented out, _revamp.py:1270-1275)
    fx_train_all = np.zeros((R, keep, n_train)) if (keep_fx and R * keep * n_train <= (1 << 26)) else None
    fx_test_all = np.zeros((R, keep, n_test)) if (keep_fx and R * keep * n_test <= (1 << 26)) else None
    col = lambda k: np.asarray(res[k])[:, burnin:].reshape(R * keep, 1)
    rmse_train, rmse_test, acc_train, acc_test = col("rmse_train"), col("rmse_test"), col("acc_train"), col("acc_test")
    likelihood_rep = np.asarray(res["likeh_list"])[:, 1:, :]
    likelihood_vec = likelihood_rep.transpose(2, 0, 1).reshape(2, -1)
    surg = np.asarray(res["surg_likeh_list"])[:, 1:, 0:2]
    surg_vec = surg.transpose(2, 0, 1).reshape(2, -1)
    accept_list = np.asarray(res["accept_list"])
    rmse_surr, surr_list = 0, []
    if use_surrogate:
        surr_list = surg_vec.T
        sl = surr_list[~np.isnan(surr_list).any(axis=1)]
        rmse_surr = float(np.sqrt(((sl[:, 1] - sl[:, 0]) ** 2).mean())) if len(sl) else 0.0
    if accept_percent is None:
        accept_percent = np.asarray(res["naccept"]) / (S * 1.0) * 100     # _revamp.py:740
    accept = float(np.sum(accept_percent) / R)
    return (posterior, fx_train_all, fx_test_all, rmse_train, rmse_test, acc_train, acc_test, likelihood_vec.T, accept_list,
            rmse_surr, surr_list, accept)


def load_chains(path: str, temperatures, with_pos: bool = True) -> dict:
    """Read the per-chain files of `dump_chains` (or of a reference run) back into the [chain, sample, ...] layout."""
    res = {k: [] for k in ("rmse_test", "rmse_train", "acc_test", "acc_train", "surg_likeh_list", "likeh_list", "accept_list")}
    pos, acc = [], []
    for T in temperatures:
        t = str(T)
        if with_pos:
            pos.append(np.loadtxt(os.path.join(path, "posterior/pos_w", f"chain_{t}.txt")))
        for key in ("rmse_test", "rmse_train", "acc_test", "acc_train"):
            res[key].append(np.loadtxt(os.path.join(path, "predictions", f"{key}_chain_{t}.txt")))
        res["surg_likeh_list"].append(np.loadtxt(os.path.join(path, "posterior/surg_likelihood", f"chain_{t}.txt")))
        res["likeh_list"].append(np.loadtxt(os.path.join(path, "posterior/pos_likelihood", f"chain_{t}.txt")))
        res["accept_list"].append(np.loadtxt(os.path.join(path, "posterior/accept_list", f"chain_{t}.txt")))
        acc.append(float(np.loadtxt(os.path.join(path, "posterior/accept_list", f"chain_{t}_accept.txt"))))
    out = {k: np.array(v) for k, v in res.items()}
    if with_pos:
        out["pos_w"] = np.array(pos)
    out["accept_percent"] = np.array(acc)
    return out


def show_results(path: str, temperatures, samples: int, num_param: int, burn_in: float, use_surrogate: bool, n_train: int, n_test: int,
                 write: bool = True):
    """The file-based `show_results` of the reference: per-chain files in, the 12-tuple out, and (write=True) the aggregated
    files pos_param.txt / likelihood.txt / acc_train.txt / acc_test.txt / surrogate/surg_likelihood.txt (_revamp.py:1365-1389)."""
    res = load_chains(path, temperatures)
    out = aggregate(res, samples, len(temperatures), num_param, burn_in, use_surrogate, n_train, n_test, accept_percent=res["accept_percent"])
    if write:
        dump_aggregate(path, out[0], out[7], out[5], out[6])
        if use_surrogate and len(out[10]):
            sl = out[10][~np.isnan(out[10]).any(axis=1)]
            os.makedirs(os.path.join(path, "surrogate"), exist_ok=True)
            np.savetxt(os.path.join(path, "surrogate", "surg_likelihood.txt"), sl, fmt="%1.5f")
    return out
