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
