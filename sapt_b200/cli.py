"""Command line of the reference (_revamp.py:1402-1411,1563-1570):

    python -m sapt_b200.cli problem samples surrogate_prob surrogate_intervalratio foldername num_chains
           [--data-root DIR] [--variant rw|lg|rv|ls] [--hidden H] [--seed S] [--no-dump]

Same six positional arguments; constants as in the chosen fork's main() (rv: maxtemp 2, swap_interval 5,
burn_in 0.6, learn_rate 0.01; rw: 4, 50, 0.2, 0.5; lg: as rw with Langevin gradients, swap_interval 100000; ls: the local-surrogate fork, 4, 100000, 0.6, 0.01)."""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

FORK_CONSTANTS = {  # (maxtemp, swap_interval, burn_in, use_langevin, learn_rate)
    "rv": (2, 5, 0.6, False, 0.01),          # _revamp.py:1555-1608
    "rw": (4, 50, 0.2, False, 0.5),          # surrogate_rw_common_interval.py:1484-1486
    "lg": (4, 100000, 0.2, True, 0.5),       # surrogate_pt_classifier_langevingrad.py:1672,1723-1725
    "ls": (4, 100000, 0.6, False, 0.01),     # localsurrogates/surrogate_pt_classifier_rw_localsurrogate.py:1543-1546,1600-1602
}


def main(argv=None):
    ap = argparse.ArgumentParser(prog="sapt_b200.cli", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("problem", type=int); ap.add_argument("samples", type=int); ap.add_argument("surrogate_prob", type=float)
    ap.add_argument("surrogate_intervalratio", type=float); ap.add_argument("foldername"); ap.add_argument("num_chains", type=int)
    ap.add_argument("--data-root", default=os.environ.get("SAPT_DATA_ROOT", "DATA"))
    ap.add_argument("--variant", default="rv", choices=list(FORK_CONSTANTS))
    ap.add_argument("--hidden", type=int, default=None)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-dump", action="store_true", help="skip the per-chain text dumps (large)")
    a = ap.parse_args(argv)

    from .parallel_tempering import ParallelTempering
    from .problems import load_problem
    from . import results as R

    rng = np.random.RandomState(a.seed)
    traindata, testdata, topology, sur_topo, name = load_problem(a.problem, a.data_root, a.hidden, rng)
    maxtemp, swap_interval, burn_in, use_lg, lr = FORK_CONSTANTS[a.variant]
    surrogate_interval = int(a.surrogate_intervalratio * (a.samples / a.num_chains))       # _revamp.py:1570
    run_nb = 0
    while os.path.exists(f"{a.foldername}{name}_{run_nb}"):
        run_nb += 1
    path = f"{a.foldername}{name}_{run_nb}"
    R.make_directories(path)
    t0 = time.time()
    pt = ParallelTempering(a.surrogate_prob > 0, use_lg, lr, False, traindata, testdata, topology, a.num_chains, maxtemp, a.samples,
                           swap_interval, max(1, surrogate_interval), a.surrogate_prob, path, path, sur_topo, variant=a.variant, seed=a.seed,
                           record_pos=not a.no_dump or a.variant != "ls")
    pt.initialize_chains(burn_in)
    (pos_w, fx_train, fx_test, rmse_train, rmse_test, acc_train, acc_test, accept_list, swap_perc, likelihood_rep, rmse_surr,
     surr_list, accept) = pt.run_chains()
    span = time.time() - t0
    if not a.no_dump:
        R.dump_chains(path, pt.temperatures, pt.results, pt.NumSamples)
    R.dump_aggregate(path, pos_w, likelihood_rep, acc_train, acc_test)
    row = R.result_row(a.problem, a.samples, maxtemp, swap_interval, a.surrogate_intervalratio, a.surrogate_prob, use_lg, lr, acc_train,
                       acc_test, swap_perc, accept, rmse_surr, span / 60, int(span))
    with open(os.path.join(path, "result.txt"), "a+") as f:
        np.savetxt(f, row, fmt="%1.2f", newline=" ")
    with open("master_result_file.txt", "a+") as f:
        np.savetxt(f, row, fmt="%1.2f", newline=" ")
        mse = getattr(pt._trainer, "train_mse", []) if pt._trainer is not None else []
        np.savetxt(f, (np.mean(mse) if mse else 0.0, np.std(mse) if mse else 0.0), fmt="%1.2e", newline=" ")
        np.savetxt(f, [f"{name}_{run_nb}"], fmt="%s", newline=" \n")
    print(f"{name}: train {row[8]:.2f} +- {row[9]:.2f} (max {row[10]:.2f})  test {row[11]:.2f} +- {row[12]:.2f} (max {row[13]:.2f})  "
          f"swap {swap_perc:.1f}%  accept {accept:.1f}%  {span:.1f}s -> {path}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
