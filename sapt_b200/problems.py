"""Problem table + preprocessing of the reference's main() (_revamp.py:1425-1542): 8 problems
(file, topology, surrogate topology), z-scoring, 60/40 split. Pure numpy (host side, off the hot path)."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np


@dataclass(frozen=True)
class ProblemSpec:
    number: int
    name: str
    files: Tuple[str, ...]
    delimiter: str
    ip: int
    hidden: int
    output: int
    surrogate_topology: int
    separate: bool          # one file, z-score + random 60/40 split (separate_flag, _revamp.py:1521)
    skip_header: bool = False
    label_offset: int = 0   # iris labels are 1-based in the file (_revamp.py:1441)


# hidden sizes are the reference's defaults; its commented alternatives (Cancer 12, Ionosphere 50) are what
# BASELINE.json's C1/C2 use — pass `hidden=` to load_problem to select them.
PROBLEMS = {
    1: ProblemSpec(1, "winequality-red", ("winequality-red.csv",), ";", 11, 50, 10, 2, True, skip_header=True),
    2: ProblemSpec(2, "winequality-white", ("winequality-white.csv",), ";", 11, 50, 10, 2, True, skip_header=True),
    3: ProblemSpec(3, "iris", ("iris.csv",), ";", 4, 8, 3, 1, True, label_offset=-1),
    4: ProblemSpec(4, "Ionosphere", ("Ions/Ions/ftrain.csv", "Ions/Ions/ftest.csv"), ",", 34, 15, 2, 1, False),
    5: ProblemSpec(5, "Cancer", ("Cancer/ftrain.txt", "Cancer/ftest.txt"), " ", 9, 8, 2, 1, False),
    6: ProblemSpec(6, "bank-additional", ("Bank/bank-processed.csv",), ";", 20, 50, 2, 2, True),
    7: ProblemSpec(7, "PenDigit", ("PenDigit/train.csv", "PenDigit/test.csv"), ",", 16, 30, 10, 2, False),
    8: ProblemSpec(8, "chess", ("chess.csv",), ";", 6, 25, 18, 3, True),
}


def zscore_columns(a: np.ndarray, ncols: int) -> np.ndarray:
    a = a.copy()
    for k in range(ncols):
        a[:, k] = (a[:, k] - np.mean(a[:, k])) / np.std(a[:, k])
    return a


def split_60_40(features: np.ndarray, classes: np.ndarray, rng: Optional[np.random.RandomState] = None, train_ratio: float = 0.6):
    """_revamp.py:1531-1541: z-score every feature over the whole file, random permutation, 60/40."""
    rng = rng or np.random
    feats = zscore_columns(features, features.shape[1])
    idx = rng.permutation(feats.shape[0])
    ntr = int(train_ratio * feats.shape[0])
    tr = np.hstack([feats[idx[:ntr], :], classes[idx[:ntr], :]])
    te = np.hstack([feats[idx[ntr:], :], classes[idx[ntr:], :]])
    return tr, te


def load_problem(problem: int, data_root: str, hidden: Optional[int] = None, rng: Optional[np.random.RandomState] = None):
    """-> (traindata, testdata, topology, surrogate_topology, name) exactly as main() builds them."""
    spec = PROBLEMS[int(problem)]
    paths = [os.path.join(data_root, f) for f in spec.files]
    for p in paths:
        if not os.path.isfile(p):
            raise FileNotFoundError(f"{p} (problem {spec.number} '{spec.name}'); the reference repository itself lacks chess.csv, "
                                    "bank-processed.csv (.MISSING_LARGE_BLOBS)")
    if spec.separate:
        data = np.genfromtxt(paths[0], delimiter=spec.delimiter)
        if spec.skip_header:
            data = data[1:, :]
        classes = data[:, spec.ip].reshape(-1, 1) + spec.label_offset
        features = data[:, 0:spec.ip]
        tr, te = split_60_40(features, classes, rng)
    else:
        tr = np.genfromtxt(paths[0], delimiter=spec.delimiter)
        te = np.genfromtxt(paths[1], delimiter=spec.delimiter)
        if spec.number in (4, 5):            # Ionosphere / Cancer files carry a trailing one-hot column (_revamp.py:1456,1468)
            tr, te = tr[:, :-1], te[:, :-1]
        if spec.number == 7:                 # PenDigit: z-score train and test separately (_revamp.py:1494-1503)
            tr, te = zscore_columns(tr, 16), zscore_columns(te, 16)
    topology = [spec.ip, int(hidden) if hidden else spec.hidden, spec.output]
    return tr, te, topology, spec.surrogate_topology, spec.name
