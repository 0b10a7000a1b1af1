"""TEST INFRASTRUCTURE ONLY — loads the *unmodified* reference scripts for oracle pinning.

Only usable where ``/root/reference`` exists (the build container). Nothing in the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` may import this module at run time; it is
used by ``oracle/make_golden.py`` (fixture generation) and by the ``not gpu`` tests that
re-validate the numpy restatement against the live reference when it is present.

The reference scripts import keras / matplotlib / nn_mcmc_plots at module top
(/root/reference/_revamp.py:13-37); none is installed here and none is needed by the hot
path, so attribute-chaining stub modules are registered in ``sys.modules`` first.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SAPT_REFERENCE_ROOT", "/root/reference")

VARIANTS = {
    # abbreviations follow SURVEY.md
    "RV": "_revamp.py",
    "RW": "misc_codeapril2020/surrogate_rw_common_interval.py",
    "LG": "misc_codeapril2020/surrogate_pt_classifier_langevingrad.py",
    "LS": "localsurrogates/surrogate_pt_classifier_rw_localsurrogate.py",
}


class _Anything:
    """Callable, attribute-chaining, dict-like dummy used for every stubbed symbol."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __getitem__(self, key):
        return _Anything()

    def __setitem__(self, key, value):
        pass

    def update(self, *a, **k):
        pass

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()


_STUBS = [
    "matplotlib", "matplotlib.mlab", "matplotlib.pyplot", "matplotlib.ticker",
    "matplotlib.patches", "matplotlib.collections",
    "keras", "keras.models", "keras.layers", "keras.objectives", "keras.callbacks",
    "nn_mcmc_plots",
]


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, VARIANTS["RV"]))


def _install_stubs():
    for name in _STUBS:
        if name not in sys.modules:
            mod = _StubModule(name)
            mod.__path__ = []  # make it look like a package so "from a.b import c" works
            sys.modules[name] = mod


_cache = {}


def load_reference(variant: str = "LG"):
    """Return the reference script ``variant`` as a module (classes Network, ptReplica,
    ParallelTempering, surrogate). ``main()`` is never executed (guarded by __name__)."""
    if variant in _cache:
        return _cache[variant]
    if not reference_available():
        raise FileNotFoundError(f"reference not present under {REFERENCE_ROOT}")
    _install_stubs()
    path = os.path.join(REFERENCE_ROOT, VARIANTS[variant])
    spec = importlib.util.spec_from_file_location(f"_sapt_reference_{variant}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache[variant] = mod
    return mod


class DuckReplica:
    """Duck-typed stand-in for ``ptReplica`` so its unbound methods can be driven without
    spawning a process (ptReplica.__init__ takes 24 positional IPC arguments,
    misc_codeapril2020/surrogate_pt_classifier_langevingrad.py:332)."""

    def __init__(self, ref_module, topology, temperature=1.0):
        self._ref = ref_module
        self.topology = list(topology)
        self.temperature = float(temperature)
        self.adapttemp = float(temperature)

    def rmse(self, pred, actual):
        return self._ref.ptReplica.rmse(self, pred, actual)

    def accuracy(self, pred, actual):
        return self._ref.ptReplica.accuracy(self, pred, actual)

    def likelihood_func(self, fnn, data, w):
        return self._ref.ptReplica.likelihood_func(self, fnn, data, w)

    def prior_likelihood(self, sigma_squared, nu_1, nu_2, w):
        return self._ref.ptReplica.prior_likelihood(self, sigma_squared, nu_1, nu_2, w)
