"""Importable alias of the package directory ``surrogate-assisted-parallel-tempering_b200/``
(a hyphenated directory name cannot be imported directly). ``import sapt_b200`` executes that
directory's ``__init__.py`` with this package's ``__path__`` pointing at it, so
``sapt_b200.engine`` etc. resolve to the files there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "surrogate-assisted-parallel-tempering_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
