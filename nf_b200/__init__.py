"""Import alias: the package directory is named after the reference repository
(``normalising-flows-4-reinforcement-learning_b200``), which is not a valid Python identifier, so
``nf_b200`` re-homes its module search path there."""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "normalising-flows-4-reinforcement-learning_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
