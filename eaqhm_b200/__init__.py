"""Importable alias of the package directory ``eaqhm-analysis-and-synthesis-in-python_b200/``.

The contract fixes the directory name, which is not a valid Python identifier;
this stub makes ``import eaqhm_b200`` resolve to it (same files, no copy).
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "eaqhm-analysis-and-synthesis-in-python_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
