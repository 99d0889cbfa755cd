"""Import alias: ``import jfnk_b200`` loads the package in ``iterative-solvers-summer-2020_b200/``
(whose directory name is not a valid Python identifier)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_real = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "iterative-solvers-summer-2020_b200")
_spec = _ilu.spec_from_file_location("jfnk_b200", _os.path.join(_real, "__init__.py"), submodule_search_locations=[_real])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["jfnk_b200"] = _mod
_spec.loader.exec_module(_mod)
