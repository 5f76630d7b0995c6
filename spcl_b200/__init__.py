"""Import alias: the product package lives in ``self-paced-contrastive-learning_b200/`` (not a valid
Python identifier), this shim loads it under the importable name ``spcl_b200``."""
import importlib.util
import pathlib
import sys

_root = pathlib.Path(__file__).resolve().parent.parent / "self-paced-contrastive-learning_b200"
_spec = importlib.util.spec_from_file_location("spcl_b200", _root / "__init__.py",
                                               submodule_search_locations=[str(_root)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["spcl_b200"] = _mod
_spec.loader.exec_module(_mod)
