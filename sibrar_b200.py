"""Import alias: the package directory is named ``sibrar---single-branch-recommender_b200`` (not a Python
identifier), so this module loads it under the importable name ``sibrar_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sibrar---single-branch-recommender_b200")
_spec = importlib.util.spec_from_file_location("sibrar_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_pkg = importlib.util.module_from_spec(_spec)
sys.modules["sibrar_b200"] = _pkg
_spec.loader.exec_module(_pkg)
