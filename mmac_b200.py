"""Import alias for the package directory ``multi-modal-art-classifier_b200/``.

The directory name carries hyphens (it follows the upstream repository's name), so it
cannot be written in an ``import`` statement.  ``import mmac_b200`` executes this file, which
loads that directory as a regular package under the name ``mmac_b200`` and replaces itself in
``sys.modules``; ``import mmac_b200.nn`` etc. then resolve inside the package directory.
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "multi-modal-art-classifier_b200")
_spec = importlib.util.spec_from_file_location(
    "mmac_b200", os.path.join(_PKG_DIR, "__init__.py"),
    submodule_search_locations=[_PKG_DIR])
_pkg = importlib.util.module_from_spec(_spec)
sys.modules["mmac_b200"] = _pkg
_spec.loader.exec_module(_pkg)
