"""Importable alias of ``graph-convolutional-network-for-multi-camera-vehicle-tracking_b200/`` (a directory name
Python cannot import directly).  ``import gcn_mtmc_b200`` loads that package's modules from there."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "graph-convolutional-network-for-multi-camera-vehicle-tracking_b200")
__path__.insert(0, _PKG_DIR)
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
