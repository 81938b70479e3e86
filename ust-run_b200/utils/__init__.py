"""Drop-in replacements for the hot-path parts of the reference's ``utils`` package."""

# Tier A: this package shadows the reference's package of the same name, which is a namespace directory
# (no __init__.py) further down sys.path.  Only the hot-path modules live here; every other module of the reference's
# package (metrics.py, util.py, wrn.py, deeplabv2.py, backbone/, ...) keeps resolving to the reference's own file because
# those directories are appended to this package's search path -- no copies or symlinks.
import os as _os
import sys as _sys

for _p in list(_sys.path):
    _d = _os.path.join(_p or ".", __name__.split(".")[-1])
    if _os.path.isdir(_d) and _os.path.realpath(_d) not in [_os.path.realpath(_q) for _q in __path__]:
        __path__.append(_d)
