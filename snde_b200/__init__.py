"""Importable alias of the package directory ``self-supervised-normal-depth-estimation_b200/``.

The repository layout names the package after the reference repo (hyphens and all),
which is not a valid Python identifier; this stub points ``snde_b200``'s search path at
that directory so ``import snde_b200.loss_functions`` etc. work.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "self-supervised-normal-depth-estimation_b200")
__path__ = [_PKG_DIR]
PACKAGE_DIR = _PKG_DIR
__version__ = "0.1.0"
