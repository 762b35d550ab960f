"""Importable alias of the product package.

The product lives in ``polycracker-unofficial-mirror_b200/`` (the directory name the build contract asks for);
a hyphen is not importable, so this shim makes ``import polycracker_b200.<module>`` resolve there.
"""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
PACKAGE_DIR = _os.path.join(_os.path.dirname(_here), "polycracker-unofficial-mirror_b200")
__path__.insert(0, PACKAGE_DIR)
__version__ = "0.1.0"
