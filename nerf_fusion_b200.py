"""Importable alias of the ``nerf-fusion_b200/`` package (a hyphen is not a valid identifier):
``import nerf_fusion_b200 as dfb`` gives the same module object as importlib.import_module("nerf-fusion_b200")."""
import importlib
import sys

_pkg = importlib.import_module("nerf-fusion_b200")
sys.modules[__name__] = _pkg
