"""Loads the package directory ``pti-ldm-vae_b200/`` (hyphens are not importable) under the module
name ``pti_ldm_vae_b200``.  Usage:  ``import _pkg; b200 = _pkg.load()``."""
import importlib.util
import pathlib
import sys

NAME = "pti_ldm_vae_b200"
ROOT = pathlib.Path(__file__).resolve().parent / "pti-ldm-vae_b200"


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, ROOT / "__init__.py", submodule_search_locations=[str(ROOT)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod
