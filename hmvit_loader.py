"""Registers the package directory `hm-vit_b200/` (hyphenated, so not importable by name) as the
module `hmvit_b200`.  Usage:  import hmvit_loader; hmvit = hmvit_loader.load()"""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(_ROOT, "hm-vit_b200")


def load():
    if "hmvit_b200" in sys.modules:
        return sys.modules["hmvit_b200"]
    spec = importlib.util.spec_from_file_location("hmvit_b200", os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["hmvit_b200"] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        del sys.modules["hmvit_b200"]
        raise
    return mod
