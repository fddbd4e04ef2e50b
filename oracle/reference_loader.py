"""Import the UNMODIFIED Python reference by file path.  TEST INFRASTRUCTURE ONLY.

Works only in the build container (where /root/reference is mounted); used by
tests/golden/make_golden.py to generate the committed golden vectors and by the
optional `-m "not gpu"` cross-checks that skip when the reference is absent.
Nothing here is reachable from the GPU-box run-time paths.

* `ot` (POT) is imported at ot_solvers.py:1 but only used by dead code
  (compute_transport_map_pot, ot_solvers.py:74-92); an empty stub module stands in.
* SpaDOT/__init__.py pulls anndata/scanpy (absent), so modules are loaded by path
  under a private package name instead of `import SpaDOT`.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("SPADOT_REFERENCE", "/root/reference")
_OT_DIR = os.path.join(REF_ROOT, "SpaDOT", "utils", "OT_loss")
_MODEL_DIR = os.path.join(REF_ROOT, "SpaDOT", "model")


def available() -> bool:
    return os.path.exists(os.path.join(_OT_DIR, "ot_solvers.py"))


def _load(name, path, package=None):
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=None)
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_ot_solvers():
    """Returns the reference `ot_solvers` module (with its ctypes-bound shipped libot.so)."""
    if "_spadot_ref_ot.ot_solvers" in sys.modules:
        return sys.modules["_spadot_ref_ot.ot_solvers"]
    sys.modules.setdefault("ot", types.ModuleType("ot"))
    pkg = types.ModuleType("_spadot_ref_ot")
    pkg.__path__ = [_OT_DIR]
    sys.modules["_spadot_ref_ot"] = pkg
    _load("_spadot_ref_ot.ot_func", os.path.join(_OT_DIR, "ot_func.py"), "_spadot_ref_ot")
    return _load("_spadot_ref_ot.ot_solvers", os.path.join(_OT_DIR, "ot_solvers.py"), "_spadot_ref_ot")


def load_svgp():
    """Returns the reference `svgp` module (SpaDOT/model/svgp.py; depends on torch only)."""
    if "_spadot_ref_svgp" in sys.modules:
        return sys.modules["_spadot_ref_svgp"]
    return _load("_spadot_ref_svgp", os.path.join(_MODEL_DIR, "svgp.py"))


def load_model():
    """Returns the reference `SpaDOT.model.SpaDOT` module (SpaDOT/model/SpaDOT.py) importable in the build
    container.  torch_geometric is absent, so `torch_geometric.nn.GATConv` is stubbed with the oracle's
    plain-torch restatement (oracle/gat_ref.py; parity of that piece is unpinned) — everything else
    (SVGPEncoder, Decoder, SVGP, the forward pass and its losses) is the unmodified reference."""
    if "_spadot_ref_model.SpaDOT" in sys.modules:
        return sys.modules["_spadot_ref_model.SpaDOT"]
    from . import gat_ref
    tg = types.ModuleType("torch_geometric")
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_nn.GATConv = gat_ref.GATConvRef
    tg.nn = tg_nn
    sys.modules.setdefault("torch_geometric", tg)
    sys.modules.setdefault("torch_geometric.nn", tg_nn)
    pkg = types.ModuleType("_spadot_ref_model")
    pkg.__path__ = [_MODEL_DIR]
    sys.modules["_spadot_ref_model"] = pkg
    for name in ("svgp", "encoder", "decoder"):
        _load(f"_spadot_ref_model.{name}", os.path.join(_MODEL_DIR, f"{name}.py"), "_spadot_ref_model")
    return _load("_spadot_ref_model.SpaDOT", os.path.join(_MODEL_DIR, "SpaDOT.py"), "_spadot_ref_model")
