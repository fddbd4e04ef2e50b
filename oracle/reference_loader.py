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


_PYREF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "pyref")


def available() -> bool:
    return os.path.exists(os.path.join(_OT_DIR, "ot_solvers.py"))


def compiled_available() -> bool:
    """oracle/_ref/pyref/*.pycode: the reference's ot_func.py / ot_solvers.py byte-compiled UNMODIFIED by oracle/Makefile
    (`make pyref`, build container only).  Like oracle/_ref/libot_ref.so they are git-ignored build outputs that travel
    to the GPU box, where /root/reference does not exist."""
    return all(os.path.exists(os.path.join(_PYREF, f)) for f in ("ot_func.pycode", "ot_solvers.pycode"))


def _load(name, path, package=None):
    if path.endswith(".pycode"):
        from importlib.machinery import SourcelessFileLoader
        spec = importlib.util.spec_from_loader(name, SourcelessFileLoader(name, path))
    else:
        spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=None)
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_ot_solvers(lib_path=None, tag=None):
    """Returns the reference `ot_solvers` module.  By default it binds its own shipped libot.so (ot_func.py:10).

    lib_path: load THAT shared library in place of libot.so — the one-line swap of INTEGRATION.md section 5 done from the
    outside: `ctypes.cdll.LoadLibrary` is intercepted while the reference's ot_func.py executes, nothing in the
    reference is edited.  `tag` names the private package (one module instance per library).
    Source files under /root/reference are used when mounted, else the byte-compiled copies in oracle/_ref/pyref."""
    tag = tag or ("_spadot_ref_ot" if lib_path is None else "_spadot_ref_ot_" + str(abs(hash(lib_path))))
    if tag + ".ot_solvers" in sys.modules:
        return sys.modules[tag + ".ot_solvers"]
    if available():
        d, ext = _OT_DIR, ".py"
    elif compiled_available():
        d, ext = _PYREF, ".pycode"        # (not *.pyc: snapshot tools drop that extension)
        if lib_path is None:
            lib_path = os.path.join(os.path.dirname(_PYREF), "libot_ref.so")    # the reference's ot_func.cpp, compiled
    else:
        raise ImportError("neither /root/reference nor oracle/_ref/pyref is present")
    sys.modules.setdefault("ot", types.ModuleType("ot"))
    pkg = types.ModuleType(tag)
    pkg.__path__ = [d]
    sys.modules[tag] = pkg
    import ctypes
    orig = ctypes.cdll.LoadLibrary

    def redirect(path, *a, **k):
        if lib_path is not None and os.path.basename(str(path)) == "libot.so":
            return orig(lib_path, *a, **k)
        return orig(path, *a, **k)
    ctypes.cdll.LoadLibrary = redirect
    try:
        _load(tag + ".ot_func", os.path.join(d, "ot_func" + ext), tag)
    finally:
        ctypes.cdll.LoadLibrary = orig
    return _load(tag + ".ot_solvers", os.path.join(d, "ot_solvers" + ext), tag)


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


def load_cal_spatial_net():
    """Returns the reference's `_Cal_Spatial_Net` (SpaDOT/utils/_utils.py:52-100) as a callable.

    `_utils.py` imports scanpy / chi2comb / statsmodels at module level (absent here), so the function's own source
    text is cut out of the file with `ast` and compiled UNMODIFIED in a namespace that holds exactly the names it uses
    (pd, np, sp = scipy, sklearn with sklearn.neighbors loaded)."""
    import ast
    import numpy as np
    import pandas as pd
    import scipy as sp
    import scipy.sparse  # noqa: F401
    import sklearn
    import sklearn.neighbors  # noqa: F401
    path = os.path.join(REF_ROOT, "SpaDOT", "utils", "_utils.py")
    src = open(path).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "_Cal_Spatial_Net")
    code = compile(ast.Module(body=[fn], type_ignores=[]), path, "exec")
    ns = dict(pd=pd, np=np, sp=sp, sklearn=sklearn)
    exec(code, ns)
    return ns["_Cal_Spatial_Net"]


def reference_edge_index(coords, k_cutoff, max_neigh=30):
    """edge_index exactly as the reference's train loop builds it: `_Cal_Spatial_Net` on a duck-typed AnnData, then
    `dense_to_sparse(torch.tensor(adata.uns['adj']))` (utils/_train_utils.py:69-72) — PyG's dense_to_sparse of a 2-D
    matrix is `adj.nonzero().t()`, i.e. the non-zeros in row-major order (restated here: torch_geometric is absent)."""
    import contextlib
    import io
    import types
    import numpy as np
    import pandas as pd
    import scipy.sparse
    import torch
    n = coords.shape[0]
    names = [f"spot{i}" for i in range(n)]
    adata = types.SimpleNamespace(obsm={"spatial": np.asarray(coords)}, obs=pd.DataFrame(index=names), n_obs=n, uns={},
                                  layers={"counts": scipy.sparse.csr_matrix((n, 2))}, var=pd.DataFrame(index=["g0", "g1"]))
    with contextlib.redirect_stdout(io.StringIO()) as out:
        load_cal_spatial_net()(adata, k_cutoff=k_cutoff, max_neigh=max_neigh)
    adj = torch.tensor(np.asarray(adata.uns["adj"]), dtype=torch.float64)
    return adj.nonzero().t().contiguous().numpy(), out.getvalue()
