"""Drop-in shims that let the reference's source run unchanged on top of fpsg_b200.

``install()`` does two things:
  * registers ``kaolin``, ``kaolin.metrics`` and ``kaolin.metrics.pointcloud`` in ``sys.modules`` (unless a real
    Kaolin is importable and ``force`` is False), so that ``from kaolin.metrics.pointcloud import
    chamfer_distance`` (reference src/models/few_shot.py:13) resolves to the B200 implementation;
  * if the reference's ``dgcnn.model`` module is already imported, rebinds its module-level ``knn`` to
    ``fpsg_b200.knn`` -- ``get_graph_feature`` looks ``knn`` up at call time (src/dgcnn/model.py:28). If the module is
    imported later, call ``patch_dgcnn_knn()`` again.

Alternatively put ``fpsg_b200/shims`` on ``PYTHONPATH``: the ``kaolin`` package in this directory is a regular
importable package.
"""
from __future__ import annotations

import importlib
import importlib.util
import sys


def install(force: bool = False, patch_dgcnn: bool = True) -> None:
    have_real = False
    if not force:
        loaded = sys.modules.get("kaolin")
        if loaded is not None:
            # something called kaolin is already imported: it is "real" unless it is this shim
            have_real = not getattr(loaded, "__file__", "").startswith(__path__[0])
        else:
            try:
                have_real = importlib.util.find_spec("kaolin") is not None
            except (ImportError, ValueError):
                have_real = False
    if force or not have_real:
        from .kaolin import metrics as _metrics  # noqa: F401
        from . import kaolin as _kaolin
        from .kaolin.metrics import pointcloud as _pc
        sys.modules["kaolin"] = _kaolin
        sys.modules["kaolin.metrics"] = _metrics
        sys.modules["kaolin.metrics.pointcloud"] = _pc
    if patch_dgcnn:
        patch_dgcnn_knn()


def install_neuralnet_pytorch(force: bool = False) -> bool:
    """Registers the ``neuralnet_pytorch`` shim (``metrics.emd_loss`` with ``sinkhorn=True`` -> ``fpsg_b200.emd_sinkhorn``;
    reference src/models/utils.py:9-13) unless the real package is importable and ``force`` is False. Returns True if the
    shim is in place. Kept separate from ``install()``: the metric is evaluation-only and parity-unpinned."""
    if not force:
        if "neuralnet_pytorch" in sys.modules:
            return getattr(sys.modules["neuralnet_pytorch"], "__file__", "").startswith(__path__[0])
        try:
            if importlib.util.find_spec("neuralnet_pytorch") is not None:
                return False
        except (ImportError, ValueError):
            pass
    from . import neuralnet_pytorch as _nnt
    from .neuralnet_pytorch import metrics as _m
    sys.modules["neuralnet_pytorch"] = _nnt
    sys.modules["neuralnet_pytorch.metrics"] = _m
    return True


def patch_dgcnn_knn(module=None, graph_feature: bool = False) -> bool:
    """Rebinds ``knn`` in the reference's dgcnn.model (already imported, or given). Returns True if patched.
    With ``graph_feature=True`` the module's ``get_graph_feature`` (src/dgcnn/model.py:23-42, looked up as a module
    global by ``DGCNNfeat.forward``, :62-74) is replaced by the fused B200 version as well."""
    from ..knn import knn
    from ..graph_feature import get_graph_feature
    mods = [module] if module is not None else [m for n, m in list(sys.modules.items())
                                                if n in ("dgcnn.model", "src.dgcnn.model") and m is not None]
    done = False
    for m in mods:
        if hasattr(m, "knn") and hasattr(m, "get_graph_feature"):
            m.knn = knn
            if graph_feature:
                m.get_graph_feature = get_graph_feature
            done = True
    return done
