"""Minimal stand-in for the ``kaolin`` package: only ``kaolin.metrics.pointcloud``, which is all the reference
imports (src/models/few_shot.py:13). Backed by fpsg_b200's sm_100a kernels."""
from . import metrics  # noqa: F401

__version__ = "0.9.0+fpsg_b200"
