"""fpsg_b200 -- B200 (sm_100a) implementation of FPSG's point-set-distance hot path: batched Chamfer distance
forward/backward and the DGCNN kNN graph, behind the reference's own call surfaces.

    from fpsg_b200 import chamfer_distance, chamfer_loss, sided_distance, knn
    import fpsg_b200.shims; fpsg_b200.shims.install()   # makes the reference source run unchanged
"""
from .chamfer import chamfer_distance, chamfer_loss, sided_distance
from .knn import knn

__all__ = ["chamfer_distance", "chamfer_loss", "sided_distance", "knn"]
__version__ = "0.1.0"
