"""``kaolin.metrics.pointcloud`` as the reference consumes it (Kaolin 0.9.0 API):

    chamfer_distance(p1, p2, w1=1., w2=1.) -> Tensor (B,)      squared-L2, w1*mean + w2*mean
    sided_distance(p1, p2)                  -> (dist (B,N), idx (B,N) int64)

Call sites in the reference: src/models/few_shot.py:57 (bound as ``pc_metric``), :110, :117, :167
(``self.pc_metric(syn_pc, ref_pc).sum()``).
"""
import os
import sys

try:
    from fpsg_b200.chamfer import chamfer_loss as _chamfer_loss, sided_distance as _sided_distance
except ImportError:  # used straight from PYTHONPATH=fpsg_b200/shims without the repo root on sys.path
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__)))))))
    from fpsg_b200.chamfer import chamfer_loss as _chamfer_loss, sided_distance as _sided_distance


def chamfer_distance(p1, p2, w1=1., w2=1., squared=True):
    if not squared:
        raise NotImplementedError("only the squared Chamfer distance (Kaolin 0.9.0 default) is provided")
    return _chamfer_loss(p1, p2, w1, w2)


def sided_distance(p1, p2):
    return _sided_distance(p1, p2)
