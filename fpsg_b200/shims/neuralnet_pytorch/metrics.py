"""``neuralnet_pytorch.metrics.emd_loss`` as the reference calls it (reference src/models/utils.py:12-13):

    emd_loss(pc1, pc2, reduce='sum', sinkhorn=True) -> Tensor (B,)     (the reference then takes .sum(), few_shot.py:168)

With ``sinkhorn=True`` the original hands the clouds to ``geomloss.SamplesLoss()`` and returns its per-pair values --
the debiased Sinkhorn divergence with geomloss' defaults -- which is what ``fpsg_b200.emd_sinkhorn`` implements
(parity unpinned: neither package is available, see fpsg_b200/emd.py). The ``sinkhorn=False`` branch (an auction-based
CUDA EMD) is not used by the reference and is not provided."""
import os
import sys

try:
    from fpsg_b200.emd import emd_sinkhorn as _emd_sinkhorn
except ImportError:  # used straight from PYTHONPATH=fpsg_b200/shims without the repo root on sys.path
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
    from fpsg_b200.emd import emd_sinkhorn as _emd_sinkhorn


def emd_loss(xyz1, xyz2, reduce='mean', sinkhorn=False):
    if reduce not in ('mean', 'sum'):
        raise AssertionError('Unknown reduce method')
    if not sinkhorn:
        raise NotImplementedError("only the sinkhorn=True branch (the one the reference uses) is provided")
    return _emd_sinkhorn(xyz1, xyz2)
