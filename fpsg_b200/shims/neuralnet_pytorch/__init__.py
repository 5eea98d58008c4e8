"""The sliver of ``neuralnet_pytorch`` the reference imports (reference src/models/utils.py:9:
``from neuralnet_pytorch.metrics import emd_loss``), backed by fpsg_b200. See ``metrics``."""
from . import metrics  # noqa: F401
