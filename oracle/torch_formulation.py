"""The reference's PyTorch formulation of the hot path, on CPU. TEST INFRASTRUCTURE / CPU BASELINE ONLY.

* ``chamfer_distance_torch``: the squared-L2, mean+mean Chamfer the reference obtains from
  ``kaolin.metrics.pointcloud.chamfer_distance`` (reference src/models/few_shot.py:13,57,110,117,167), written
  with the in-tree difference-form distance ``torch.pow(x - y, 2).sum(2)`` (reference src/models/utils.py:98-109)
  so that torch.autograd gives an independent check of the hand-derived backward, and so that bench.py can time
  "the reference's PyTorch CPU path" named by BASELINE.md section 3.
* ``knn_torch``: restatement of reference src/dgcnn/model.py:13-20 (same three tensor expressions, same topk).
  tests/golden/make_golden.py imports the real function from /root/reference to pin this one.
"""
from __future__ import annotations

import torch


def sided_distance_torch(p1: torch.Tensor, p2: torch.Tensor):
    """(dist (B,N), idx (B,N)): min / argmin over p2 of the difference-form squared distance."""
    d = torch.pow(p1[:, :, None, :] - p2[:, None, :, :], 2).sum(-1)  # utils.py:106-109, batched
    dist, idx = d.min(dim=2)
    return dist, idx


def chamfer_distance_torch(p1: torch.Tensor, p2: torch.Tensor, w1: float = 1.0, w2: float = 1.0,
                           chunk: int | None = None) -> torch.Tensor:
    """(B,) = w1*mean_i min_j d + w2*mean_j min_i d. ``chunk`` bounds the (chunk,N,M) temporary."""
    B = p1.shape[0]
    step = chunk or B
    out = []
    for s in range(0, B, step):
        a, b = p1[s:s + step], p2[s:s + step]
        d = torch.pow(a[:, :, None, :] - b[:, None, :, :], 2).sum(-1)
        out.append(w1 * d.min(dim=2)[0].mean(dim=-1) + w2 * d.min(dim=1)[0].mean(dim=-1))
    return torch.cat(out) if len(out) > 1 else out[0]


def knn_torch(x: torch.Tensor, k: int) -> torch.Tensor:
    inner = -2 * torch.matmul(x.transpose(2, 1), x)             # model.py:14
    xx = torch.sum(x ** 2, dim=1, keepdim=True)                 # model.py:15
    pairwise_distance = -xx - inner - xx.transpose(2, 1)        # model.py:16
    return pairwise_distance.topk(k=k, dim=-1)[1]               # model.py:18


def get_graph_feature_torch(x: torch.Tensor, k: int = 20, idx: torch.Tensor | None = None) -> torch.Tensor:
    """Restatement of reference src/dgcnn/model.py:23-42 with the hard-coded ``torch.device('cuda')`` (:29)
    replaced by ``x.device`` so that it also runs on CPU (the reference function itself cannot: it allocates
    ``idx_base`` on 'cuda' unconditionally). Same tensor expressions otherwise; used as the autograd cross-check
    and CPU statement of SURVEY.md 8f row f1."""
    batch_size, num_dims, num_points = x.size()
    x = x.view(batch_size, -1, num_points)                                            # model.py:25
    if idx is None:
        idx = knn_torch(x, k=k)                                                      # model.py:27-28
    idx_base = torch.arange(0, batch_size, device=x.device).view(-1, 1, 1) * num_points  # model.py:31
    idx = (idx + idx_base).view(-1)                                                   # model.py:33-34
    x = x.transpose(2, 1).contiguous()                                                # model.py:36
    feature = x.view(batch_size * num_points, -1)[idx, :]                             # model.py:37
    feature = feature.view(batch_size, num_points, k, num_dims)                       # model.py:38
    x = x.view(batch_size, num_points, 1, num_dims).repeat(1, 1, k, 1)                # model.py:39
    return torch.cat((feature - x, x), dim=3).permute(0, 3, 1, 2).contiguous()        # model.py:41
