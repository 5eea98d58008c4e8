"""Float64 CPU statement of the Sinkhorn EMD the product implements (fpsg_b200/csrc/emd_sinkhorn.cu).

TEST INFRASTRUCTURE, not product code. PARITY UNPINNED: the reference computes this metric through
``neuralnet_pytorch.metrics.emd_loss(..., sinkhorn=True)`` (reference src/models/utils.py:9-13, used at
src/models/few_shot.py:168), which delegates to ``geomloss.SamplesLoss()``; neither package is in the reference tree nor
in this image, so nothing here could be checked against the reference's own output. The recurrences below are geomloss'
published default (debiased Sinkhorn divergence, p = 2, symmetrised log-domain updates, epsilon-scaling) as recalled."""
from __future__ import annotations

import numpy as np
from scipy.special import logsumexp


def _softmin(eps, C, h):
    """-eps * log sum_j exp(h_j - C_ij / eps), row-wise."""
    return -eps * logsumexp(h[None, :] - C / eps, axis=1)


def sinkhorn_divergence(x, y, eps_schedule, debias=True):
    """One pair: x (N,3), y (M,3) -> (value, b_i (N,), a_j (M,)); float64 throughout."""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    N, M = len(x), len(y)
    Cxy = 0.5 * ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)
    Cyx = Cxy.T
    Cxx = 0.5 * ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    Cyy = 0.5 * ((y[:, None, :] - y[None, :, :]) ** 2).sum(-1)
    la, lb = np.full(N, -np.log(N)), np.full(M, -np.log(M))
    eps = eps_schedule[0]
    a_i, b_j = _softmin(eps, Cxx, la), _softmin(eps, Cyy, lb)
    a_j, b_i = _softmin(eps, Cyx, la), _softmin(eps, Cxy, lb)
    for eps in eps_schedule:
        at_i = _softmin(eps, Cxx, la + a_i / eps)
        bt_j = _softmin(eps, Cyy, lb + b_j / eps)
        at_j = _softmin(eps, Cyx, la + b_i / eps)
        bt_i = _softmin(eps, Cxy, lb + a_j / eps)
        a_i, b_j = 0.5 * (a_i + at_i), 0.5 * (b_j + bt_j)
        a_j, b_i = 0.5 * (a_j + at_j), 0.5 * (b_i + bt_i)
    eps = eps_schedule[-1]
    a_i, b_j = _softmin(eps, Cxx, la + a_i / eps), _softmin(eps, Cyy, lb + b_j / eps)
    a_j, b_i = _softmin(eps, Cyx, la + b_i / eps), _softmin(eps, Cxy, lb + a_j / eps)
    if debias:
        val = (b_i - a_i).mean() + (a_j - b_j).mean()
    else:
        val = b_i.mean() + a_j.mean()
    return val, b_i, a_j


def emd_sinkhorn(p1, p2, eps_schedule, debias=True):
    """Batched: (B,N,3), (B,M,3) -> (B,) float64."""
    return np.array([sinkhorn_divergence(a, b, eps_schedule, debias)[0] for a, b in zip(np.asarray(p1), np.asarray(p2))])
