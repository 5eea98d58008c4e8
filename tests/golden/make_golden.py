"""Generates the committed golden fixtures in tests/golden/. Run once in the build container:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

* knn_*.npz   -- inputs and the outputs of the REAL reference function ``knn`` imported from
                 /root/reference/src/dgcnn/model.py (lines 13-20), run on CPU. These pin the kNN oracle.
* chamfer_*.npz -- inputs and the outputs of the pure-PyTorch difference-form Chamfer (forward via
                 min/argmin, backward via torch.autograd). The reference's own Chamfer is Kaolin 0.9.0,
                 which is absent here, so these pin the *formulation* (reference src/models/utils.py:98-109
                 + the Kaolin semantics restated in oracle/fpsg_oracle.c), not Kaolin's binary.
/root/reference does not exist on the GPU box; nothing at test time reads it -- only this script does.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.dont_write_bytecode = True

from oracle import torch_formulation as TF  # noqa: E402


def load_reference_knn():
    spec = importlib.util.spec_from_file_location("ref_dgcnn_model", "/root/reference/src/dgcnn/model.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.knn


def unit_ball_cloud(g, B, N):
    p = torch.randn(B, N, 3, generator=g)
    p = p - p.mean(1, keepdim=True)
    return p / p.norm(dim=2).amax(1)[:, None, None]


def main():
    torch.set_num_threads(1)  # fixed reduction order for the record
    knn = load_reference_knn()
    g = torch.Generator().manual_seed(20261018)

    cases = {
        # name: (x, k)
        "knn_c3_k20": (unit_ball_cloud(g, 2, 256).transpose(1, 2).contiguous(), 20),
        "knn_c64_k20": (torch.randn(2, 64, 128, generator=g), 20),
        "knn_c128_k20": (torch.randn(1, 128, 96, generator=g), 20),
        "knn_c3_k1": (unit_ball_cloud(g, 1, 64).transpose(1, 2).contiguous(), 1),
        "knn_c3_kN": (unit_ball_cloud(g, 1, 24).transpose(1, 2).contiguous(), 24),
        "knn_c3_far": ((unit_ball_cloud(g, 1, 128) + 100.0).transpose(1, 2).contiguous(), 20),
    }
    for name, (x, k) in cases.items():
        with torch.no_grad():
            idx = knn(x, k)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x.numpy(), k=np.int64(k), idx=idx.numpy())
        print(name, tuple(x.shape), k, tuple(idx.shape))

    chamfer_cases = {
        "chamfer_small": (torch.tanh(0.5 * torch.randn(2, 64, 3, generator=g)), unit_ball_cloud(g, 2, 48)),
        "chamfer_ragged": (torch.tanh(0.5 * torch.randn(3, 33, 3, generator=g)), unit_ball_cloud(g, 3, 130)),
    }
    for name, (p1, p2) in chamfer_cases.items():
        p1 = p1.clone().requires_grad_(True)
        p2 = p2.clone().requires_grad_(True)
        d1, i1 = TF.sided_distance_torch(p1, p2)
        d2, i2 = TF.sided_distance_torch(p2, p1)
        loss = TF.chamfer_distance_torch(p1, p2)
        loss.sum().backward()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), p1=p1.detach().numpy(), p2=p2.detach().numpy(),
                            dist1=d1.detach().numpy(), dist2=d2.detach().numpy(), idx1=i1.numpy(), idx2=i2.numpy(),
                            loss=loss.detach().numpy(), grad_p1=p1.grad.numpy(), grad_p2=p2.grad.numpy())
        print(name, tuple(p1.shape), tuple(p2.shape))


if __name__ == "__main__":
    main()
