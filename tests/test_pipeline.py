"""Input pipeline: reference src/datasets/modelnet.py ``ply_reader`` (:15-29) and the centre /
scale block of ``FewShotSubModelNet.__getitem__`` (:66-69).

CPU: the oracle restatements and the library's host-side PLY parser against golden outputs of the reference's own
code (tests/golden/make_golden_pipeline.py). GPU: the normalisation kernel bit-for-bit against the same golden
outputs and against the oracle on larger seeded batches."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN


def _norm_cases():
    z = np.load(os.path.join(GOLDEN, "pipeline_norm.npz"))
    return {k[2:]: (z[k], z["y_" + k[2:]]) for k in z.files if k.startswith("x_")}


def test_oracle_normalisation_matches_reference_outputs_bitwise():
    from oracle import cpu_oracle as O
    for name, (x, y) in _norm_cases().items():
        got = O.normalize_clouds(x[None])[0]
        assert np.array_equal(got, y), f"{name}: max diff {np.abs(got - y).max()}"


@pytest.mark.parametrize("which", ["a", "b"])
def test_ply_parsers_match_reference_reader(which):
    from oracle import cpu_oracle as O
    import fpsg_b200
    want = np.load(os.path.join(GOLDEN, "pipeline_ply.npz"))[which + "_xyz"]
    path = os.path.join(GOLDEN, f"cloud_{which}.ply")
    assert np.array_equal(O.ply_reader_py(path), want)       # the line-for-line restatement
    got = fpsg_b200.ply_reader(path)                         # the product's C parser (host code, no GPU needed)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got, want)
    assert np.array_equal(fpsg_b200.ply_reader(path, max_points=100), want[:100])


def test_ply_parser_errors(tmp_path):
    import fpsg_b200
    with pytest.raises(RuntimeError, match="cannot open"):
        fpsg_b200.ply_reader(str(tmp_path / "missing.ply"))
    p = tmp_path / "short.ply"
    p.write_text("ply\nelement vertex 3\nend_header\n0 0 0\n1 1 1\n")
    with pytest.raises(RuntimeError, match="3 vertices announced, 2 found"):
        fpsg_b200.ply_reader(str(p))
    p = tmp_path / "wide.ply"
    p.write_text("ply\nelement vertex 1\nend_header\n0 0 0 1 1 1\n")
    with pytest.raises(RuntimeError, match="more than 3"):
        fpsg_b200.ply_reader(str(p))
    p = tmp_path / "nohdr.ply"
    p.write_text("ply\nelement vertex 1\n0 0 0\n")
    with pytest.raises(RuntimeError, match="end_header"):
        fpsg_b200.ply_reader(str(p))


def test_normalise_refuses_cpu_tensors():
    import fpsg_b200
    with pytest.raises(RuntimeError, match="CUDA"):
        fpsg_b200.normalize_point_clouds(torch.zeros(1, 4, 3))


@pytest.mark.gpu
def test_gpu_normalisation_matches_reference_outputs_bitwise():
    import fpsg_b200
    for name, (x, y) in _norm_cases().items():
        got = fpsg_b200.normalize_point_clouds(torch.from_numpy(x).cuda()).cpu().numpy()
        assert np.array_equal(got, y), f"{name}: max diff {np.abs(got - y).max()}"


@pytest.mark.gpu
@pytest.mark.parametrize("B,N", [(1, 1), (3, 2), (37, 2048), (5, 2049), (2, 5000), (4, 777)])
def test_gpu_normalisation_matches_oracle_bitwise(B, N):
    from oracle import cpu_oracle as O
    import fpsg_b200
    g = torch.Generator().manual_seed(B * 1000 + N)
    x = torch.randn(B, N, 3, generator=g) * torch.tensor([1.0, 30.0, 0.02]) + torch.tensor([5.0, -200.0, 0.1])
    got = fpsg_b200.normalize_point_clouds(x.cuda()).cpu().numpy()
    want = O.normalize_clouds(x.numpy())
    assert np.array_equal(got, want, equal_nan=True), f"max diff {np.nanmax(np.abs(got - want))}"
    if N > 1:
        norms = np.sqrt((got.astype(np.float64) ** 2).sum(-1))
        assert np.allclose(norms.max(1), 1.0, atol=1e-6) and np.allclose(got.mean(1), 0.0, atol=1e-4)


@pytest.mark.gpu
def test_pipeline_feeds_the_chamfer_path():
    """file -> parser -> GPU normalisation -> Chamfer loss, all through the product."""
    import fpsg_b200
    a = fpsg_b200.ply_reader(os.path.join(GOLDEN, "cloud_b.ply"))
    ref = fpsg_b200.normalize_point_clouds(torch.from_numpy(a).cuda()[None])
    syn = torch.tanh(torch.randn(1, 2048, 3, generator=torch.Generator().manual_seed(1))).cuda().requires_grad_(True)
    loss = fpsg_b200.chamfer_loss(syn, ref).sum()
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(syn.grad).all()
