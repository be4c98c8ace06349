"""Test-time score fusion kernels (lecb_block_fuse, lecb_cooc_adjust) vs the reference-generated golden and the oracle."""
import numpy as np
import pytest
import torch

from oracle import restatement as R

from . import _cases as C

pytestmark = pytest.mark.gpu


def test_fusion_kernels_match_reference_golden():
    from lecb200 import postprocess as PP
    g = C.load("fusion.npz")
    data, sims, out = (torch.from_numpy(g[k]).cuda() for k in ("data", "sims", "output"))
    np.testing.assert_allclose(PP.fuse(data, sims).cpu().numpy(), g["fuse"], atol=2e-6)
    np.testing.assert_allclose(PP.fuse(data, sims, threshold=0.5).cpu().numpy(), g["fuse_t05"], atol=2e-6)
    np.testing.assert_allclose(PP.fuse6(data, sims).cpu().numpy(), g["fuse6"], atol=2e-6)
    p = PP.normalized_cooccurrence(g["adj"], g["nums"])
    np.testing.assert_allclose(PP.adjust_predictions(out, p, 0.5).cpu().numpy(), g["adjusted"], atol=2e-6)
    want = R.aggregate_blocks(out.cpu(), data.cpu(), 0.3, 1.4)
    np.testing.assert_allclose(PP.aggregate_blocks(out, data).cpu().numpy(), want.numpy(), atol=1e-6)


@pytest.mark.parametrize("b,nb,k", [(1, 1, 2), (3, 7, 6), (5, 116, 80), (2, 300, 128)])
def test_fusion_kernels_edge_shapes(b, nb, k):
    """Single window (max == min), tiny / maximal class counts, more windows than warps."""
    from lecb200 import postprocess as PP
    gen = torch.Generator().manual_seed(b * 100 + nb)
    data = torch.rand((b, nb, k), generator=gen) - 0.3
    sims = torch.rand((b, nb, 5), generator=gen)
    base = torch.rand((b, k), generator=gen)
    for fn_gpu, fn_ref in ((PP.fuse, R.fuse), (PP.fuse6, R.fuse6)):
        got = fn_gpu(data.cuda(), sims.cuda()).cpu()
        torch.testing.assert_close(got, fn_ref(data, sims).float(), atol=3e-6, rtol=1e-5)
    got = PP.aggregate_blocks(base.cuda(), data.cuda(), threshold=0.1, weight=1.4).cpu()
    torch.testing.assert_close(got, R.aggregate_blocks(base, data, 0.1, 1.4), atol=1e-6, rtol=1e-6)


def test_fusion_rejects_cpu_and_bad_shapes():
    from lecb200 import LecbError, postprocess as PP
    with pytest.raises(LecbError):
        PP.aggregate_blocks(torch.zeros((2, 80)), torch.zeros((2, 4, 80)))
    with pytest.raises(LecbError):
        PP.aggregate_blocks(torch.zeros((2, 200), device="cuda"), torch.zeros((2, 4, 200), device="cuda"))     # K > 128
