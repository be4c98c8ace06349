"""oracle/pil_resize.py (numpy restatement of Pillow's 8-bit resize, the reference's test-time `Resize`) against Pillow
itself: bit-exact, up- and down-scaling, both filters the reference's INTERPOLATION_MODES can select for CLIP inputs."""
import numpy as np
import pytest

from oracle import pil_resize as PR

PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402

CASES = [
    # (in_h, in_w) -> (out_h, out_w): the whole image to the network input, windows of several scales (upscaling and mild
    # downscaling), extreme aspect ratios, identity along one axis
    ((375, 500), (448, 448)), ((480, 640), (224, 224)), ((187, 250), (448, 448)), ((93, 125), (448, 448)),
    ((75, 250), (224, 224)), ((1024, 683), (448, 448)), ((224, 300), (224, 224)), ((31, 17), (64, 96)),
]


@pytest.mark.parametrize("filt", ["bicubic", "bilinear"])
@pytest.mark.parametrize("src,dst", CASES)
def test_resize_matches_pillow_bit_for_bit(src, dst, filt):
    rng = np.random.default_rng(src[0] * 1000 + dst[1])
    img = rng.integers(0, 256, size=(src[0], src[1], 3), dtype=np.uint8)
    img[: src[0] // 3] = np.where(rng.random((src[0] // 3, src[1], 1)) < 0.5, 0, 255)      # saturating edges: the clamp matters
    want = np.asarray(Image.fromarray(img).resize((dst[1], dst[0]), resample=getattr(Image, filt.upper())))
    got = PR.resize_u8(img, dst[0], dst[1], filt)
    assert got.dtype == np.uint8 and got.shape == want.shape
    assert np.array_equal(got, want), f"{int((got != want).sum())} of {got.size} bytes differ (max {int(np.abs(got.astype(int) - want).max())})"


def test_transform_matches_torchvision():
    tv = pytest.importorskip("torchvision")
    import torch
    import torchvision.transforms as T
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=(333, 500, 3), dtype=np.uint8)
    mean, std = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)     # CLIP's PIXEL_MEAN / PIXEL_STD
    tfm = T.Compose([T.Resize((448, 448), interpolation=T.InterpolationMode.BICUBIC), T.ToTensor(), T.Normalize(mean, std)])
    want = tfm(Image.fromarray(img)).numpy()
    got = PR.test_transform(img, (448, 448), mean, std)
    assert np.abs(got - want).max() <= 1e-6


@pytest.mark.parametrize("filt", ["bicubic", "bilinear"])
@pytest.mark.parametrize("n_in,n_out", [(375, 448), (640, 224), (93, 448), (1024, 448), (224, 224), (17, 96), (500, 7), (3, 448)])
def test_library_taps_equal_the_oracle(n_in, n_out, filt):
    """lecb_resize_plan (host code of liblecb.so, the taps a GPU resize will consume) == oracle/pil_resize.precompute_coeffs,
    integer for integer; and applying the library's plan along one axis reproduces Pillow's bytes."""
    from lecb200 import windows
    bounds, coeffs = windows.resize_plan(n_in, n_out, filt)
    want_b, want_k = PR.precompute_coeffs(n_in, 0.0, float(n_in), n_out, filt)
    assert coeffs.shape == want_k.shape and np.array_equal(bounds, want_b) and np.array_equal(coeffs, want_k)
    rng = np.random.default_rng(n_in + n_out)
    img = rng.integers(0, 256, size=(n_in, 5, 3), dtype=np.uint8)
    acc = np.full((n_out, 5, 3), 1 << 21, dtype=np.int64)
    for t in range(coeffs.shape[1]):
        idx = np.minimum(bounds[:, 0] + t, n_in - 1)                 # taps past the count are zero: the index is irrelevant
        acc += coeffs[:, t, None, None].astype(np.int64) * img[idx].astype(np.int64)
    got = np.clip(acc >> 22, 0, 255).astype(np.uint8)
    want = np.asarray(Image.fromarray(img).resize((5, n_out), resample=getattr(Image, filt.upper())))
    assert np.array_equal(got, want)


def test_library_plan_rejects_bad_arguments():
    from lecb200 import LecbError, _lib, windows
    assert _lib.lib.lecb_resize_ksize(0, 10, 1) < 0 and _lib.lib.lecb_resize_ksize(10, 10, 7) < 0
    assert _lib.lib.lecb_resize_plan(10, 10, 1, 0, 0, 5) == -1 and b"null" in _lib.lib.lecb_last_error()
    with pytest.raises(LecbError):
        windows.resize_plan(0, 4)
