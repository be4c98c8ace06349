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
