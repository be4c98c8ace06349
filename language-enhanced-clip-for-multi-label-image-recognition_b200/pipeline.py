"""The reference's test-time multi-window scoring of ONE decoded image, end to end on the GPU.

Reference flow: `DatasetWrapperWithBlock.__getitem__` / `_transform_image` (dassl/data/data_manager.py:311-492) cuts the
image into sliding windows of every scale of `multi_scale` and pushes the whole image and each window through the dataset
transform (PIL `Resize(INPUT.SIZE, bicubic)`, `ToTensor`, `Normalize`) on the host; `Caption_distill_double.test`
(T:641-673) scores the image and every window batch with `model_inference` and combines them per class with the
max / min / threshold rule (`output_final = 1.4 * s_ag + output`, T:655-662), optionally after the co-occurrence adjustment of
the local scores (T:627-636).  Here: one upload of the decoded uint8 image, `lecb_crop_resize_u8` (Pillow's bytes) straight
into uint8 network inputs, `DenseCLIPB200.forward(image_u8, if_test=True)` in chunks, `lecb_block_fuse` / `lecb_cooc_adjust`.
Nothing runs on the host except the window geometry (integer arithmetic) and the cached resampling plans."""
from __future__ import annotations

from typing import Sequence

import torch

from . import postprocess, windows


@torch.no_grad()
def score_image_with_windows(model, img_u8: torch.Tensor, size: int, multi_scale: Sequence[int] = (2, 3, 4, 5), chunk: int = 256,
                             cooc_p: torch.Tensor = None, cooc_weight: float = 0.5, threshold: float = 0.3, weight: float = 1.4):
    """img_u8: decoded image, uint8 CUDA tensor [H,W,3] (what `read_image` yields before the transform).
    -> dict(output [1,K], output_pos [1,K] of the whole image; output_blocks / output_pos_blocks [1,NB,K], sims_blocks [1,NB,10];
            output_final, output_pos_final [1,K]) — the tensors T:641-673 builds per sample."""
    h, w = int(img_u8.shape[0]), int(img_u8.shape[1])
    wins = [windows.whole_image(h, w)] + [x for s in multi_scale for x in windows.sliding_windows(h, w, s)]
    batch, _ = windows.crop_resize(img_u8, wins, size, want_u8=True, want_f32=False)
    outs, outs_pos, sims = [], [], []
    for i in range(0, batch.shape[0], chunk):
        o, o_pos, _, _, sim = model(batch[i:i + chunk], if_test=True)
        if cooc_p is not None:                                     # T:631-636 / T:650-652 (`TEST.use_freq`)
            o_pos = postprocess.adjust_predictions(o_pos, cooc_p, cooc_weight)
        outs.append(o)
        outs_pos.append(o_pos)
        sims.append(sim)
    o, o_pos, sim = torch.cat(outs), torch.cat(outs_pos), torch.cat(sims)
    output, output_pos = o[:1].contiguous(), o_pos[:1].contiguous()
    blocks, blocks_pos = o[1:].unsqueeze(0).contiguous(), o_pos[1:].unsqueeze(0).contiguous()
    return {"output": output, "output_pos": output_pos, "output_blocks": blocks, "output_pos_blocks": blocks_pos,
            "sims": sim[:1], "sims_blocks": sim[1:].unsqueeze(0),
            "output_final": postprocess.aggregate_blocks(output, blocks, threshold, weight),
            "output_pos_final": postprocess.aggregate_blocks(output_pos, blocks_pos, threshold, weight),
            "n_windows": len(wins) - 1}
