"""Caption feature bank builder (reference: generate_caption_text_features.py:72-96).

The reference pushes every tokenised caption through the CLIP text tower with `if_sequence=True`, keeps the EOT
row, L2-normalises it and pickles the stacked tensor; `Caption_distill_double.py:35-36` loads that pickle at import
and the retrieval step (T:444-448) reads it.  Here the same rows come from `TextEncoder(ids, None, if_embedding=False)`
— only the positions up to the last EOT of the batch are run (exact under the causal mask) — normalised by
`lecb_l2norm_rows`, and the file written is the same single-object pickle of a CPU tensor."""
from __future__ import annotations

import pickle

import torch

from . import ops


@torch.no_grad()
def build_caption_bank(text_encoder, captions: torch.Tensor, batch_size: int = 4096, dtype=torch.float16) -> torch.Tensor:
    """captions int64 [N,77] token ids (SOT ... EOT, zero padded) -> unit-norm features [N,D] on the encoder's device."""
    dev = text_encoder.positional_embedding.device
    out = []
    for i in range(0, captions.shape[0], batch_size):
        ids = captions[i:i + batch_size].to(dev)
        feats = text_encoder(ids, None, if_embedding=False)                    # [n, D] fp32 at the EOT rows
        out.append(ops.l2norm_rows(feats).to(dtype))
    return torch.cat(out, 0)


def save_caption_bank(path: str, bank: torch.Tensor) -> None:
    """Same on-disk format as generate_caption_text_features.py:95-96: pickle.dump of one CPU tensor."""
    with open(path, "wb") as f:
        pickle.dump(bank.detach().cpu(), f)


def load_caption_bank(path: str, device="cuda") -> torch.Tensor:
    """What Caption_distill_double.py:35-36 does at import (`pickle.load(f).cuda()`), as fp16 for `DenseCLIPB200(caption_bank=)`."""
    with open(path, "rb") as f:
        bank = pickle.load(f)
    return bank.to(device=device, dtype=torch.float16).contiguous()
