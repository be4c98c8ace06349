"""__graft_entry__.smoke(): one small pass of the hot path on cuda:0, checked against the CPU oracle.

Case: the width-64 one-block-per-stage ModifiedResNet at 128x128, 4 images, 80 classes, evidence on.
The CUDA path runs end to end (stem, tcgen05 GEMMs / im2col convs, attnpool, text tower for the 240
prompts, dual-prompt head); the oracle (oracle/restatement.py) recomputes the image path in fp32 on the
CPU with the reference-generated prompt features from tests/golden/head_small.npz."""
import torch

from oracle import restatement as R

from . import _cases as C
from ._gpu_common import LOGIT_TOL, build_model


def run():
    import lecb200
    c = C.head_case("small")
    g = c["gold"]
    n0 = lecb200.launch_count()
    model = build_model(c, use_evidence=True)
    out = model(c["image"].cuda(), if_test=True)
    torch.cuda.synchronize()
    launches = lecb200.launch_count() - n0
    assert launches > 50, f"only {launches} lecb kernels launched"
    arch = c["arch"]
    with torch.no_grad():
        feat = R.rn_trunk(c["sd"], c["image"], arch.vision_layers)
        ref = R.head_test(R.attnpool_global(c["sd"], feat, arch.vision_width * 32 // 64), R.local_features(c["sd"], feat),
                          torch.from_numpy(g["text_features_ev"]), torch.from_numpy(g["text_features_neg_ev"]),
                          torch.from_numpy(g["text_features_evidence_ev"]))
    errs = [(o.float().cpu() - r).abs().max().item() for o, r in zip(out[:4], ref[:4])]
    print(f"smoke: {launches} lecb kernel launches; max abs err vs oracle: logits_ {errs[0]:.5f} "
          f"logits_local {errs[1]:.5f} neg_map {errs[2]:.5f} pos_map {errs[3]:.5f} (gate {LOGIT_TOL})")
    assert max(errs) <= LOGIT_TOL, errs
