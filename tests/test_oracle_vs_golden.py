"""Pin oracle/restatement.py against the reference-generated fixtures (CPU; no GPU, no reference tree).

Tolerances: the restatement runs the same fp32 math as the reference in a different op order
(e.g. single-query attention instead of full MHA), so agreement is ~1e-5 relative; 2e-4 absolute on
logits (range ±4) is the gate here, 50x tighter than the 1e-2 product tolerance."""
import numpy as np
import pytest
import torch

from oracle import restatement as R

from . import _cases as C

ATOL = 2e-4


@pytest.mark.parametrize("tag", ["tiny", "small", "rn50_224"])
def test_head_matches_reference(tag):
    c = C.head_case(tag)
    g = c["gold"]
    for ev in c["modes"]:
        sfx = "_ev" if ev else ""
        with torch.no_grad():
            out = R.dense_clip_test(c["sd"], c["arch"], c["image"], c["pl_state"], c["tokens"],
                                    use_evidence=ev, bank=c["bank"])
        for name, t in zip(("logits", "logits_local", "neg_map", "pos_map", "topk_scores"), out):
            np.testing.assert_allclose(t.numpy(), g[name + sfx], atol=ATOL, rtol=1e-4, err_msg=f"{tag}{sfx}:{name}")


def test_prompt_features_match_reference():
    c = C.head_case("tiny")
    g = c["gold"]
    with torch.no_grad():
        f = R.prompt_features(c["sd"], c["pl_state"], c["tokens"], c["arch"].transformer_heads, use_evidence=True)
    for t, name in zip(f, ("text_features_ev", "text_features_neg_ev", "text_features_evidence_ev")):
        t = t / t.norm(dim=-1, keepdim=True)
        np.testing.assert_allclose(t.numpy(), g[name], atol=1e-5)


@pytest.mark.parametrize("tag", ["tiny", "rn50"])
def test_train_path_matches_reference(tag):
    c = C.train_case(tag)
    g = c["gold"]
    for ev in (False, True):
        for loss_name in ("ranking", "asl"):
            pl = {k: (v.clone().requires_grad_(True) if k.startswith("ctx") else v) for k, v in c["pl_state"].items()}
            out = R.dense_clip_train(c["sd"], c["arch"], c["captions"], pl, c["tokens"], use_evidence=ev)
            if loss_name == "ranking":
                loss = R.ranking_loss(out[0], c["labels"], 1.0, 1.0) + R.ranking_loss(out[1], c["labels"], 1.0, 1.0)
            else:
                loss = R.asl_loss(out[0], c["labels"]) + R.asl_loss(out[1], c["labels"])
            loss.backward()
            sfx = ("_ev" if ev else "") + "_" + loss_name
            assert abs(loss.item() - float(g["loss" + sfx])) < 1e-4 * max(1.0, abs(float(g["loss" + sfx])))
            for pname in ("ctx", "ctx_double", "ctx_evidence"):
                gref = g[f"grad_{pname}" + sfx]
                got = pl[pname].grad
                if bool(g[f"gradnone_{pname}" + sfx]):
                    assert got is None or float(got.abs().max()) == 0.0
                else:
                    scale = max(np.abs(gref).max(), 1e-8)
                    np.testing.assert_allclose(got.numpy() / scale, gref / scale, atol=2e-4, err_msg=f"{tag}{sfx}:{pname}")
            if loss_name == "ranking":
                s2 = "_ev" if ev else ""
                np.testing.assert_allclose(out[0].detach().numpy(), g["logits" + s2], atol=ATOL)
                np.testing.assert_allclose(out[1].detach().numpy(), g["logits_local" + s2], atol=ATOL)
                np.testing.assert_allclose(out[2].detach().numpy()[:, :2], g["seq_feats" + s2], atol=1e-5)
                np.testing.assert_allclose(out[3].detach().numpy(), g["text_features" + s2], atol=1e-5)


def test_losses_match_reference():
    g = C.load("losses.npz")
    x0, y, yp = (torch.from_numpy(g[k]) for k in ("x", "y", "y_partial"))
    for name, fn in (("ranking_s1", lambda a: R.ranking_loss(a, y, 1.0, 1.0)),
                     ("ranking_s2", lambda a: R.ranking_loss(a, y)),
                     ("asl", lambda a: R.asl_loss(a, y)),
                     ("dualcoop", lambda a: R.dualcoop_loss(a, yp))):
        a = x0.clone().requires_grad_(True)
        loss = fn(a)
        loss.backward()
        assert abs(loss.item() - float(g["loss_" + name])) < 1e-5 * max(1.0, abs(float(g["loss_" + name]))), name
        np.testing.assert_allclose(a.grad.numpy(), g["grad_" + name], atol=1e-6, rtol=1e-5, err_msg=name)
        # U:86 scales its argument in place; the restatement must not
        np.testing.assert_array_equal(a.detach().numpy(), g["x"])


def test_map_matches_reference():
    g = C.load("map.npz")
    got = R.mean_average_precision(g["targets"], g["scores"])
    assert abs(got - float(g["mAP"])) < 1e-9


@pytest.mark.parametrize("tag,arch_fn", [("tiny", lambda: C.synth.tiny_vit()), ("b16_224", lambda: C.synth.VITB16(224)),
                                         ("l14_224", lambda: C.synth.VITL14(224))])
def test_vit_global_matches_reference(tag, arch_fn):
    """The ViT path's global (class-token) feature is the reference VisionTransformer.forward output (M:259-276)."""
    g = C.load(f"vit_{tag}.npz")
    arch = arch_fn()
    sd = C.synth.clip_state_dict(arch, 0)
    img = C.synth.images(int(g["batch"]), arch.image_resolution, int(g["seed"]))
    np.testing.assert_allclose(C.checksum(img), g["image_checksum"], rtol=1e-12, err_msg="RNG drift: images")
    with torch.no_grad():
        glob, local = R.vit_dense(sd, img, arch.vision_patch_size, arch.vision_width // 64)
    np.testing.assert_allclose(glob.numpy(), g["global_feat"], atol=ATOL, rtol=1e-4)
    grid = arch.image_resolution // arch.vision_patch_size
    assert tuple(local.shape) == (grid * grid, int(g["batch"]), arch.embed_dim)
    assert torch.isfinite(local).all()


def test_fusion_matches_reference():
    """Test-time fusion: the restatement vs the reference's own fuse / fuse6 / adjust_predictions."""
    g = C.load("fusion.npz")
    data, sims, out = (torch.from_numpy(g[k]) for k in ("data", "sims", "output"))
    np.testing.assert_allclose(R.fuse(data, sims).numpy(), g["fuse"], atol=1e-6)
    np.testing.assert_allclose(R.fuse(data, sims, 0.5).numpy(), g["fuse_t05"], atol=1e-6)
    np.testing.assert_allclose(R.fuse6(data, sims).numpy(), g["fuse6"], atol=1e-6)
    np.testing.assert_allclose(R.cooccurrence_adjust(out, g["adj"], g["nums"], 0.5).numpy(), g["adjusted"], atol=1e-6)
    # aggregation without re-weighting (T:655-662) == fuse with zero similarity and the variance factor divided out
    agg = R.aggregate_blocks(out, data, 0.3, 1.4)
    alpha, beta = data.max(1)[0], data.min(1)[0]
    want = 1.4 * torch.where(alpha > 0.3, alpha, beta) + out
    np.testing.assert_allclose(agg.numpy(), want.numpy(), atol=1e-6)
