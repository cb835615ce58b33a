"""GPU parity for ViT shapes beyond ViT-B/32: the any-T online-softmax attention kernel and the
generic (patch 14) im2col, i.e. the reference's default encoder ViT-L/14 (graph.py:98-104,
clip_feat_dim 768, 257 tokens) vs the fp32 torch oracle (A9)."""
import dataclasses

import numpy as np
import pytest
import torch

from oracle import hmsg_oracle as O
from holoagent_b200 import synth

pytestmark = pytest.mark.gpu


def _load(engine, shape, seed=0):
    sd = synth.make_vit_weights(shape, seed=seed)
    kw = dataclasses.asdict(shape)
    engine.encoder_load(sd, **kw)
    return sd


def _check(out, ref, tol=1e-3):
    err = np.abs(out - ref).max()
    cos = np.sum(out * ref, axis=-1)
    print("max abs err", err, "min cos", cos.min())
    assert err <= tol                     # contract: 1e-3 of the unit norm per component
    assert np.all(cos > 1 - 1e-5)
    assert np.allclose(np.linalg.norm(out, axis=-1), 1.0, atol=1e-5)


def test_flash_attention_agrees_on_vit_b32(engine):
    """attn_variant 5 forces the any-T kernel on the 50-token model: same result as the T<=64 kernel."""
    sd = _load(engine, synth.VitB32Shape())
    x = torch.randn(70, 3, 224, 224, generator=torch.Generator().manual_seed(5))
    base = engine.encode_images(x.numpy())
    engine.set_option("attn_variant", 5)
    try:
        alt = engine.encode_images(x.numpy())
    finally:
        engine.set_option("attn_variant", 0)
    assert np.abs(alt - base).max() < 5e-4      # fp16 residual stream: see test_vit_variants_agree
    _check(alt[:4], O.get_img_feats_batch_tensor(sd, x[:4]))


def test_short_sequence_vit(engine):
    """17 and 37 tokens: the T <= 64 kernel's generic path with several all-padding key tiles."""
    for image, patch in ((128, 32), (192, 32)):
        shape = synth.VitB32Shape(image=image, patch=patch, width=256, layers=2, heads=4, mlp=1024, out_dim=256)
        sd = _load(engine, shape, seed=image)
        x = torch.randn(6, 3, image, image, generator=torch.Generator().manual_seed(image)) * 1.2
        _check(engine.encode_images(x.numpy()), O.get_img_feats_batch_tensor(sd, x, heads=shape.heads))


@pytest.mark.parametrize("patch,B", [(14, 3), (14, 37), (16, 5), (28, 4)])
def test_small_long_sequence_vit(engine, patch, B):
    """2-layer towers with 257 / 197 / 65 tokens (ragged last key block, ragged last query tile)."""
    shape = synth.VitB32Shape(image=224, patch=patch, width=256, layers=2, heads=4, mlp=1024, out_dim=256)
    sd = _load(engine, shape, seed=patch)
    x = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(B)) * 1.2
    ref = O.get_img_feats_batch_tensor(sd, x, heads=shape.heads)
    _check(engine.encode_images(x.numpy()), ref)
    # device-resident input takes the same path
    assert np.array_equal(engine.encode_images(x.cuda()).cpu().numpy(), engine.encode_images(x.numpy()))


def test_vit_l14_full_size(engine):
    """The reference's default encoder at full size (24 layers, width 1024, 16 heads, d = 768)."""
    sd = _load(engine, synth.VIT_L14, seed=3)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(9)) * 1.2
    ref = O.get_img_feats_batch_tensor(sd, x, heads=16)
    out = engine.encode_images(x.numpy())
    assert out.shape == (2, 768)
    _check(out, ref)


@pytest.mark.parametrize("image,B", [(56, 5), (112, 3)])
def test_head_dim_80_tower(engine, image, B):
    """ViT-H/14 geometry (graph.py:105-111: width 1280, 16 heads of 80, patch 14) on a 2-layer tower: 17 and 65 tokens"""
    shape = synth.VitB32Shape(image=image, patch=14, width=1280, layers=2, heads=16, mlp=1280, out_dim=1024)
    sd = _load(engine, shape, seed=80)
    x = torch.randn(B, 3, image, image, generator=torch.Generator().manual_seed(B)) * 1.2
    _check(engine.encode_images(x.numpy()), O.get_img_feats_batch_tensor(sd, x, heads=shape.heads))


def test_encoder_load_rejects_unsupported_shapes(engine):
    sd = synth.make_vit_weights(synth.VitB32Shape(width=256, layers=1, heads=4, mlp=1024, out_dim=256))
    with pytest.raises(RuntimeError):      # head dim must be 64 or 80
        engine.encoder_load(sd, image=224, patch=32, width=256, layers=1, heads=8, mlp=1024, out_dim=256)
    with pytest.raises(RuntimeError):      # image not a multiple of the patch
        engine.encoder_load(sd, image=230, patch=32, width=256, layers=1, heads=4, mlp=1024, out_dim=256)


@pytest.mark.parametrize("shape,B", [(synth.VitB32Shape(), 70), (synth.VitB32Shape(), 300),
                                     (synth.VitB32Shape(image=224, patch=14, width=256, layers=2, heads=4, mlp=1024, out_dim=256), 9)])
def test_last_layer_class_token_pruning_is_exact(engine, shape, B):
    """The last block computes Q / attention / out-proj / MLP for the class-token row only (the only row
    ln_post + proj consume).  Switching the pruning off must give the same embeddings bit for bit."""
    _load(engine, shape)
    x = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(21))
    pruned = engine.encode_images(x.numpy())
    engine.set_option("last_layer_cls_only", 0)
    try:
        full = engine.encode_images(x.numpy())
    finally:
        engine.set_option("last_layer_cls_only", 1)
    assert np.array_equal(pruned, full)


def test_encode_crops_with_patch14_tower(engine):
    """hmsg_encode_crops with a patch-14 tower (ViT-L/14 geometry: K = 588 padded to 640, 257 tokens) takes the
    crops -> fp32 NCHW -> generic im2col route; it must equal make_crops + encode_images on the same crops."""
    from tests.scenes import scene, load_scene
    H, W, M, n = 240, 320, 5, 3
    sc = scene(n_frames=3, H=H, W=W)
    load_scene(engine, sc)
    engine.voxel_build(); engine.radius_filter(50, 0.5)
    shape = synth.VitB32Shape(image=224, patch=14, width=256, layers=2, heads=4, mlp=1024, out_dim=256)
    sd = _load(engine, shape, seed=14)
    boxes = np.stack([synth.make_mask_boxes(int(sc["ids"][f]) + 5, H, W, M) for f in range(n)])
    engine.masks_boxes(0, boxes)
    B = n * (2 * M + 1)
    a = torch.empty((B, 256), dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
    engine.encode_crops(0, n, M, boxes, 50, a)
    ptr = engine.make_crops(0, n, M, boxes, 50)
    engine.encode_images_ptr(ptr, B, b)
    engine.sync()
    assert torch.equal(a, b)
    crops = engine.crops_read(B)
    ref = O.get_img_feats_batch_tensor(sd, torch.from_numpy(crops[:4]), heads=shape.heads)
    _check(a[:4].cpu().numpy(), ref)
