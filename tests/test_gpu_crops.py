"""GPU parity: A8 (N3) device-side crops + preprocessing vs real cv2 + PIL (bit-exact)."""
import numpy as np
import pytest

from holoagent_b200 import synth
from oracle import hmsg_oracle as O
from tests.scenes import scene, load_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[1, 0], ids=["imma", "scalar"])
def crops_mma(engine, request):
    """both resampling back ends: int8 tensor-core banded PIL passes (default) and the scalar kernels"""
    engine.set_option("crops_mma", request.param)
    yield request.param
    engine.set_option("crops_mma", 1)


@pytest.mark.parametrize("H,W", [(240, 320), (480, 640), (720, 1280)])
def test_crops_bit_exact(engine, crops_mma, H, W):
    sc = scene(n_frames=3, H=H, W=W)
    load_scene(engine, sc)
    engine.voxel_build()
    engine.radius_filter(50, 0.5)
    M = 7
    n = 2
    boxes = np.stack([synth.make_mask_boxes(int(sc["ids"][f]) + 77, H, W, M) for f in range(n)])
    boxes[0, 0] = (0, 0, 40, 30)                      # touches the top-left corner (margin clamp)
    boxes[0, 1] = (W - 45, H - 37, 45, 37)            # touches the bottom-right corner (slicing truncation)
    boxes[1, 2] = (5, H // 2, W - 10, 20)             # wide
    engine.masks_boxes(0, boxes)
    engine.make_crops(0, n, M, boxes, 50)
    got = engine.crops_read(n * (2 * M + 1)).reshape(n, 2 * M + 1, 3, 224, 224)
    for f in range(n):
        valid = sc["depth"][f] > 0
        masks = []
        for (x, y, w, h) in boxes[f]:
            seg = np.zeros((H, W), bool); seg[y:y + h, x:x + w] = valid[y:y + h, x:x + w]
            masks.append({"bbox": [int(x), int(y), int(w), int(h)], "segmentation": seg})
        img = sc["rgb"][f]
        masked = O.crop_all_bounding_boxs(img, masks, True, 50)
        plain = O.crop_all_bounding_boxs(img, masks, False, 50)
        ref = [O.clip_preprocess(c).numpy() for c in masked] + [O.clip_preprocess(c).numpy() for c in plain] + [O.clip_preprocess(img).numpy()]
        ref = np.stack(ref)
        assert np.array_equal(got[f], ref), np.abs(got[f] - ref).max()


def test_fused_crops_encoder_equals_unfused(engine, crops_mma):
    """hmsg_encode_crops (crops -> fp16 patch matrix -> encoder) == make_crops + encode_images, bit for bit"""
    import torch
    H, W, M, n = 240, 320, 5, 3
    sc = scene(n_frames=3, H=H, W=W)
    load_scene(engine, sc)
    engine.voxel_build(); engine.radius_filter(50, 0.5)
    engine.encoder_load(synth.make_vit_weights())
    boxes = np.stack([synth.make_mask_boxes(int(sc["ids"][f]) + 5, H, W, M) for f in range(n)])
    engine.masks_boxes(0, boxes)
    B = n * (2 * M + 1)
    a = torch.empty((B, 512), dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
    engine.encode_crops(0, n, M, boxes, 50, a)
    ptr = engine.make_crops(0, n, M, boxes, 50)
    engine.encode_images_ptr(ptr, B, b)
    engine.sync()
    assert torch.equal(a, b)


def test_imma_and_scalar_patch_matrices_agree(engine):
    """fused path (crops -> fp16 patch matrix -> encoder): both resampling back ends give identical embeddings"""
    import torch
    H, W, M, n = 480, 640, 9, 4
    sc = scene(n_frames=4, H=H, W=W)
    load_scene(engine, sc)
    engine.voxel_build(); engine.radius_filter(50, 0.5)
    engine.encoder_load(synth.make_vit_weights())
    boxes = np.stack([synth.make_mask_boxes(int(sc["ids"][f]) + 11, H, W, M) for f in range(n)])
    boxes[0, 0] = (0, 0, W, H)                         # whole frame as a mask (down-sampling in both directions)
    boxes[1, 1] = (W - 3, H - 2, 3, 2)                 # tiny crop at the corner (up-sampling x170)
    engine.masks_boxes(0, boxes)
    B = n * (2 * M + 1)
    a = torch.empty((B, 512), dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
    engine.encode_crops(0, n, M, boxes, 50, a)
    engine.set_option("crops_mma", 0)
    try:
        engine.encode_crops(0, n, M, boxes, 50, b)
    finally:
        engine.set_option("crops_mma", 1)
    engine.sync()
    assert torch.equal(a, b)
