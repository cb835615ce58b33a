"""Drop-in for fsr_vln/perception/models/sam_clip_feats_extractor.py:extract_feats_per_pixel
(:82-191).  SAM is outside the hot path: ``mask_generator`` is any object with
``generate(image) -> list of {"segmentation","bbox",...}``.  Crops, preprocessing, the encoder,
the fusion and the dense per-pixel map all run in libhmsg_b200.so."""
from __future__ import annotations

import numpy as np


def extract_feats_per_pixel(image, mask_generator, clip_model, preprocess, clip_feat_dim=768, bbox_margin=0, maskedd_weight=0.75):
    """Returns (outfeat [H,W,d] fp16 CPU tensor, F_p [M,d] CPU tensor, masks, F_g np.float32 [1,d]).
    With no masks the reference returns a 3-tuple of None (extractor.py:163-164); kept."""
    import torch
    eng = clip_model.engine
    H, W = image.shape[0], image.shape[1]
    masks = mask_generator.generate(image)
    M = len(masks)
    if M == 0:
        return None, None, None
    d = clip_feat_dim
    # a private 1-frame scene: the image rides in the rgb slot, depth=1 everywhere (the mask
    # bitset must not be ANDed with a depth validity here)
    if getattr(eng, "_xf_shape", None) != (H, W):
        eng.scene_begin(H, W, np.eye(3), 1000.0, 0.05, 1)
        eng._xf_shape = (H, W)
    eng.scene_reset_frames()
    eng.add_frames(np.ones((1, H, W), np.uint16), np.ascontiguousarray(image, dtype=np.uint8)[None], np.eye(4).reshape(1, 16))
    eng.voxel_build(); eng.radius_filter(0, 1.0)
    seg = np.stack([np.asarray(m["segmentation"]).astype(np.uint8) for m in masks])
    boxes = np.array([[int(v) for v in m["bbox"]] for m in masks], dtype=np.int32)[None]
    eng.masks_dense(0, seg[None])
    crops_ptr = eng.make_crops(0, 1, M, boxes, int(bbox_margin))
    feats = torch.empty((2 * M + 1, d), dtype=torch.float32, device=f"cuda:{eng.device}")
    eng.encode_images_ptr(crops_ptr, 2 * M + 1, feats)
    eng.features_begin(d)
    Fp = torch.empty((1, M, d), dtype=torch.float32, device=feats.device)
    eng.fuse_scatter(0, 1, M, feats.view(1, 2 * M + 1, d), float(maskedd_weight), Fp_out=Fp)
    outfeat = torch.from_numpy(eng.pixel_feature_map(0)).reshape(H, W, d)
    eng.torch_wait()
    return outfeat, Fp[0].cpu(), masks, feats[2 * M:2 * M + 1].cpu().numpy()
