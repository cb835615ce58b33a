"""Drop-in for the encoder call sites of fsr_vln/memory/hmsg/utils/clip_utils.py
(get_img_feats :63-80, get_img_feats_batch :83-94, get_imgs_feats_batch :109-140,
get_text_feats_multiple_templates :257-349).  ``clip_model`` is a B200ClipModel (the visual
tower lives in libhmsg_b200.so); ``preprocess`` is whatever PIL->tensor transform the caller
uses (open_clip's), exactly as in the reference."""
from __future__ import annotations

import numpy as np


class B200ClipModel:
    """Holds the encoder loaded into an HmsgEngine.  ``encode_text`` is delegated to a user
    supplied callable (the text tower is outside the hot path: queries enter as vectors)."""

    def __init__(self, engine, visual_state_dict, text_encoder=None, **shape):
        self.engine = engine
        engine.encoder_load(visual_state_dict, **shape)
        self.text_encoder = text_encoder

    @classmethod
    def from_open_clip(cls, engine, open_clip_model, text_encoder=None, quick_gelu=False):
        v = open_clip_model.visual
        sd = {k: t.float() for k, t in v.state_dict().items()}
        width = sd["conv1.weight"].shape[0]
        layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.resblocks."))
        return cls(engine, sd, text_encoder, image=v.image_size[0] if hasattr(v.image_size, "__len__") else v.image_size,
                   patch=sd["conv1.weight"].shape[-1], width=width, layers=layers, heads=width // 64,
                   mlp=sd["transformer.resblocks.0.mlp.c_fc.weight"].shape[0], out_dim=sd["proj"].shape[1], quick_gelu=quick_gelu)

    def encode_image_normalized(self, x):
        """x: torch tensor [B,3,S,S] (CPU or CUDA) -> np.float32 [B,d] unit rows."""
        if getattr(x, "is_cuda", False):
            return self.engine.encode_images(x.float().contiguous()).cpu().numpy()
        return self.engine.encode_images(x.float().contiguous().numpy())


def get_img_feats(img, preprocess, clip_model):
    """clip_utils.py:63-80 -> np.float32 [1,d]"""
    from PIL import Image
    img_in = preprocess(Image.fromarray(np.uint8(img)))[None, ...]
    return np.float32(clip_model.encode_image_normalized(img_in))


def get_img_feats_batch(imgs, preprocess, clip_model):
    """clip_utils.py:83-94 -> np.float32 [B,d]"""
    import torch
    from PIL import Image
    imgs_in = torch.stack([preprocess(Image.fromarray(np.uint8(i))) for i in imgs])
    return np.float32(clip_model.encode_image_normalized(imgs_in))


def get_imgs_feats_batch(raw_imgs, preprocess, clip_model, clip_feat_dim, batch_size=64):
    """clip_utils.py:109-140 (float64 output array like the reference's np.zeros)."""
    import torch
    from PIL import Image
    out = np.zeros((len(raw_imgs), clip_feat_dim))
    batch = []
    for i, img in enumerate(raw_imgs):
        if img.shape[0] == 0 or img.shape[1] == 0:
            img = [[[0, 0, 0]]]
        batch.append(preprocess(Image.fromarray(np.uint8(img))))
    if batch:
        out[:] = clip_model.encode_image_normalized(torch.stack(batch))
    return out


def get_text_feats(in_text, clip_model, clip_feat_dim, batch_size=64):
    """clip_utils.py:143-162: delegated to the caller's text tower; rows L2-normalised."""
    if clip_model.text_encoder is None:
        raise RuntimeError("B200ClipModel has no text_encoder: pass query vectors directly (text tower is out of scope)")
    f = np.asarray(clip_model.text_encoder(list(in_text)), dtype=np.float32).reshape(len(in_text), clip_feat_dim)
    return f / np.linalg.norm(f, axis=-1, keepdims=True)


def get_text_feats_multiple_templates(in_text, clip_model, clip_feat_dim, batch_size=64):
    """clip_utils.py:257-349: two templates, mean over templates WITHOUT re-normalising."""
    templates = ["{}", "a photo of {} in the scene."]
    texts = [t.format(lm) for lm in in_text for t in templates]
    f = get_text_feats(texts, clip_model, clip_feat_dim)
    return np.mean(f.reshape((-1, len(templates), f.shape[-1])), axis=1)
