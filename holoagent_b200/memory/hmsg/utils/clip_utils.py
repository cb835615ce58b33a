"""Drop-in for the encoder call sites of fsr_vln/memory/hmsg/utils/clip_utils.py
(get_img_feats :63-80, get_img_feats_batch :83-94, get_imgs_feats_batch :109-140,
get_text_feats_multiple_templates :257-349).  ``clip_model`` is a B200ClipModel (the visual
tower lives in libhmsg_b200.so); ``preprocess`` is whatever PIL->tensor transform the caller
uses (open_clip's), exactly as in the reference."""
from __future__ import annotations

import numpy as np


def visual_tower_shape(v):
    """(fp32 state dict, encoder_load keyword arguments) of an open_clip VisionTransformer (graph.py:98-119 builds ViT-L/14,
    ViT-H-14 or ViT-B-32).  The head count comes from the attention module (ViT-H/14: 16 heads of 80, not width / 64), the
    activation from the block's MLP (open_clip's `*-quickgelu` configs / OpenAI weights use QuickGELU)."""
    sd = {k: t.float() for k, t in v.state_dict().items()}
    width = int(sd["conv1.weight"].shape[0])
    layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.resblocks."))
    image = v.image_size[0] if hasattr(v.image_size, "__len__") else v.image_size
    blocks = getattr(getattr(v, "transformer", None), "resblocks", None)
    blk = blocks[0] if blocks is not None and len(blocks) else None
    heads = getattr(getattr(blk, "attn", None), "num_heads", None) or max(1, width // 64)
    act = getattr(getattr(blk, "mlp", None), "gelu", None)
    quick = type(act).__name__ == "QuickGELU"
    return sd, dict(image=int(image), patch=int(sd["conv1.weight"].shape[-1]), width=width, layers=int(layers), heads=int(heads),
                    mlp=int(sd["transformer.resblocks.0.mlp.c_fc.weight"].shape[0]), out_dim=int(sd["proj"].shape[1]), quick_gelu=quick)


class B200ClipModel:
    """Holds the encoder loaded into an HmsgEngine.  ``encode_text`` is delegated to a user
    supplied callable (the text tower is outside the hot path: queries enter as vectors)."""

    def __init__(self, engine, visual_state_dict, text_encoder=None, **shape):
        self.engine = engine
        engine.encoder_load(visual_state_dict, **shape)
        self.text_encoder = text_encoder

    @classmethod
    def from_open_clip(cls, engine, open_clip_model, text_encoder=None, quick_gelu=None):
        sd, shape = visual_tower_shape(open_clip_model.visual)
        if quick_gelu is not None:
            shape["quick_gelu"] = bool(quick_gelu)
        return cls(engine, sd, text_encoder, **shape)

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
