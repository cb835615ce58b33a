"""CPU tests: the oracle against hand-computed known answers and the committed golden
fixtures (tests/golden/, made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from holoagent_b200 import synth
from oracle import hmsg_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def test_unproject_known_answer():
    # fx=fy=2, cx=cy=1 ; depth 2000 mm at (x=3,y=1) -> X=(3-1)*2/2=2, Y=0, Z=2 ; + translation
    K = np.array([[2.0, 0, 1.0], [0, 2.0, 1.0], [0, 0, 1]])
    depth = np.zeros((3, 4), np.uint16); depth[1, 3] = 2000; depth[2, 0] = 500
    rgb = np.zeros((3, 4, 3), np.uint8); rgb[1, 3] = (255, 0, 51)
    T = np.eye(4); T[:3, 3] = (10, 20, 30)
    p, c, m = O.create_pcd(rgb, depth, K, 1000.0, T)
    assert m.sum() == 2
    # row-major pixel order: (y=1,x=3) then (y=2,x=0)
    assert np.array_equal(p[0], [12.0, 20.0, 32.0])
    assert np.array_equal(p[1], [10 + (0 - 1) * 0.5 / 2, 20 + (2 - 1) * 0.5 / 2, 30.5])
    assert np.array_equal(c[0], [1.0, 0.0, 0.2])
    # depth is a float32 quotient (generic.py:111): 1 mm is not exactly 0.001
    depth[:] = 0; depth[0, 0] = 1
    p, _, _ = O.create_pcd(rgb, depth, K, 1000.0, np.eye(4))
    assert p[0, 2] == float(np.float32(1) / np.float32(1000.0)) and p[0, 2] != 0.001
    # mask_img: "rgb" is a bool mask that multiplies the depth
    depth[:] = 1500
    msk = np.zeros((3, 4), bool); msk[0, 1] = True
    p, c, m = O.create_pcd(msk, depth, K, 1000.0, np.eye(4), mask_img=True)
    assert len(p) == 1 and c is None and m[0, 1]


def test_voxel_down_sample_known_answer():
    # min=0 -> voxel_min_bound=-0.125 ; keys floor((p+0.125)/0.25)
    pts = np.array([[0.0, 0, 0], [0.26, 0, 0], [0.49, 0, 0], [0.12, 0, 0], [0.3, 0.0, 0.0]])
    cols = np.arange(15, dtype=np.float64).reshape(5, 3) / 15
    p, c, ijk, inv = O.voxel_down_sample(pts, cols, 0.25)
    assert ijk.tolist() == [[0, 0, 0], [1, 0, 0], [2, 0, 0]]
    assert np.array_equal(inv, [0, 1, 2, 0, 1])
    assert np.allclose(p[:, 0], [0.06, 0.28, 0.49]) and np.allclose(c[0], (cols[0] + cols[3]) / 2)
    # canonical order is ascending (i,j,k) regardless of input order
    p2, _, ijk2, _ = O.voxel_down_sample(pts[::-1].copy(), cols[::-1].copy(), 0.25)
    assert np.array_equal(ijk2, ijk) and np.allclose(p2, p)


def test_radius_outlier_known_answer():
    # 3x3x1 lattice with spacing 1: centre has 4 neighbours at d=1 ... strict d2 < r2
    g = np.array([[x, y, 0.0] for x in range(3) for y in range(3)], dtype=np.float64)
    cnt = O.radius_counts(g, radius=1.0)          # d2 < 1 -> only self
    assert cnt.tolist() == [1] * 9
    cnt = O.radius_counts(g, radius=1.0000001)
    assert cnt.reshape(3, 3).tolist() == [[3, 4, 3], [4, 5, 4], [3, 4, 3]]
    keep = O.radius_outlier_keep(g, nb_points=3, radius=1.0000001)   # keep iff count > nb_points
    assert keep.tolist() == [1, 3, 4, 5, 7]


def test_last_writer_wins_rule_matches_torch():
    s = torch.zeros(4, 2); c = torch.zeros(4, 1)
    f = torch.arange(8.0).reshape(4, 2)
    idx = np.array([1, 1, 2, 1])
    O.scatter_node_feats(s, c, idx, f)
    assert s.tolist() == [[0, 0], [6, 7], [4, 5], [0, 0]] and c.reshape(-1).tolist() == [0, 1, 1, 0]
    nodes, pos = O.winners(idx, 4)
    assert nodes.tolist() == [1, 2] and pos.tolist() == [3, 2]


def test_dense_map_equals_sparse_form():
    rs = np.random.RandomState(0)
    H, W, M, d = 24, 32, 5, 128
    segs = rs.rand(M, H, W) > 0.6
    Fp = rs.randn(M, d).astype(np.float32); Fp /= np.linalg.norm(Fp, axis=1, keepdims=True)
    dense = O.pixel_feature_map(Fp, segs, H, W)
    pix = np.arange(H * W)[::7]
    assert torch.equal(dense[pix], O.pixel_features_at(Fp, segs.reshape(M, -1)[:, pix]))
    assert dense.dtype == torch.float16
    none = ~segs.reshape(M, -1).any(0)
    assert not dense[torch.from_numpy(none)].any()         # pixels in no mask stay zero


def test_retrieval_known_answers():
    E = np.eye(4, dtype=np.float32)[[0, 1, 2, 3, 0]] * np.array([1, 2, 3, 4, 1], np.float32)[:, None]
    ids, sc = O.query_topk(np.array([1, 1, 1, 1], np.float32), E, 3)
    assert ids.tolist() == [3, 2, 1] and sc.tolist() == [4, 3, 2]
    ids, _ = O.query_topk(np.array([1, 0, 0, 0], np.float32), E, 2)     # tie -> lower index first
    assert ids.tolist() == [0, 4]
    # negative prompts: objects whose column argmax is the query, by descending max score
    q = np.array([[1, 0, 0, 0], [0, 1, 1, 1]], np.float32)
    top, s = O.query_object_core(q, E, 0, 5, True)
    assert top.tolist() == [0, 4] and s.tolist() == [1, 1]
    top, s = O.query_object_core(q, E, 1, 2, True)
    assert top.tolist() == [3, 2]
    assert O.identify_object(E[2], E) == 2
    tm = O.template_mean(np.array([[1, 0], [0, 1], [2, 2], [4, 4]], np.float32))
    assert tm.tolist() == [[0.5, 0.5], [3, 3]]


def test_crop_geometry_known_answers():
    assert O.increase_bbox_by_margin((10, 5, 20, 20), 50) == (0, 0, 80, 75)   # negative x/y clamped, w/h shrink
    img = (np.arange(60 * 80 * 3) % 251).astype(np.uint8).reshape(60, 80, 3)
    seg = np.zeros((60, 80), bool); seg[10:30, 20:50] = True
    masks = [{"bbox": [20, 10, 30, 20], "segmentation": seg}]
    plain = O.crop_all_bounding_boxs(img, masks, False, 50)[0]
    blocked = O.crop_all_bounding_boxs(img, masks, True, 50)[0]
    assert plain.shape == blocked.shape == (512, 512, 3)
    t = O.clip_preprocess(plain)
    assert t.shape == (3, 224, 224) and t.dtype == torch.float32
    assert O.clip_preprocess(img).shape == (3, 224, 224)     # non-square: resize shorter side + centre crop


def test_golden_tiny_scene():
    g = np.load(os.path.join(G, "tiny_scene.npz"))
    vs = float(g["voxel_size"])
    geo = O.build_geometry(g["depth"], g["rgb"], g["poses"], g["K"], 1000.0, vs, int(g["nb_points"]), float(g["radius"]))
    assert np.array_equal(geo["voxel_ijk"], g["voxel_ijk"]) and np.array_equal(geo["keep"], g["keep"])
    assert np.array_equal(geo["voxel_xyz"], g["voxel_xyz"]) and np.array_equal(geo["min_bound"], g["min_bound"])
    tree = O.build_tree(geo["node_xyz"])
    f = 2
    p, _, m = O.create_pcd(g["rgb"][f], g["depth"][f], g["K"], 1000.0, g["poses"][f])
    assert np.array_equal(O.pixel_to_node(tree, p)[1], g["pix_idx"][f][m.reshape(-1)])
    M = g["boxes"].shape[1]
    Fp = O.fuse_mask_feats(g["feats"][f, :M], g["feats"][f, M:2 * M], g["feats"][f, 2 * M:], float(g["maskedd_weight"]))
    assert np.allclose(Fp, g["F_p"][f], atol=1e-7)
    for i in range(3):
        ids, sc = O.query_topk(g["query"][i], g["full_feats"], 4)
        assert np.array_equal(ids, g["top_ids"][i]) and np.allclose(sc, g["top_scores"][i], atol=1e-7)


def test_golden_small_vit():
    g = np.load(os.path.join(G, "small_vit.npz"))
    sh = synth.VitB32Shape(image=64, patch=32, width=256, layers=2, heads=4, mlp=512, out_dim=256)
    sd = synth.make_vit_weights(sh, seed=3)
    out = O.get_img_feats_batch_tensor(sd, torch.from_numpy(g["x"]), heads=4)
    assert np.allclose(out, g["out"], atol=1e-6)
    assert np.allclose(np.linalg.norm(out, axis=1), 1, atol=1e-6)


def test_vit_oracle_matches_transformers_clip():
    """Independent restatement check: same weights through HF CLIPVisionModelWithProjection."""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    sh = synth.VitB32Shape(image=64, patch=32, width=128, layers=2, heads=2, mlp=256, out_dim=64)
    sd = synth.make_vit_weights(sh, seed=1)
    cfg = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, image_size=64, patch_size=32,
                           projection_dim=64, hidden_act="gelu", attn_implementation="eager")
    m = CLIPVisionModelWithProjection(cfg).eval()
    hf = m.state_dict()
    hf["vision_model.embeddings.class_embedding"].copy_(sd["class_embedding"])
    hf["vision_model.embeddings.patch_embedding.weight"].copy_(sd["conv1.weight"])
    hf["vision_model.embeddings.position_embedding.weight"].copy_(sd["positional_embedding"])
    hf["vision_model.pre_layrnorm.weight"].copy_(sd["ln_pre.weight"]); hf["vision_model.pre_layrnorm.bias"].copy_(sd["ln_pre.bias"])
    hf["vision_model.post_layernorm.weight"].copy_(sd["ln_post.weight"]); hf["vision_model.post_layernorm.bias"].copy_(sd["ln_post.bias"])
    hf["visual_projection.weight"].copy_(sd["proj"].T)
    for i in range(2):
        p = f"transformer.resblocks.{i}."; h = f"vision_model.encoder.layers.{i}."
        wq, wk, wv = sd[p + "attn.in_proj_weight"].chunk(3); bq, bk, bv = sd[p + "attn.in_proj_bias"].chunk(3)
        for n, w_, b_ in (("q_proj", wq, bq), ("k_proj", wk, bk), ("v_proj", wv, bv)):
            hf[h + f"self_attn.{n}.weight"].copy_(w_); hf[h + f"self_attn.{n}.bias"].copy_(b_)
        hf[h + "self_attn.out_proj.weight"].copy_(sd[p + "attn.out_proj.weight"]); hf[h + "self_attn.out_proj.bias"].copy_(sd[p + "attn.out_proj.bias"])
        hf[h + "layer_norm1.weight"].copy_(sd[p + "ln_1.weight"]); hf[h + "layer_norm1.bias"].copy_(sd[p + "ln_1.bias"])
        hf[h + "layer_norm2.weight"].copy_(sd[p + "ln_2.weight"]); hf[h + "layer_norm2.bias"].copy_(sd[p + "ln_2.bias"])
        hf[h + "mlp.fc1.weight"].copy_(sd[p + "mlp.c_fc.weight"]); hf[h + "mlp.fc1.bias"].copy_(sd[p + "mlp.c_fc.bias"])
        hf[h + "mlp.fc2.weight"].copy_(sd[p + "mlp.c_proj.weight"]); hf[h + "mlp.fc2.bias"].copy_(sd[p + "mlp.c_proj.bias"])
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        ref = m(pixel_values=x).image_embeds
    out = O.vit_forward(sd, x, heads=2)
    assert torch.allclose(out, ref, atol=2e-5), (out - ref).abs().max()


def test_device_resampler_arithmetic_matches_cv2_and_pil():
    """crops.cu restates cv2 INTER_LINEAR (8U) and PIL antialiased bicubic in fixed point; the
    same arithmetic in numpy (tests/resample_ref.py) must equal the real libraries bit for bit."""
    import cv2
    from PIL import Image
    from tests import resample_ref as R
    rs = np.random.RandomState(0)
    for (h, w) in [(3, 5), (57, 200), (150, 356), (356, 356), (480, 640), (700, 33), (1024, 1024)]:
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        assert np.array_equal(cv2.resize(img, (512, 512)), R.cv_resize_linear(img, 512, 512)), (h, w)
    for (h, w, nw, nh) in [(512, 512, 224, 224), (480, 640, 298, 224), (240, 320, 298, 224), (120, 160, 298, 224)]:
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((nw, nh), Image.BICUBIC))
        assert np.array_equal(ref, R.pil_resize_bicubic(img, nw, nh)), (h, w)


def test_config0_plumbing_oracle_end_to_end():
    """BASELINE configs[0] (reference CPU path, no GPU), scaled to keep the CPU suite fast:
    8 frames 320x240, M=3 masks, a 2-layer ViT.  The whole build-and-retrieve path runs on the
    oracle alone and is deterministic."""
    H, W, F, M = 240, 320, 8, 3
    depth, rgb, T, K = synth.make_frames_np(np.arange(F) * 5, H, W)
    geo = O.build_geometry(depth, rgb, T, K, 1000.0, 0.05, nb_points=300, radius=0.6)
    n = len(geo["node_xyz"])
    assert 1000 < n <= len(geo["voxel_xyz"])
    tree = O.build_tree(geo["node_xyz"])
    sh = synth.VitB32Shape(image=224, patch=32, width=128, layers=2, heads=2, mlp=256, out_dim=128)
    sd = synth.make_vit_weights(sh, seed=2)
    sum_f = torch.zeros(n, 128); cnt = torch.zeros(n, 1)
    for f in range(2):
        masks = synth.make_masks(f, depth[f], M)
        crops = O.crop_all_bounding_boxs(rgb[f], masks, True, 50) + O.crop_all_bounding_boxs(rgb[f], masks, False, 50) + [rgb[f]]
        fe = O.get_img_feats_batch_tensor(sd, torch.stack([O.clip_preprocess(c) for c in crops]), heads=2)
        Fp = O.fuse_mask_feats(fe[:M], fe[M:2 * M], fe[2 * M:], 0.4418)
        O.ingest_frame(sum_f, cnt, tree, n, depth[f], rgb[f], T[f], K, 1000.0, Fp, np.stack([m["segmentation"] for m in masks]))
        m3 = O.create_3d_masks([m["segmentation"] for m in masks], depth[f], K, 1000.0, T[f], geo["node_xyz"], geo["node_rgb"], tree, 0.05)
        assert len(m3) == M and all(len(p) <= int(m["segmentation"].sum()) for (p, _, _), m in zip(m3, masks))
    full = O.finalize_node_feats(sum_f, cnt)
    hit = cnt.numpy().reshape(-1) > 0
    assert 0 < hit.sum() < n and np.isfinite(full).all() and not full[~hit].any()
    ids, sc = O.query_topk(full[hit][0] * 0.8, full, 5)
    assert sc[0] >= sc[-1] and len(set(ids.tolist())) == 5


def test_text_template_shim_matches_reference_formula():
    """A10 (clip_utils.py:257-349): two templates per label, rows L2-normalised per text, mean over templates WITHOUT
    re-normalising - the host shim with an injected text tower vs the oracle's template_mean."""
    import types
    from holoagent_b200.memory.hmsg.utils.clip_utils import get_text_feats_multiple_templates
    rs = np.random.RandomState(0)
    d = 64
    bank = {}

    def text_encoder(texts):
        for t in texts:
            bank.setdefault(t, rs.randn(d).astype(np.float32) * 3.0)
        return np.stack([bank[t] for t in texts])

    labels = ["chair", "coffee mug", "plant"]
    out = get_text_feats_multiple_templates(labels, types.SimpleNamespace(text_encoder=text_encoder), d)
    asked = [t.format(lm) for lm in labels for t in ("{}", "a photo of {} in the scene.")]
    assert list(bank) == asked                                    # template order of clip_utils.py:272-276, label-major
    raw = np.stack([bank[t] for t in asked])
    ref = O.template_mean(raw / np.linalg.norm(raw, axis=-1, keepdims=True), 2)
    assert out.shape == (3, d) and np.allclose(out, ref, atol=1e-7)
    assert np.all(np.linalg.norm(out, axis=-1) < 1.0)             # a mean of two unit vectors is not re-normalised (H8)


def test_clip_preprocess_equals_the_torchvision_pipeline_open_clip_builds():
    """open_clip's eval `image_transform` (un-vendored; what the reference's `preprocess` is, clip_utils.py:72-73, :88-89) is a
    torchvision Compose: Resize(224, BICUBIC) -> CenterCrop(224) -> RGB -> ToTensor -> Normalize(CLIP mean / std).  torchvision
    is installed here, so the oracle's restatement is held to the real thing bit for bit - at the 512x512 size every mask crop
    has (sam_utils.py:144), at full-frame shapes (wide, tall) and at odd sizes that exercise the rounding of the resize / crop."""
    tv = pytest.importorskip("torchvision.transforms")
    from PIL import Image
    pipe = tv.Compose([tv.Resize(224, interpolation=tv.InterpolationMode.BICUBIC), tv.CenterCrop(224), lambda im: im.convert("RGB"),
                       tv.ToTensor(), tv.Normalize(mean=O.CLIP_MEAN, std=O.CLIP_STD)])
    rs = np.random.RandomState(8)
    for (h, w) in [(512, 512), (480, 640), (640, 480), (720, 1280), (224, 224), (301, 517), (517, 301), (225, 999)]:
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        ref = pipe(Image.fromarray(img))
        got = O.clip_preprocess(img)
        assert got.shape == ref.shape == (3, 224, 224)
        assert torch.equal(got, ref), (h, w, float((got - ref).abs().max()))
