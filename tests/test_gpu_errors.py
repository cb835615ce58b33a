"""GPU: the C-ABI error convention: every entry point returns a status, hmsg_last_error carries the text,
the Python layer raises; calls out of order or out of range fail loudly and leave the ctx usable."""
import numpy as np
import pytest

from holoagent_b200.engine import HmsgEngine, HmsgError
from tests.scenes import scene, load_scene

pytestmark = pytest.mark.gpu

ARG, STATE, CAPACITY = 1, 3, 4        # include/hmsg_b200.h: HMSG_ERR_*


def _raises(code, fn, *a, **kw):
    with pytest.raises(HmsgError) as e:
        fn(*a, **kw)
    msg = str(e.value)
    assert msg.startswith(f"[{code}] ") and len(msg) > 6, msg
    return msg


def test_calls_out_of_order_fail_with_state_errors():
    eng = HmsgEngine(0)
    try:
        sc = scene(n_frames=2, H=60, W=80)
        z16 = np.zeros((1, 60, 80), np.uint16); z8 = np.zeros((1, 60, 80, 3), np.uint8); T = np.eye(4).reshape(1, 16)
        assert "hmsg_scene_begin" in _raises(STATE, eng.add_frames, z16, z8, T)
        assert "hmsg_scene_begin" in _raises(STATE, eng.put_rgb_host, 0, z8)
        _raises(STATE, eng.encode_images, np.zeros((1, 3, 224, 224), np.float32))          # no encoder loaded
        _raises(STATE, eng.query_topk, np.zeros((1, 8), np.float32), 1)                   # no index
        _raises(ARG, eng.set_option, "no_such_option", 1)
        eng.scene_begin(60, 80, sc["K"], 1000.0, 0.05, 2)
        assert "no frames" in _raises(STATE, eng.voxel_build)
        _raises(STATE, eng.radius_filter, 5, 0.5)                                          # voxel table missing
        eng.add_frames(sc["depth"], sc["rgb"], sc["poses"].reshape(-1, 16))
        _raises(CAPACITY, eng.add_frames, z16, z8, T)                                      # capacity was 2 frames
        _raises(ARG, eng.put_rgb_host, 1, np.zeros((2, 60, 80, 3), np.uint8))              # frames 1..2: frame 2 is not stored
        _raises(STATE, eng.nodes_read)
        _raises(STATE, eng.pixel_to_node, 0)
        _raises(STATE, eng.features_begin, 128)
        # ... and the ctx still works after all of that
        nv, _ = eng.voxel_build()
        eng.radius_filter(0, 0.5)
        xyz, _, _, _ = eng.nodes_read()
        assert nv > 0 and len(xyz) > 0
        # an all-zero depth scene has no point at all: the voxel build says so instead of building an empty table
        eng.scene_begin(60, 80, sc["K"], 1000.0, 0.05, 1)
        eng.add_frames(z16, z8, T)
        assert "no valid depth" in _raises(STATE, eng.voxel_build)
    finally:
        eng.close()


def test_mask_batch_misuse(engine):
    sc = scene(n_frames=3, H=60, W=80)
    load_scene(engine, sc)
    engine.voxel_build(); engine.radius_filter(0, 0.5)
    engine.features_begin(128)
    boxes = np.tile(np.array([[5, 5, 20, 20]], np.int32), (2, 3, 1))
    _raises(ARG, engine.masks_boxes, 2, boxes)                      # frames 2..3 of a 3-frame scene
    engine.masks_boxes(0, boxes)
    _raises(ARG, engine.masks_counts, 0, np.array([1, 4], np.int32))    # count above M
    import torch
    feats = torch.zeros((2, 7, 128), device="cuda")
    _raises(STATE, engine.fuse_scatter, 1, 2, 3, feats, 0.4418)    # batch 1..2 was never set
