"""Shared small synthetic scenes for the parity tests (seeded, CPU-generated)."""
from __future__ import annotations

import functools

import numpy as np

from holoagent_b200 import synth


@functools.lru_cache(maxsize=4)
def scene(n_frames=10, H=240, W=320, step=3):
    ids = np.arange(0, n_frames * step, step)
    d, c, T, K = synth.make_frames_np(ids, H, W)
    return {"ids": ids, "depth": d, "rgb": c, "poses": T, "K": K, "H": H, "W": W, "scale": 1000.0, "vs": 0.05}


def load_scene(engine, sc, cap=None):
    engine.scene_begin(sc["H"], sc["W"], sc["K"], sc["scale"], sc["vs"], cap or len(sc["ids"]))
    engine.add_frames(sc["depth"], sc["rgb"], sc["poses"].reshape(-1, 16))
