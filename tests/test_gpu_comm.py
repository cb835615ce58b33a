"""GPU: the C-ABI collective layer (csrc/comm.cu).  World 1 runs on any box (NCCL with a single rank:
the sharded entry points must reduce to the single-GPU ones); the multi-rank equality test spawns torchrun and skips
when the box has fewer than 2 GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from holoagent_b200 import synth
from tests.scenes import scene, load_scene

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world1_sharded_entry_points_equal_single_gpu(engine):
    sc = scene()
    load_scene(engine, sc)
    F = len(sc["ids"])
    nv, mb = engine.voxel_build()
    a = engine.voxels_read()
    nn = engine.radius_filter(150, 0.4)
    na = engine.nodes_read()
    engine.comm_init_local()
    assert engine.comm_info()[:2] == (0, 1)
    nv2, mb2 = engine.voxel_build_sharded_c([(0, 4), (4, F - 4)])
    b = engine.voxels_read()
    assert nv2 == nv and np.array_equal(mb, mb2)
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    assert np.allclose(a[0], b[0], rtol=1e-12, atol=1e-12)
    assert engine.radius_filter_sharded(150, 0.4) == nn
    nb = engine.nodes_read()
    assert np.array_equal(na[2], nb[2]) and np.array_equal(na[3], nb[3])
    # node-embedding merge with one rank: the partial is already the sum
    d, M = 128, 4
    engine.features_begin(d)
    rs = np.random.RandomState(0)
    feats = rs.randn(3, 2 * M + 1, d).astype(np.float32)
    feats /= np.linalg.norm(feats, axis=-1, keepdims=True)
    boxes = np.stack([synth.make_mask_boxes(int(sc["ids"][i]), sc["H"], sc["W"], M) for i in range(3)])
    engine.masks_boxes(0, boxes)
    engine.fuse_scatter(0, 3, M, feats, 0.4418)
    s0, c0 = engine.node_feats_raw()
    engine.allgather_nodes()
    s1, c1 = engine.node_feats_raw()
    assert np.array_equal(s0, s1) and np.array_equal(c0, c1)
    assert engine.comm_info()[2] == 0.0


@pytest.mark.parametrize("collective", ["c", "torch"])
def test_multi_gpu_ingest_equals_single(collective):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "scripts", "check_multigpu.py"), "--collective", collective, "--frames", "48"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-1500:] + "\n" + "\n".join(l for l in r.stderr.splitlines() if "Error" in l or "error" in l or "assert" in l.lower())[-2500:])
    assert "identical_across_ranks=True" in r.stdout
