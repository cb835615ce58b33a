"""CPU tests: the C-ABI shared library loads and exports exactly the symbols that
include/hmsg_b200.h declares; host-side packing logic; loud failure without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

from holoagent_b200 import _lib, build, synth
from holoagent_b200.engine import HmsgEngine, HmsgError, pack_vit_blob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if build.needs_build():
        build.build()
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "hmsg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hmsg_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(lib):
    syms = header_symbols()
    assert len(syms) >= 30
    assert sorted(_lib.SIGNATURES) == syms            # ctypes prototypes mirror the header one to one
    for s in syms:
        assert hasattr(lib, s), f"{s} not exported"
    assert lib.hmsg_version() >= 100


def test_library_is_sm100a_native():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only container check")
def test_no_cpu_fallback():
    with pytest.raises(HmsgError) as e:
        HmsgEngine(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_pack_vit_blob_layout():
    sh = synth.VitB32Shape(image=64, patch=32, width=256, layers=2, heads=4, mlp=512, out_dim=256)
    sd = synth.make_vit_weights(sh, seed=3)
    blob = pack_vit_blob(sd, 2)
    W, T, Kc = 256, 5, 3 * 32 * 32
    expect = W * Kc + W + T * W + 2 * W + 2 * (2 * W + 3 * W * W + 3 * W + W * W + W + 2 * W + 512 * W + 512 + W * 512 + W) + 2 * W + W * 256
    assert blob.size == expect and blob.dtype == np.float32
    assert np.array_equal(blob[: W * Kc], sd["conv1.weight"].numpy().reshape(-1))
    assert np.array_equal(blob[-W * 256:], sd["proj"].numpy().reshape(-1))
    # every weight is fp16-representable (graph.py:117 precision='fp16')
    assert np.array_equal(blob, blob.astype(np.float16).astype(np.float32))


def test_synth_is_deterministic_and_well_formed():
    d1, c1, T1, K1 = synth.make_frames_np(np.array([3, 11]), 60, 80)
    d2, c2, T2, K2 = synth.make_frames_np(np.array([3, 11]), 60, 80)
    assert np.array_equal(d1, d2) and np.array_equal(c1, c2) and np.array_equal(T1, T2)
    assert d1.dtype == np.uint16 and c1.dtype == np.uint8 and d1.shape == (2, 60, 80)
    assert 0.02 < (d1 == 0).mean() < 0.5 and d1.max() <= 10000
    R = T1[0, :3, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.isclose(np.linalg.det(R), 1.0)
    assert np.allclose(K1, [[40, 0, 40], [0, 40, 30], [0, 0, 1]])
    m = synth.make_masks(5, d1[0], 4)
    assert all(set(x) >= {"segmentation", "bbox", "predicted_iou"} for x in m)
    assert not (m[0]["segmentation"] & (d1[0] == 0)).any()


def test_sass_carries_the_blackwell_instructions_the_design_claims():
    """cuobjdump -sass of the shipped library: tcgen05 MMAs (UTCHMMA, also the cta_group::2 form), TMA loads / stores /
    reduce-add (UTMALDG / UTMASTG / UTMAREDG), TMEM loads (LDTM), the int8 IMMA of the crop resampler and the packed
    FFMA2 of the GELU epilogue.  (PTX names never appear in SASS: /opt/skills/guides/B200_PROFILING.md.)"""
    import collections
    import subprocess
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    ops = collections.Counter(re.findall(r"\b(UTCHMMA(?:\.2CTA)?|UTMALDG\.2D(?:\.2CTA)?|UTMASTG\.2D|UTMAREDG\.2D\.ADD|LDTM|IMMA\.16832\.U8\.[US]8|FFMA2|HMMA\.16816\.F32)", out))
    for need in ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG.2D", "UTMALDG.2D.2CTA", "UTMASTG.2D", "UTMAREDG.2D.ADD", "LDTM",
                 "IMMA.16832.U8.S8", "IMMA.16832.U8.U8", "FFMA2", "HMMA.16816.F32"):
        assert ops[need] > 0, (need, dict(ops))


def test_header_is_plain_c_and_the_c_example_links(tmp_path):
    """include/hmsg_b200.h must be usable from C (cgo / JNI / ctypes bind C, not C++): examples/hmsg_from_c.c compiles as
    pedantic C99 and links against the built library.  Without a GPU the program has to stop at hmsg_ctx_create with the
    library's own message - there is no CPU path to fall into."""
    import shutil
    import subprocess
    import torch
    from holoagent_b200 import build as _b
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    libdir = os.path.dirname(_b.LIB)
    exe = str(tmp_path / "hmsg_from_c")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "hmsg_from_c.c"),
           "-o", exe, "-L", libdir, "-lhmsg_b200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if torch.cuda.is_available():
        return                      # running it is the GPU suite's business
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr)
