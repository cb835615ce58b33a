"""CPU: the int8 tensor-core form of PIL's antialiased bicubic 512 -> 224 pass is exact.

libhmsg_b200.so exports the coefficient fragments it feeds to `mma.sync.m16n8k32.s32.u8.{s8,u8}` (host code, no
GPU needed).  This test replays the kernel's arithmetic in numpy - fragment layouts of PTX m16n8k32, three 8-bit
digit planes recombined by shifting the int32 accumulator, rounding constant, clip - and holds the result to the
real PIL `Image.resize(..., BICUBIC)` the reference's open_clip preprocess calls (utils/clip_utils.py:88-89),
bit for bit, in both directions (horizontal pass on rows, vertical pass on the transposed intermediate)."""
import ctypes as C

import numpy as np
import pytest
from PIL import Image

from holoagent_b200 import _lib, build


@pytest.fixture(scope="module")
def table():
    if build.needs_build():
        build.build()
    lib = _lib.load()
    frag = np.zeros((28, 3, 32, 2), np.uint32)
    x0 = np.zeros(28, np.int32)
    assert lib.hmsg_debug_pil_mma_table(frag.ctypes.data_as(C.c_void_p), x0.ctypes.data_as(C.c_void_p)) == 0
    return frag, x0


def _b_matrix(frag_jq, signed):
    """B fragment (32 x 8, col): lane = n*4 + t; reg 0 = k 4t..4t+3, reg 1 = k 16+4t..16+4t+3 (low byte = low k)."""
    B = np.zeros((32, 8), np.int64)
    for lane in range(32):
        n, t = lane >> 2, lane & 3
        for reg, kb in ((0, 4 * t), (1, 16 + 4 * t)):
            for b in range(4):
                v = (int(frag_jq[lane, reg]) >> (8 * b)) & 255
                B[kb + b, n] = v - 256 if (signed and v >= 128) else v
    return B


def _banded_pass(rows_u8, frag, x0):
    """rows_u8 [R, 512] -> [R, 224]: what k_crop_rows_mma / k_crop_cols_mma compute per channel."""
    R = rows_u8.shape[0]
    padded = np.zeros((R, 560), np.int64); padded[:, :512] = rows_u8
    out = np.zeros((R, 224), np.uint8)
    for j in range(28):
        A = padded[:, x0[j]:x0[j] + 32]                                   # the 32-byte window of the tile
        acc = (A @ _b_matrix(frag[j, 0], True)).astype(np.int64)          # d2 plane (signed)
        acc = ((acc << 8) + A @ _b_matrix(frag[j, 1], False))             # d1
        acc = ((acc << 8) + (1 << 21) + A @ _b_matrix(frag[j, 2], False)) # d0 + PIL rounding constant
        acc = ((acc + (1 << 31)) % (1 << 32)) - (1 << 31)                 # int32 wrap-around (never triggers: the sum fits)
        out[:, 8 * j:8 * j + 8] = np.clip(acc >> 22, 0, 255)
    return out


def test_window_starts_are_word_aligned(table):
    frag, x0 = table
    assert np.all(x0 % 4 == 0) and np.all(np.diff(x0) > 0) and x0[-1] + 32 <= 560


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_banded_imma_equals_pil_bicubic(table, seed):
    frag, x0 = table
    rs = np.random.RandomState(seed)
    img = rs.randint(0, 256, (512, 512), dtype=np.uint8)
    if seed == 1:
        img = ((np.add.outer(np.arange(512), np.arange(512)) % 2) * 255).astype(np.uint8)     # worst-case ringing: clip8 on both sides
    if seed == 2:
        img[:, :256] = 255; img[:256, :] = 0
    # PIL does the horizontal pass first, then the vertical pass, with a clip to uint8 in between (Resample.c)
    ref_h = np.asarray(Image.fromarray(img).resize((224, 512), Image.BICUBIC))
    got_h = _banded_pass(img, frag, x0)
    assert np.array_equal(got_h, ref_h)
    ref = np.asarray(Image.fromarray(img).resize((224, 224), Image.BICUBIC))
    got = _banded_pass(got_h.T.copy(), frag, x0).T                                            # vertical pass = same table on columns
    assert np.array_equal(got, ref)
