"""numpy restatement of the fixed-point resamplers exactly as crops.cu computes them
(cv2 INTER_LINEAR 8U and PIL antialiased bicubic).  Used by CPU tests to prove the device
arithmetic is bit-exact against the real cv2 / PIL the reference calls."""
import math

import numpy as np


def cv_coefs(src, dst, clamp_coef):
    scale = 1.0 / (dst / src)
    sx = np.zeros(dst, np.int32); a = np.zeros((dst, 2), np.int32)
    for dx in range(dst):
        fx = np.float32((dx + 0.5) * scale - 0.5)
        s = int(math.floor(fx)); fx = np.float32(fx - np.float32(s))
        if clamp_coef:
            if s < 0:
                fx = np.float32(0); s = 0
            if s >= src - 1:
                fx = np.float32(0); s = src - 1
        sx[dx] = s
        a[dx, 0] = int(np.rint(np.float32((np.float32(1.0) - fx) * np.float32(2048))))
        a[dx, 1] = int(np.rint(np.float32(fx * np.float32(2048))))
    return sx, a


def cv_resize_linear(img, dw, dh):
    sh, sw = img.shape[:2]
    sx, ax = cv_coefs(sw, dw, True)
    sy, ay = cv_coefs(sh, dh, False)
    I = img.astype(np.int32)
    x1 = np.minimum(sx + 1, sw - 1)
    rows = I[:, sx, :] * ax[:, 0][None, :, None] + I[:, x1, :] * ax[:, 1][None, :, None]
    y0 = np.clip(sy, 0, sh - 1); y1 = np.clip(sy + 1, 0, sh - 1)
    b0 = ay[:, 0][:, None, None]; b1 = ay[:, 1][:, None, None]
    return ((((b0 * (rows[y0] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2).astype(np.uint8)


def _bicubic(x):
    a = -0.5
    if x < 0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_coeffs(in_size, out_size):
    scale = in_size / out_size; fs = max(scale, 1.0); support = 2.0 * fs
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = []; kk = np.zeros((out_size, ksize), np.int64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale; ss = 1.0 / fs
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << 22)) if v < 0 else int(0.5 + v * (1 << 22))
        bounds.append((xmin, xmax))
    return bounds, kk


def _resample_axis(img, out_size, axis):
    img = np.moveaxis(img, axis, 0).astype(np.int64)
    b, kk = pil_coeffs(img.shape[0], out_size)
    out = np.zeros((out_size,) + img.shape[1:], np.int64)
    for xx in range(out_size):
        xmin, n = b[xx]
        acc = np.full(img.shape[1:], 1 << 21, np.int64)
        for x in range(n):
            acc += img[x + xmin] * kk[xx, x]
        out[xx] = np.clip(acc >> 22, 0, 255)
    return np.moveaxis(out.astype(np.uint8), 0, axis)


def pil_resize_bicubic(img, nw, nh):
    return _resample_axis(_resample_axis(img, nw, 1), nh, 0)
