"""ORACLE - test infrastructure, not product code.

numpy restatement of the image pre-processing that precedes the hot path:
reference ``simlingo_training/utils/internvl2_utils.py:179-267`` (``preprocess_image_batch`` ->
``dynamic_preprocess`` -> ``build_transform``), i.e.

    PIL ``Image.resize((448*gw, 448*gh))`` (default filter BICUBIC)  ->  crop into gw*gh tiles of 448x448
    -> ``T.Resize((448, 448), BICUBIC)`` (a no-op on a 448x448 tile)  ->  ToTensor (/255)  ->  Normalize(ImageNet).

The arithmetic lives in Pillow (third-party, installed here: Pillow 12.2.0, ``src/libImaging/Resample.c``): separable
two-pass resampling, horizontal then vertical, 8-bit intermediate, fixed-point coefficients with
``PRECISION_BITS = 32 - 8 - 2``; bicubic kernel a = -0.5, support 2.0 stretched by the down-scaling factor (antialias).
PINNED: ``tests/test_preprocess.py`` checks this restatement bit-for-bit against ``tests/golden/preprocess.npz``, which
``tests/golden/make_golden_preprocess.py`` produced by running the reference's own function in this container.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # internvl2_utils.py:17-18
IMAGENET_STD = (0.229, 0.224, 0.225)
PRECISION_BITS = 32 - 8 - 2

# name -> (height, width, seed): the agent's cropped front camera (512 - 512*4.8//16 = 359 rows of a 1024x512 frame,
# agent_simlingo.py:470), the un-cropped frame, and a small odd-sized image that is up-scaled in both directions
CASES: Dict[str, Tuple[int, int, int]] = {"agent_359x1024": (359, 1024, 11), "full_512x1024": (512, 1024, 12), "small_101x203": (101, 203, 13)}


from simlingo_b200.spec import synth_camera  # noqa: E402,F401  (deterministic uint8 test image)


def find_closest_aspect_ratio(aspect_ratio, target_ratios, width, height, image_size):
    """internvl2_utils.py:215-229"""
    best_diff, best = float("inf"), (1, 1)
    area = width * height
    for ratio in target_ratios:
        diff = abs(aspect_ratio - ratio[0] / ratio[1])
        if diff < best_diff:
            best_diff, best = diff, ratio
        elif diff == best_diff and area > 0.5 * image_size * image_size * ratio[0] * ratio[1]:
            best = ratio
    return best


def tile_grid(width: int, height: int, min_num: int = 1, max_num: int = 2, image_size: int = 448) -> Tuple[int, int]:
    """internvl2_utils.py:231-247: (tiles across, tiles down) of the closest aspect ratio with <= max_num tiles."""
    ratios = sorted({(i, j) for n in range(min_num, max_num + 1) for i in range(1, n + 1) for j in range(1, n + 1)
                     if min_num <= i * j <= max_num}, key=lambda r: r[0] * r[1])
    return find_closest_aspect_ratio(width / height, ratios, width, height, image_size)


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Pillow ``precompute_coeffs`` + ``normalize_coeffs_8bpc``: per output index the first input index, the number of
    taps and the fixed-point taps (int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        n = hi - lo
        w = [_bicubic((x + lo - center + 0.5) * ss) for x in range(n)]
        ww = sum(w)                                    # same left-to-right double accumulation as the C loop
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        xmin[xx], cnt[xx] = lo, n
    return xmin, cnt, kk


def _resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """one 8-bit pass of ``ImagingResampleHorizontal_8bpc`` / ``Vertical``: ss = 2^(P-1) + sum pix*k ; clip8(ss >> P)"""
    in_size = img.shape[axis]
    if in_size == out_size:
        return img
    xmin, cnt, kk = resample_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, -1).astype(np.int64)
    out = np.empty(src.shape[:-1] + (out_size,), np.int64)
    for xx in range(out_size):
        n = cnt[xx]
        acc = (src[..., xmin[xx]:xmin[xx] + n] * kk[xx, :n].astype(np.int64)).sum(-1) + (1 << (PRECISION_BITS - 1))
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out.astype(np.uint8), -1, axis)


def pil_resize_bicubic(img_chw: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """``PIL.Image.resize((out_w, out_h))`` on an RGB uint8 image: horizontal pass first, then vertical."""
    return _resample_axis(_resample_axis(img_chw, out_w, 2), out_h, 1)


def preprocess_tiles_u8(img_chw: np.ndarray, max_num_grid: int = 2, image_size: int = 448, use_thumbnail: bool = False) -> np.ndarray:
    """-> uint8 [tiles, 3, 448, 448]: the resized image cut into its grid, row-major (internvl2_utils.py:249-262); with
    ``use_thumbnail`` and more than one tile, the whole image resized to 448 x 448 is appended (:263-265)."""
    _, h, w = img_chw.shape
    gw, gh = tile_grid(w, h, 1, max_num_grid, image_size)
    big = pil_resize_bicubic(img_chw, gh * image_size, gw * image_size)
    tiles: List[np.ndarray] = []
    for i in range(gw * gh):
        x0, y0 = (i % gw) * image_size, (i // gw) * image_size
        tiles.append(big[:, y0:y0 + image_size, x0:x0 + image_size])
    if use_thumbnail and len(tiles) != 1:
        tiles.append(pil_resize_bicubic(img_chw, image_size, image_size))
    return np.stack(tiles)


def normalize(tiles_u8: np.ndarray) -> np.ndarray:
    """ToTensor + Normalize in float32 with torch's operation order: (u8 / 255 - mean) / std."""
    x = tiles_u8.astype(np.float32) / np.float32(255.0)
    mean = np.asarray(IMAGENET_MEAN, np.float32).reshape(1, 3, 1, 1)
    std = np.asarray(IMAGENET_STD, np.float32).reshape(1, 3, 1, 1)
    return (x - mean) / std
