"""GPU camera-frame pre-processing: drop-in for ``preprocess_image_batch`` of the reference's
``simlingo_training/utils/internvl2_utils.py:179-205`` (and the inline copy in ``team_code/agent_simlingo.py:483-502``).

Host side: the tile grid (``dynamic_preprocess`` / ``find_closest_aspect_ratio``, :215-247) and Pillow's resampling
tables (``precompute_coeffs`` + ``normalize_coeffs_8bpc`` of Pillow ``src/libImaging/Resample.c``, bicubic a = -0.5 with
the support stretched by the down-scaling factor), cached per input size.  Device side: ``slb_preprocess_frames``.
The result equals the reference's float32 ``pixel_values`` rounded to bf16 (the dtype the agent casts to,
agent_simlingo.py:751); there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import math
from functools import lru_cache
from typing import Dict, List, Sequence, Tuple, Union

import numpy as np
import torch

from . import lib

PRECISION_BITS = 32 - 8 - 2


class _ResampleTable(C.Structure):
    _fields_ = [("first", C.c_void_p), ("count", C.c_void_p), ("taps", C.c_void_p), ("ksize", C.c_int32)]


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_table(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(first input index, tap count, fixed-point taps [out, ksize]) of a Pillow BICUBIC resize in_size -> out_size."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    first = np.zeros(out_size, np.int32)
    count = np.zeros(out_size, np.int32)
    taps = np.zeros((out_size, ksize), np.int32)
    inv = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        w = [_bicubic((x + lo - center + 0.5) * inv) for x in range(hi - lo)]
        ww = sum(w)
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            taps[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        first[xx], count[xx] = lo, hi - lo
    if in_size == out_size:  # Pillow skips the pass; an identity table keeps the kernel uniform
        first, count = np.arange(out_size, dtype=np.int32), np.ones(out_size, np.int32)
        taps = np.full((out_size, 1), 1 << PRECISION_BITS, np.int32)
    return first, count, taps


def tile_grid(width: int, height: int, min_num: int = 1, max_num: int = 2, image_size: int = 448) -> Tuple[int, int]:
    """(tiles across, tiles down): the aspect ratio with <= max_num tiles closest to width / height."""
    ratios = sorted({(i, j) for n in range(min_num, max_num + 1) for i in range(1, n + 1) for j in range(1, n + 1)
                     if min_num <= i * j <= max_num}, key=lambda r: r[0] * r[1])
    best_diff, best = float("inf"), (1, 1)
    for r in ratios:
        diff = abs(width / height - r[0] / r[1])
        if diff < best_diff:
            best_diff, best = diff, r
        elif diff == best_diff and width * height > 0.5 * image_size * image_size * r[0] * r[1]:
            best = r
    return best


@lru_cache(maxsize=16)
def _plan(height: int, width: int, max_num: int, device_index: int):
    """max_num = 0: the 1 x 1 "thumbnail" grid (whole frame resized to one 448 x 448 tile)"""
    gw, gh = tile_grid(width, height, 1, max_num, 448) if max_num > 0 else (1, 1)
    dev = torch.device("cuda", device_index)
    tabs = []
    for in_size, out_size in ((width, 448 * gw), (height, 448 * gh)):
        first, count, taps = resample_table(in_size, out_size)
        t = [torch.from_numpy(a).to(dev) for a in (first, count, np.ascontiguousarray(taps))]
        tabs.append((t, _ResampleTable(t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), taps.shape[1])))
    return gw, gh, tabs


def preprocess_frames(frames: torch.Tensor, max_num_grid: int = 2, use_global_img: bool = False) -> torch.Tensor:
    """frames uint8 [B, 3, H, W] (CUDA) -> bf16 [B, tiles, 3, 448, 448].  ``use_global_img``: the reference's thumbnail
    (``dynamic_preprocess(use_thumbnail=True)``, internvl2_utils.py:262-265): when the grid has more than one tile, the whole
    frame resized to 448 x 448 is appended as an extra tile (same resampler, 1 x 1 grid)."""
    if not (frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[1] == 3):
        raise RuntimeError("simlingo_b200.preprocess: frames must be a uint8 CUDA tensor [B, 3, H, W] (no CPU fallback)")
    frames = frames.contiguous()
    B, _, H, W = frames.shape

    def run(max_num):
        gw, gh, tabs = _plan(H, W, max_num, frames.device.index or 0)
        tmp = torch.empty((B, 3, H, 448 * gw), device=frames.device, dtype=torch.uint8)
        out = torch.empty((B, gw * gh, 3, 448, 448), device=frames.device, dtype=torch.bfloat16)
        lib._check(lib.load().slb_preprocess_frames(lib._p(frames), lib._p(tmp), C.byref(tabs[0][1]), C.byref(tabs[1][1]), lib._p(out), B, H, W, gw, gh,
                                                   lib._stream()), "preprocess_frames", 2)
        return out

    out = run(max_num_grid)
    if use_global_img and out.shape[1] != 1:
        out = torch.cat([out, run(0)], dim=1)
    return out


def preprocess_image_batch(images_batch_list: Union[Sequence[torch.Tensor], torch.Tensor], input_size: int = 448, use_global_img: bool = False,
                           max_num_grid: int = 2, device: Union[str, torch.device, None] = None) -> Dict[str, torch.Tensor]:
    """Reference signature (internvl2_utils.py:179-205): list of uint8 [3, H, W] images (all of one size) ->
    {'pixel_values': [B, tiles, 3, 448, 448] bf16 on the GPU, 'image_sizes': [B, 2] (height, width)}."""
    if input_size != 448:
        raise NotImplementedError("InternVL2-1B tiles are 448 x 448")
    imgs: List[torch.Tensor] = list(images_batch_list) if not torch.is_tensor(images_batch_list) else list(images_batch_list.unbind(0))
    dev = torch.device(device) if device is not None else (imgs[0].device if imgs[0].is_cuda else torch.device("cuda", torch.cuda.current_device()))
    batch = torch.stack([i.to(torch.uint8) for i in imgs]).to(dev, non_blocking=True)
    sizes = torch.tensor([[int(i.shape[1]), int(i.shape[2])] for i in imgs])
    return {"pixel_values": preprocess_frames(batch, max_num_grid, use_global_img), "image_sizes": sizes}
