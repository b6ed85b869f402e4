"""Generates ``tests/golden/preprocess.npz`` by running the REFERENCE's own image pre-processing in this container:
``/root/reference/simlingo_training/utils/internvl2_utils.py::preprocess_image_batch`` (PIL bicubic resize to the
tile grid -> 448x448 tiles -> ToTensor -> ImageNet normalisation), imported unmodified (``hydra`` is stubbed: the
module only imports ``to_absolute_path`` from it).  This is the step right before the hot path
(``team_code/agent_simlingo.py:483-502``).

Inputs are regenerated from seeds by ``oracle.preprocess.synth_camera``.  The outputs are reduced to the uint8 tiles the
normalisation was applied to (exact inverse of ToTensor/Normalize: every value is k/255) and stored as a SHA-256 over
all of them plus every 5th pixel for diagnostics (50 KB per case instead of 1.2 MB).

    python tests/golden/make_golden_preprocess.py
"""
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from oracle.preprocess import CASES, synth_camera  # noqa: E402

hydra = types.ModuleType("hydra")
hydra.utils = types.ModuleType("hydra.utils")
hydra.utils.to_absolute_path = lambda p: p
sys.modules["hydra"], sys.modules["hydra.utils"] = hydra, hydra.utils
spec = importlib.util.spec_from_file_location("ref_internvl2_utils", "/root/reference/simlingo_training/utils/internvl2_utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

out = {}
mean = torch.tensor(ref.IMAGENET_MEAN).view(1, 3, 1, 1)
std = torch.tensor(ref.IMAGENET_STD).view(1, 3, 1, 1)
for name, (h, w, seed) in CASES.items():
    img = torch.from_numpy(synth_camera(h, w, seed))
    res = ref.preprocess_image_batch([img], input_size=448, use_global_img=False, max_num_grid=2)
    pv = res["pixel_values"][0]                      # [tiles, 3, 448, 448] float32
    u8 = torch.round((pv * std + mean) * 255.0)
    assert torch.equal(((u8 / 255.0) - mean) / std, pv), "uint8 inverse is not exact"
    u8n = u8.to(torch.uint8).numpy()
    out[name + "_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(u8n).tobytes()).digest(), dtype=np.uint8)
    out[name + "_sub5"] = u8n[:, :, ::5, ::5].copy()
    out[name + "_norm_row0"] = pv[:, :, 0, :16].numpy()   # a few normalised float32 values (operation order of ToTensor / Normalize)
    out[name + "_image_sizes"] = res["image_sizes"].numpy()
    print(name, tuple(pv.shape), res["image_sizes"].tolist())
    # use_global_img=True (reference dynamic_preprocess(use_thumbnail=True), :262-265): the thumbnail tile is appended
    res_t = ref.preprocess_image_batch([img], input_size=448, use_global_img=True, max_num_grid=2)
    pv_t = res_t["pixel_values"][0]
    u8_t = torch.round((pv_t * std + mean) * 255.0)
    assert torch.equal(((u8_t / 255.0) - mean) / std, pv_t) and pv_t.shape[0] == 3 and torch.equal(pv_t[:2], pv)
    out[name + "_thumb_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(u8_t[2].to(torch.uint8).numpy()).tobytes()).digest(), dtype=np.uint8)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "preprocess.npz"), **out)
