"""Batch assembly of the reference's ``simlingo_training/dataloader/datamodule.py``: ``encode_uint8`` and
``DataModule.dl_collate_fn`` (:310-443) - the step that turns a list of ``DatasetOutput`` samples into the ``DrivingExample``
the hot path consumes.  Dataset construction, weighted sampling and the Lightning ``DataModule`` shell stay out of scope
(no dataset offline); ``Collator`` carries exactly the attributes ``dl_collate_fn`` reads from ``self`` so that the
reference's DataModule can delegate to it (``collate_fn=Collator.from_datamodule(self)``).

One deliberate difference in the wire format: the reference resizes / tiles / normalises every camera frame with Pillow
inside the dataloader workers and ships float32 tiles (4.8 MB per frame) to the GPU.  Here ``camera_images`` stays the raw
uint8 frame ``[B, T, 3, H, W]`` (1.1 MB per frame at 1024 x 359) and ``DrivingModel`` runs the Pillow-exact resize / tiling /
normalisation on the device (``slb_preprocess_frames``) when it sees a uint8 tensor: 4.4x less H2D traffic, no CUDA in forked
workers, bit-identical tiles.  ``image_sizes`` is the reference's ``[B, 2]`` (height, width) tensor."""
from typing import List, Optional

import numpy as np
import torch

from simlingo_training.utils.custom_types import DrivingExample, DrivingInput, DrivingLabel, LanguageLabel
from simlingo_training.utils.internvl2_utils import get_custom_chat_template, get_num_image_tokens_per_patch
from simlingo_training.utils.projection import get_camera_extrinsics, get_camera_intrinsics


def encode_uint8(strings: List[str], common_length: int) -> torch.Tensor:
    """uint8 [len(strings), common_length]: utf-8 bytes of each string, NUL padded (``decode_uint8`` in driving.py inverts it)."""
    longest = max(len(s) for s in strings)
    assert longest <= common_length, f"String is too long: {longest} > {common_length}"
    rows = np.zeros((len(strings), common_length), dtype=np.uint8)
    for i, s in enumerate(strings):
        raw = s.ljust(common_length, "\0").encode("utf-8")
        if len(raw) != common_length:   # the reference's torch.tensor(list of bytearrays) needs equal lengths as well
            raise ValueError("encode_uint8: multi-byte characters make the padded rows unequal in length")
        rows[i] = np.frombuffer(raw, dtype=np.uint8)
    return torch.from_numpy(rows)


class Collator:
    """``dl_collate_fn`` with the state it reads from the reference's DataModule: tokenizer, encoder variant, number of
    image tiles, ``use_global_img``, ``predict`` and the waypoint flavour of the base dataset."""
    IMAGES_TO_CONSIDER = ["image_ff"]   # front forward camera only (reference :130)
    NUM_IMAGE_PATCHES = 2               # the front camera is split into a 1 x 2 tile grid (:131)

    def __init__(self, tokenizer, encoder_variant: str = "OpenGVLab/InternVL2-1B", use_global_img: bool = False, predict: bool = False,
                 use_1d_wps: bool = False, num_image_tokens_total: Optional[int] = None):
        self.tokenizer = tokenizer
        self.encoder_variant = encoder_variant
        self.use_global_img = use_global_img
        self.predict = predict
        self.use_1d_wps = use_1d_wps
        if "internvl2" not in encoder_variant.lower():
            raise ValueError(f"Image preprocessing for {encoder_variant} not implemented")
        tiles = self.NUM_IMAGE_PATCHES + (1 if use_global_img else 0)
        self.num_image_tokens_total = num_image_tokens_total or get_num_image_tokens_per_patch(encoder_variant) * tiles

    @classmethod
    def from_datamodule(cls, dm) -> "Collator":
        return cls(dm.tokenizer, dm.encoder_variant, getattr(dm, "use_global_img", False), getattr(dm, "predict", False),
                   bool(getattr(getattr(dm, "base_dataset", None), "use_1d_wps", False)), getattr(dm, "num_image_tokens_total", None))

    def __call__(self, data):
        return self.dl_collate_fn(data)

    def dl_collate_fn(self, data) -> DrivingExample:
        n = len(data)
        ref_img = data[0].image_ff
        T, C, H, W = ref_img.shape
        assert T == 1, "Only one timestep as input supported"
        frames = np.stack([np.asarray(d.image_ff) if d.image_ff is not None else np.zeros_like(ref_img) for d in data])
        frames = frames.astype(np.uint8)   # as the reference's preprocess_image_batch (internvl2_utils.py:187): plain astype, no rounding
        camera = torch.from_numpy(np.ascontiguousarray(frames)).view(n, T, C, H, W)
        image_sizes = torch.tensor([[H, W]] * n)

        conversation, question = get_custom_chat_template([d.conversation for d in data], self.tokenizer, self.encoder_variant,
                                                          self.num_image_tokens_total)
        placeholders = [{self.tokenizer.convert_tokens_to_ids(k): v for k, v in d.placeholder_values.items()} for d in data]

        def label(tok):
            return LanguageLabel(phrase_ids=tok["phrase_ids"], phrase_valid=tok["phrase_valid"], phrase_mask=tok["phrase_mask"],
                                 placeholder_values=placeholders, language_string=tok["language_string"], loss_masking=tok["loss_masking"])

        answer = LanguageLabel(phrase_ids=None, phrase_valid=None, phrase_mask=None, placeholder_values=None,
                               language_string=[d.answer[0]["content"][0]["text"] for d in data], loss_masking=None)
        f32 = lambda field: torch.tensor(np.asarray([getattr(d, field) for d in data])).float()
        driving_input = DrivingInput(
            camera_images=camera,   # [B, T, 3, H, W] uint8: tiled and normalised on the GPU by DrivingModel
            image_sizes=image_sizes,
            camera_intrinsics=get_camera_intrinsics(W, H, 110).unsqueeze(0).repeat_interleave(n, dim=0).view(n, 3, 3).float(),
            camera_extrinsics=get_camera_extrinsics().unsqueeze(0).repeat_interleave(n, dim=0).view(n, 4, 4).float(),
            vehicle_speed=f32("speed"),
            target_point=f32("target_points"),
            prompt=label(conversation),
            prompt_inference=label(question),
        )
        driving_label = DrivingLabel(
            waypoints=f32("waypoints_1d" if self.use_1d_wps else "waypoints"),
            path=f32("path"),
            answer=answer,
            image_ff_org=torch.tensor(np.asarray([d.image_ff_org_size for d in data])),
            eval_infos=[d.eval_infos for d in data] if self.predict else None,
        )
        return DrivingExample(
            driving_input=driving_input,
            driving_label=driving_label,
            run_id=encode_uint8([d.measurement_path for d in data], 1000),
            qa_templates=[d.qa_templates[0] if d.qa_templates is not None else None for d in data] if self.predict else None,
        )
