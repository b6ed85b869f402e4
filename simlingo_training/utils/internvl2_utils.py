"""Image pre-processing entry points of the reference's ``simlingo_training/utils/internvl2_utils.py`` that sit directly
in front of the model, running on the GPU: same names and call signature, ``pixel_values`` come back
as bf16 CUDA tensors ready for ``DrivingInput.camera_images``.  Tokenizer / chat-template helpers of that file are out
of scope (no tokenizer offline)."""
from simlingo_b200.preprocess import preprocess_frames, preprocess_image_batch, tile_grid  # noqa: F401

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
