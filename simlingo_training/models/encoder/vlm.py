"""Vision-language encoder wrapper (drop-in for reference ``simlingo_training/models/encoder/vlm.py``)."""
from torch import nn

from simlingo_training.models.encoder.internvl2_model import LingoInternVLModel


class VLMEncoderModel(nn.Module):
    def __init__(self, cfg_data_module, processor, cache_dir, **cfg):
        super().__init__()
        # hydra-style config objects are flattened onto the module (reference vlm.py:15-18)
        for source in (cfg, cfg_data_module):
            for key, value in source.items():
                setattr(self, key, value)
        self.token_size = self.embed_dim
        if "internvl2" not in self.variant.lower():
            raise ValueError(f"Unknown variant {self.variant}")
        self.image_encoder = LingoInternVLModel(self.variant, *cfg)
        self.image_encoder.processor = processor
        self.image_encoder.use_global_img = self.use_global_img
        # the chat model's own LLM is dropped: SimLingo wraps a second copy with LoRA (reference vlm.py:30-31)
        self.image_encoder.language_model = None
        self.image_encoder.model.language_model = None
        print("\033[91m" + f"Using {self.variant} as the image encoder." + "\033[0m")
        if self.freeze:
            print("\033[91m" + "Image encoder weights frozen." + "\033[0m")
            for p in self.parameters():
                p.requires_grad = False
            for p in self.image_encoder.model.mlp1.parameters():
                p.requires_grad = True
