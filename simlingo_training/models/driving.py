"""SimLingo ``DrivingModel`` - drop-in for reference ``simlingo_training/models/driving.py``.

Same constructor (``hydra.utils.instantiate(cfg.model, cfg_data_module=..., processor=..., cache_dir=...)``), same
``forward`` / ``forward_model`` / ``forward_loss`` / ``training_step`` / ``configure_optimizers`` signatures and the same
``state_dict`` keys; the heavy math runs in the sm_100a kernels of ``simlingo_b200``.

Differences that do not change results:
  * generation uses a KV cache and a batched decode instead of re-forwarding the full sequence per token and per batch
    item (reference :133-176, llm.py:217-235); padded rows still get the reference's no-mask final pass (:154-156);
  * ``forward_loss`` evaluates the LM head only on rows that carry a label instead of materialising
    ``[B, L+30, vocab]`` logits (``forward_model`` still returns full logits when called directly);
  * pytorch_lightning / hydra are optional imports (absent in the build image).
"""
import importlib
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from simlingo_b200 import runtime as _rt
from simlingo_b200.spec import LLM_PREFIX
from simlingo_training.models.adaptors.adaptors import AdaptorList, DrivingAdaptor, LanguageAdaptor, WaypointInputAdaptor
from simlingo_training.models.utils import summarise_losses
from simlingo_training.utils.custom_types import DrivingExample, DrivingInput, DrivingLabel, DrivingOutput, TrainingOutput

try:  # pragma: no cover - not installed in the build image
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except ImportError:  # pragma: no cover
    pl = None

    class _Base(nn.Module):
        """nn.Module with the few LightningModule hooks the model code touches."""
        local_rank = 0
        trainer = None

        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass


def _instantiate(cfg, **kwargs):
    """``hydra.utils.instantiate(cfg, **kwargs, _recursive_=False)`` or a local ``_target_`` resolver."""
    try:  # pragma: no cover
        import hydra
        return hydra.utils.instantiate(cfg, **kwargs, _recursive_=False)
    except ImportError:
        items = dict(cfg.items()) if hasattr(cfg, "items") else dict(vars(cfg))
        target = items.pop("_target_")
        module_name, _, cls_name = target.rpartition(".")
        cls = getattr(importlib.import_module(module_name), cls_name)
        return cls(**items, **kwargs)


def decode_uint8(encoded: torch.Tensor) -> List[str]:
    return [bytes(row).decode("utf-8").rstrip("\0") for row in encoded.cpu().numpy()]


class NormZeroOne(nn.Module):
    def __init__(self, min_max: Tuple[float, float]):
        super().__init__()
        self.register_buffer("min_max", torch.tensor(min_max, dtype=torch.float), persistent=False)

    def forward(self, x: Tensor) -> Tensor:
        return (x - self.min_max[0]) / (self.min_max[1] - self.min_max[0])


class DrivingModel(_Base):
    def __init__(self, cfg_data_module, processor, cache_dir, **cfg):
        super().__init__()
        self.save_hyperparameters()
        for key, value in cfg.items():
            setattr(self, key, value)
        self.processor = processor
        self.prediction = {}
        self.predict_language = True
        self.cfg_data_module = cfg_data_module
        self.vision_model = _instantiate(self.vision_model, cfg_data_module=cfg_data_module, processor=self.processor,
                                         cache_dir=cache_dir)
        self.language_model = _instantiate(self.language_model, cache_dir=cache_dir)
        self.all_predictions, self.all_losses = {}, {}
        driving = DrivingAdaptor(self.language_model.hidden_size, speed_wps_mode=self.speed_wps_mode,
                                 predict_route_as_wps=self.predict_route_as_wps)
        self.adaptors = AdaptorList(language=LanguageAdaptor(self.language_model), driving=driving)
        self.wp_encoder = WaypointInputAdaptor(token_size=self.language_model.hidden_size, hidden_size=256, hidden_size2=512)
        self.tokenizer = self.processor.tokenizer if "tokenizer" in self.processor.__dict__ else self.processor
        self.spec = self.language_model.spec

    # ------------------------------------------------------------------------------------------------
    def _engine(self):
        """One engine over the whole model, shared with the sub-modules that can also be called on their own."""
        eng = _rt.engine_for(self, "", self.spec)
        _rt.attach_engine(self.vision_model.image_encoder.model, eng)
        _rt.attach_engine(self.language_model.model.base_model.model, eng)
        return eng

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        _rt.invalidate(self, in_place=True)
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        _rt.invalidate(self)
        return out

    def _eos(self):
        variant = self.language_model.variant
        if variant == "OpenGVLab/InternVL2-4B":
            return self.tokenizer.added_tokens_encoder["<|end|>"]
        if variant == "OpenGVLab/InternVL2-2B":
            return self.tokenizer.added_tokens_encoder["<|im_end|>"]
        return self.tokenizer.eos_token_id

    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, example: DrivingExample, return_language: Optional[bool] = None, prompt_ids: Optional[Tensor] = None):
        """Samples commentary tokens and a trajectory: -> (speed_wps [B,10,2], route [B,20,2], language list[str])."""
        self.speed_wps, self.route, self.language = None, None, []
        driving_input = getattr(example, "driving_input", example)
        eng = self._engine()
        adaptor_dict = self.adaptors(example, inference=True)
        adaptor_dict = self.vision_model.image_encoder.replace_placeholder_tokens(
            adaptor_dict=adaptor_dict,
            pixel_values=self._pixels(driving_input),
            placeholder_values=driving_input.prompt_inference.placeholder_values,
            wp_encoder=self.wp_encoder,
        )
        embeds, masks = adaptor_dict["language_inputs"], adaptor_dict["language_inputs_mask"]
        speed, route, tokens = eng.generate(embeds.to(torch.bfloat16), masks.bool(), max_new_tokens=100, eos_token_id=self._eos())
        self.speed_wps, self.route = speed, route
        self.language = [self.tokenizer.batch_decode(t.unsqueeze(0), skip_special_tokens=True)[0] for t in tokens]
        self.sampled_tokens = tokens
        return self.speed_wps, self.route, self.language

    def _pixels(self, driving_input: DrivingInput) -> Tensor:
        """``camera_images`` as tiles [B, T, NP, 3, 448, 448].  Raw uint8 frames [B, T, 3, H, W] (what
        ``simlingo_training.dataloader.datamodule.Collator`` ships) are resized / tiled / normalised here on the device with the
        Pillow-exact kernel instead of in the dataloader workers (reference internvl2_utils.py:179-267)."""
        cam = driving_input.camera_images
        if cam is not None and cam.dtype == torch.uint8 and cam.dim() == 5:
            from simlingo_b200.preprocess import preprocess_frames
            B, T, C, H, W = cam.shape
            dev = next(self.parameters()).device
            tiles = preprocess_frames(cam.reshape(B * T, C, H, W).to(dev, non_blocking=True), max_num_grid=2,
                                      use_global_img=bool(self.cfg_data_module.get("use_global_img", False)) if hasattr(self.cfg_data_module, "get") else False)
            return tiles.view(B, T, tiles.shape[1], 3, 448, 448)
        return cam

    def _substituted(self, driving_input: DrivingInput, adaptor_dict: Dict) -> Dict:
        return self.vision_model.image_encoder.replace_placeholder_tokens(
            adaptor_dict=adaptor_dict,
            pixel_values=self._pixels(driving_input),
            placeholder_values=driving_input.prompt.placeholder_values,
            wp_encoder=self.wp_encoder,
        )

    def forward_model(self, driving_input: DrivingInput, adaptor_dict: Dict, driving_labels: DrivingLabel = None,
                      want_logits: bool = True):
        """Teacher-forced pass over [valid language | 30 queries | pads] -> (features, logits) for every position."""
        if not torch.is_grad_enabled():
            self._engine()  # inference kernels with LoRA folded into scratch weights; the training path runs un-merged
        adaptor_dict = self._substituted(driving_input, adaptor_dict)
        embeds = adaptor_dict["inputs"].to(dtype=self.language_model.model.dtype)
        feats, logits = _rt.llm_forward(self.language_model.model.base_model.model, embeds, adaptor_dict["inputs_mask"],
                                        want_logits=want_logits)
        return feats, logits

    def param_store(self):
        """Flat bf16 parameter / gradient store of the trainable tensors (created on first use): the hand-written
        backward writes gradients there, ``FusedAdamW`` updates it in one kernel, and
        ``param_store().enable_data_parallel()`` turns on the bucketed NCCL gradient all-reduce."""
        from simlingo_b200 import training
        return training.ensure_store(self, self.spec)

    def forward_loss(self, example: DrivingExample, per_sample=False):
        """Forward + next-token CE on the answer tokens + smooth-L1 on route / speed waypoints."""
        if torch.is_grad_enabled():
            self.param_store()
        adaptor_dict = self.adaptors(example)
        feats, _ = self.forward_model(example.driving_input, adaptor_dict, driving_labels=example.driving_label, want_logits=False)
        loss_dict = self.adaptors.compute_loss(feats, None, adaptor_dict, example)
        losses = {k: v for k, v in loss_dict.items() if k.endswith("loss")}
        logs = {k: v for k, v in loss_dict.items() if k.endswith("log")}
        if per_sample:
            return losses, {k: v for k, v in loss_dict.items() if not k.endswith("loss") and not k.endswith("log")}
        return summarise_losses(losses), logs

    def training_step(self, batch: DrivingExample, _batch_idx: int = 0):
        output, _ = self.forward_loss(batch)
        self.log_training_output(output, "train")
        self.log("train/loss", output.loss, on_step=True, on_epoch=True, prog_bar=True, logger=True)
        return {"loss": output.loss, "outputs": output}

    def validation_step(self, batch: DrivingExample, _batch_idx: int = 0):
        output, _ = self.forward_loss(batch)
        self.log_training_output(output, "val")
        self.log("val/loss", output.loss, on_step=False, on_epoch=True, prog_bar=True, logger=True)
        return {"loss": output.loss, "outputs": output}

    def predict_step(self, batch: DrivingExample, _batch_idx: int = 0):
        speed_wps, route, language = self.forward(batch, return_language=True)
        self.num_route_points = 20
        from simlingo_b200.postprocess import equal_spacing_route  # one launch for the batch instead of the per-item numpy loop (:290-295)
        route = equal_spacing_route(route, self.num_route_points)
        label = batch.driving_label
        record = {
            "waypoints": [speed_wps], "route": [route], "language": list(language),
            "waypoints_gt": [label.waypoints], "route_gt": [label.path], "language_gt": list(label.answer.language_string),
            "prompt": list(batch.driving_input.prompt.language_string), "path": decode_uint8(batch.run_id),
            "qa_templates": batch.qa_templates, "eval_infos": label.eval_infos,
        }
        if not self.prediction:
            self.prediction = record
        else:
            for k, v in record.items():
                if isinstance(v, list):
                    self.prediction[k].extend(v)
        return speed_wps, route, language, label.waypoints, label.path, label.answer.language_string

    def equal_spacing_route(self, points):
        """Re-samples the predicted route at 1 m arc-length steps (20 points)."""
        route = np.concatenate((np.zeros_like(points[:1]), points))
        prev = np.roll(route, 1, axis=0)
        prev[0] = prev[1]
        arc = np.cumsum(np.linalg.norm(route - prev, axis=1))
        arc += np.arange(0, len(arc)) * 1e-4  # strictly increasing for np.interp
        grid = np.arange(0, 20, 1)
        return np.array([np.interp(grid, arc, route[:, 0]), np.interp(grid, arc, route[:, 1])]).T

    def log_training_output(self, training_output: TrainingOutput, mode: str, dataset: Optional[str] = None):
        losses = {k: n.detach() for k, n in training_output.loss_averages.items()}
        counts = {k: n.detach().sum() for k, n in training_output.loss_counts.items()}
        losses["loss"], counts["loss"] = training_output.loss.detach(), 1
        for k, v in sorted(losses.items()):
            self.log(f"{mode}_losses/{k}", v, batch_size=counts[k], sync_dist=True, add_dataloader_idx=False)

    def configure_gradient_clipping(self, optimizer, gradient_clip_val=None, gradient_clip_algorithm=None):
        """Lightning hook (``Trainer(gradient_clip_val=0.3)``, reference train.py:206): the global-norm clip is folded
        into the fused AdamW kernel instead of a separate pass over the gradients."""
        opt = getattr(optimizer, "optimizer", optimizer)
        if gradient_clip_algorithm not in (None, "norm"):
            raise NotImplementedError("only global-norm clipping is implemented")
        opt.max_grad_norm = float(gradient_clip_val or 0.0)

    def configure_optimizers(self):
        """AdamW(lr, weight_decay, betas) over all parameters + per-step OneCycleLR (reference :718-732), realised
        by the fused multi-tensor AdamW kernel with fp32 master weights."""
        from simlingo_b200.optim import FusedAdamW
        # all parameters, frozen ones included, exactly as the reference does: the optimizer state indices then line up with a
        # torch.optim.AdamW checkpoint of the reference (FusedAdamW.state_dict / load_state_dict use that layout)
        optimizer = FusedAdamW(list(self.parameters()), self.param_store(), lr=self.lr,
                               weight_decay=self.weight_decay, betas=tuple(self.betas),
                               max_grad_norm=float(getattr(self, "gradient_clip_val", 0.0) or 0.0))
        trainer = self.trainer
        max_steps = trainer.estimated_stepping_batches if trainer.max_steps == -1 else trainer.max_steps
        scheduler = torch.optim.lr_scheduler.OneCycleLR(optimizer, max_lr=self.lr, total_steps=max_steps, pct_start=self.pct_start)
        return {"optimizer": optimizer, "lr_scheduler": {"scheduler": scheduler, "frequency": 1, "interval": "step"}}
