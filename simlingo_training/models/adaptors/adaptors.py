"""Adaptors: turn a driving example into token embeddings and the transformer outputs into predictions / losses.

Drop-in for reference ``simlingo_training/models/adaptors/adaptors.py`` (class names, constructor arguments,
parameter names, dictionary keys).  Embedding lookup, the driving heads and the waypoint encoder run in the
sm_100a kernels of ``simlingo_b200`` when no gradient is required; under autograd (training) the tiny head MLPs
use torch ops (they are < 0.01 % of the step's FLOPs) while the transformer itself stays on the custom kernels."""
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from simlingo_b200 import lib as _lib
from simlingo_training.utils.custom_types import DrivingExample


def _needs_grad(*tensors: Tensor) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _mlp_fp32(seq: nn.Sequential, x: Tensor) -> Tensor:
    """The tiny adaptor MLPs under autograd: fp32 arithmetic on the bf16 parameters (the inference kernels
    ``slb_driving_heads`` / ``slb_wp_encoder`` also accumulate in fp32), so ReLU / SiLU gates do not flip against
    the fp32 oracle."""
    x = x.float()
    for m in seq:
        if isinstance(m, nn.Linear):
            x = F.linear(x, m.weight.float(), None if m.bias is None else m.bias.float())
        else:
            x = type(m)()(x)  # activation modules are declared inplace=True; apply a fresh out-of-place instance
    return x


def _unwrap_input(example):
    """Both ``DrivingExample`` and a bare ``DrivingInput`` are accepted (reference adaptors.py:144-147)."""
    return getattr(example, "driving_input", example)


# ---------------------------------------------------------------------------------------------------------
# helpers other reference modules import from here: cross_track_error, NormZeroOne, FocalLoss
# ---------------------------------------------------------------------------------------------------------
def cross_track_error(points: Tensor, path: Tensor) -> Tensor:
    """Distance of each point [b,n,2] to the poly-line ``path`` [b,m,2] (NaN = missing vertex), measured along the
    normal of the averaged segment direction at the closest vertex (reference adaptors.py:10-35)."""
    points, path = points.float(), path.float()
    rows = torch.arange(path.size(0), device=path.device)[:, None]
    nearest = torch.cdist(points, path).nan_to_num_(torch.inf).argmin(-1)
    prev_v = path[rows, (nearest - 1).clamp_min(0)]
    this_v = path[rows, nearest]
    next_v = path[rows, (nearest + 1).clamp_max(path.size(1) - 1)]
    direction = (next_v - this_v).nan_to_num_(0.0) + (this_v - prev_v).nan_to_num_(0.0)
    normal = torch.stack((direction[..., 1], -direction[..., 0]), dim=-1)
    normal = normal / normal.norm(p=2, dim=-1, keepdim=True).clamp_min(1e-2)
    return ((points - this_v) * normal).sum(-1).abs()


class NormZeroOne(nn.Module):
    """(x - lo) / (hi - lo) (reference adaptors.py:37-44)."""

    def __init__(self, min_max: Tuple[float, float]):
        super().__init__()
        self.register_buffer("min_max", torch.tensor(min_max, dtype=torch.float), persistent=False)

    def forward(self, x: Tensor) -> Tensor:
        lo, hi = self.min_max[0], self.min_max[1]
        return (x - lo) / (hi - lo)


class FocalLoss(nn.Module):
    """-(1 - p_t)^gamma log p_t (reference adaptors.py:46-61)."""

    def __init__(self, gamma: float = 0, size_average: bool = True):
        super().__init__()
        self.gamma = gamma
        self.size_average = size_average

    def forward(self, input: Tensor, target: Tensor) -> Tensor:
        logpt = F.log_softmax(input, dim=-1).gather(1, target.view(-1, 1)).view(-1)
        loss = -((1 - logpt.exp()) ** self.gamma) * logpt
        return loss.mean() if self.size_average else loss.sum()


# ---------------------------------------------------------------------------------------------------------
class WaypointInputAdaptor(nn.Module):
    """[B, N, 2] coordinates -> [B, N, token_size] embeddings: Linear-ReLU-Linear-ReLU-Linear
    (reference adaptors.py:64-93)."""

    def __init__(self, token_size: int = 258, hidden_size: int = 64, hidden_size2: int = 128,
                 norm_layer: Optional[nn.Module] = None):
        super().__init__()
        self.hidden_size = hidden_size
        self.norm_layer = norm_layer
        self.mlp = nn.Sequential(nn.Linear(2, hidden_size), nn.ReLU(True), nn.Linear(hidden_size, hidden_size2), nn.ReLU(True),
                                 nn.Linear(hidden_size2, token_size))

    def _kernel_ok(self, x: Tensor) -> bool:
        m = self.mlp
        return (x.is_cuda and m[0].weight.dtype == torch.bfloat16 and m[0].out_features == 256 and m[2].out_features == 512
                and m[4].out_features == 896 and not _needs_grad(x, m[0].weight))

    def forward(self, x: Tensor) -> Tensor:
        if self.norm_layer is not None:
            x = self.norm_layer(x)
        if self._kernel_ok(x):
            m = self.mlp
            ww = _lib.WpWeights(*[t.data_ptr() for t in (m[0].weight, m[0].bias, m[2].weight, m[2].bias, m[4].weight, m[4].bias)])
            flat = x.reshape(-1, 2).float().contiguous()
            return _lib.wp_encoder(flat, ww).view(*x.shape[:-1], 896)
        if x.is_cuda and self.mlp[0].weight.dtype == torch.bfloat16:
            return _mlp_fp32(self.mlp, x).to(torch.bfloat16)
        return self.mlp(x)


class DrivingAdaptor(nn.Module):
    """Learned route / speed-waypoint queries and their prediction heads (reference adaptors.py:96-221)."""

    def __init__(self, hidden_size: int, mlp_dim=256, predict_route_as_wps=False, speed_wps_mode=False):
        super().__init__()
        self.heads, self.queries, self.sizes, self.order = {}, {}, {}, []
        self.speed_wps_mode = speed_wps_mode
        self.predict_route_as_wps = predict_route_as_wps
        if predict_route_as_wps:
            self.future_waypoints = 20
            self.query_embeds_wps = nn.Parameter(0.02 * torch.randn((1, self.future_waypoints, hidden_size)))
            self.route_head = nn.Sequential(nn.Linear(hidden_size, mlp_dim * 2), nn.SiLU(True), nn.Linear(mlp_dim * 2, mlp_dim),
                                            nn.SiLU(True), nn.Linear(mlp_dim, 2, bias=False))
            self._register("route", self.query_embeds_wps, self.route_head, self.future_waypoints)
        if speed_wps_mode not in ("1d", "2d"):
            raise ValueError(f"speed_wps_mode must be '1d' or '2d', not {speed_wps_mode}")
        out_dim = 2 if speed_wps_mode == "2d" else 1
        self.future_speed_waypoints = 10
        self.query_embeds_speed = nn.Parameter(0.02 * torch.randn((1, self.future_speed_waypoints, hidden_size)))
        self.speed_wps_head = nn.Sequential(nn.Linear(hidden_size, mlp_dim), nn.SiLU(True), nn.Linear(mlp_dim, out_dim, bias=False))
        self._register("speed_wps", self.query_embeds_speed, self.speed_wps_head, self.future_speed_waypoints)

    def _register(self, name, query, head, size):
        self.queries[name], self.heads[name], self.sizes[name] = query, head, size
        self.order.append(name)

    def forward(self, driving_example: DrivingExample, **kwargs) -> Dict[str, Tensor]:
        b = _unwrap_input(driving_example).camera_images.shape[0]
        inputs = torch.cat([self.queries[k].expand(b, -1, -1) for k in self.order], dim=1)
        return {"inputs": inputs, "inputs_mask": torch.ones_like(inputs[:, :, 0], dtype=torch.bool)}

    def _kernel_ok(self, features: Tensor) -> bool:
        return (features.is_cuda and features.dtype == torch.bfloat16 and self.order == ["route", "speed_wps"]
                and self.speed_wps_mode == "2d" and features.shape[-1] == 896 and features.shape[1] >= 30
                and not _needs_grad(features, self.query_embeds_speed))

    def get_predictions(self, features: Tensor, logits: Optional[Tensor] = None) -> Dict:
        """features [B, >=30, H] (route queries first) -> {'route': [B,20,2], 'speed_wps': [B,10,2]}, each the
        cumulative sum of the per-query head outputs (reference adaptors.py:163-180)."""
        if self._kernel_ok(features):
            f = features[:, :30].contiguous()
            rh, sh = self.route_head, self.speed_wps_head
            hw = _lib.HeadsWeights(*[t.data_ptr() for t in (rh[0].weight, rh[0].bias, rh[2].weight, rh[2].bias, rh[4].weight,
                                                           sh[0].weight, sh[0].bias, sh[2].weight)])
            route, speed = _lib.driving_heads(f, 30 * 896, hw, f.shape[0])
            return {"route": route.to(features.dtype), "speed_wps": speed.to(features.dtype)}
        out, at = {}, 0
        fp32 = features.is_cuda and features.dtype == torch.bfloat16
        for name in self.order:
            n = self.sizes[name]
            f = features[:, at: at + n]
            out[name] = _mlp_fp32(self.heads[name], f).cumsum(1) if fp32 else self.heads[name](f).cumsum(1)
            at += n
        return out

    def compute_loss(self, adaptor_features: Tensor, adaptor_logits: Tensor, _inputs: Dict[str, Tensor],
                     example: DrivingExample) -> Dict[str, Tuple[Tensor, Tensor]]:
        """Smooth-L1 summed over (x, y) per waypoint between cumsum(head(features)) and the labels
        (reference adaptors.py:183-221)."""
        label = example.driving_label
        assert label is not None
        targets = {"route": label.path if self.predict_route_as_wps else None}
        if self.speed_wps_mode == "2d":
            targets["speed_wps"] = label.waypoints[:, : self.future_waypoints + 1]
        else:
            targets["speed_wps"] = label.waypoints_1d
        preds = self.get_predictions(adaptor_features)
        out = {}
        for name in self.order:
            loss = F.smooth_l1_loss(preds[name], targets[name], reduction="none").sum(-1)
            out[f"{name}_loss"] = (loss, torch.ones_like(loss, dtype=torch.long))
            out[f"{name}_prediction"] = preds[name]
            out[f"{name}_label"] = targets[name]
        return out


class LanguageAdaptor(nn.Module):
    """Token embedding lookup in, next-token cross-entropy out (reference adaptors.py:224-274)."""

    def __init__(self, language_model):
        super().__init__()
        lm = language_model.model
        self.embed_tokens = lm.embed_tokens
        if hasattr(lm, "lm_head"):
            self.lm_head = lm.lm_head
        elif hasattr(lm, "embed_out"):
            self.lm_head = lm.embed_out
        elif hasattr(lm.base_model.model, "output"):
            self.lm_head = lm.base_model.model.output
        else:
            raise ValueError("Language model must have `lm_head` or `embed_out` attribute.")

    def forward(self, example: DrivingExample, inference=False, **kwargs) -> Dict[str, Tensor]:
        driving_input = _unwrap_input(example)
        label = driving_input.prompt_inference if inference else driving_input.prompt
        ids = label.phrase_ids.long()
        inputs = self.embed_tokens(ids.clamp(min=0, max=self.embed_tokens.num_embeddings - 1))
        out = {"inputs": inputs, "inputs_mask": label.phrase_valid, "_ids": ids, "_ids_mask": label.loss_masking}
        if not inference and torch.is_grad_enabled() and ids.is_cuda and label.loss_masking is not None:
            # flat indices of the positions that carry a next-token label, taken now (one device sync while the GPU is
            # idle) instead of after the decoder pass: compute_loss can then be queued without waiting for the forward
            lab = torch.where(label.loss_masking, ids, -1)[:, 1:]
            out["_label_rows"] = lab.reshape(-1).ne(-1).nonzero().squeeze(1)
        return out

    def compute_loss(self, adaptor_features: Tensor, adaptor_logits: Tensor, inputs: Dict[str, Tensor],
                     example: DrivingExample) -> Dict[str, Tuple[Tensor, Tensor]]:
        """Per-token CE of position t against id t+1 where ``loss_masking`` is set (ignore elsewhere)."""
        del example
        labels = torch.where(inputs["_ids_mask"], inputs["_ids"], -1)[:, 1:]
        if adaptor_logits is None and inputs.get("_label_rows") is not None:
            # fused, sync-free path: label rows were located before the forward pass was queued
            idx = inputs["_label_rows"]
            rows = labels.ne(-1)
            loss = torch.zeros(labels.numel(), device=labels.device, dtype=torch.float32)
            if idx.numel() > 0:
                feats = adaptor_features[:, :-1].reshape(-1, adaptor_features.shape[-1]).index_select(0, idx)
                from simlingo_b200 import training as _tr
                loss = loss.index_put((idx,), _tr.lm_head_ce(feats, self.lm_head.weight, labels.reshape(-1).index_select(0, idx)))
            return {"language_loss": (loss.view_as(labels), rows)}
        if adaptor_logits is None:
            # fused path: logits only for the rows that carry a label (never materialises [B, L, vocab])
            rows = labels.ne(-1)
            loss = torch.zeros(labels.shape, device=labels.device, dtype=torch.float32)
            if rows.any():
                feats = adaptor_features[:, :-1][rows]
                if feats.is_cuda and self.lm_head.weight.dtype == torch.bfloat16:
                    from simlingo_b200 import training as _tr
                    row_loss = _tr.lm_head_ce(feats, self.lm_head.weight, labels[rows])
                else:
                    row_loss = F.cross_entropy(self.lm_head(feats).float(), labels[rows], reduction="none")
                loss = loss.masked_scatter(rows, row_loss)
            return {"language_loss": (loss, rows)}
        lg = adaptor_logits[:, :-1]
        loss = F.cross_entropy(lg.flatten(0, -2), labels.flatten(), ignore_index=-1, reduction="none").view_as(labels)
        return {"language_loss": (loss, labels.ne(-1))}


class AdaptorList(nn.Module):
    """Concatenates the token streams of all adaptors ("valid tokens first" per row) and routes the transformer
    outputs back to each adaptor's loss (reference adaptors.py:276-370)."""

    def __init__(self, driving: Optional[DrivingAdaptor] = None, language: Optional[LanguageAdaptor] = None):
        super().__init__()
        self.driving = driving
        self.language = language

    @property
    def adaptors(self):
        named = (("language", self.language), ("driving", self.driving))
        return {k: v for k, v in named if v is not None}

    def forward(self, example: DrivingExample, **kwargs) -> Dict[str, Tensor]:
        out: Dict[str, Tensor] = {}
        streams, masks = [], []
        for key, adaptor in self.adaptors.items():
            d = adaptor.forward(example, **kwargs)
            streams.append(d["inputs"])
            masks.append(d["inputs_mask"])
            for k, v in d.items():
                out[f"{key}_{k}"] = v
        inputs, mask = torch.cat(streams, dim=1), torch.cat(masks, dim=1)
        rows = torch.arange(inputs.size(0), device=inputs.device)[:, None]
        identity = torch.arange(inputs.size(1), device=inputs.device).expand(inputs.size(0), -1)
        # stable sort on validity: valid tokens keep their order and move in front of the padding
        order = mask[rows, identity].byte().argsort(dim=-1, descending=True, stable=True)
        perm = identity.gather(1, order)
        out["inputs"] = inputs[rows, perm]
        out["inputs_mask"] = mask[rows, perm]
        out["perm"] = perm
        out["split_sizes"] = torch.as_tensor([x.size(1) for x in streams])
        return out

    def split_outputs_by_adaptor(self, input_dict: Dict[str, Tensor], outputs: Tensor) -> Dict[str, Tensor]:
        inverse = input_dict["perm"].argsort(-1)
        rows = torch.arange(inverse.size(0), device=inverse.device)[:, None]
        parts = outputs[rows, inverse].split([int(x) for x in input_dict["split_sizes"]], dim=1)
        return {key: parts[i] for i, key in enumerate(self.adaptors.keys())}

    def compute_loss(self, features: Tensor, logits: Optional[Tensor], input_dict: Dict[str, Tensor],
                     example: DrivingExample) -> Dict[str, Tuple[Tensor, Tensor]]:
        feats = self.split_outputs_by_adaptor(input_dict, features)
        lgts = self.split_outputs_by_adaptor(input_dict, logits) if logits is not None else {k: None for k in feats}
        losses: Dict[str, Tuple[Tensor, Tensor]] = {}
        for key, adaptor in self.adaptors.items():
            own_inputs = {k[len(key) + 1:]: v for k, v in input_dict.items() if k.startswith(key + "_")}
            losses.update(adaptor.compute_loss(feats[key], lgts[key], own_inputs, example))
        return losses
