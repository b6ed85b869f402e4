"""Glue between the drop-in ``nn.Module`` containers and the kernel engine.

An ``Engine`` is built lazily from a module's live parameters (``state_dict(keep_vars=True)``, so aliases such as
``adaptors.language.lm_head.weight`` resolve to the same storage) and cached on the module.  ``DrivingModel``
shares one full engine with its sub-modules; a sub-module used on its own (``LLM``, ``VLMEncoderModel``) gets an
engine over just its part.  There is no CPU / eager fallback: parameters must be bf16 CUDA tensors."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from . import lib
from .engine import Engine
from .spec import LLM_PREFIX, ModelSpec

_KEY = "_slb_engine"


def attach_engine(module: nn.Module, engine: Engine) -> None:
    module.__dict__[_KEY] = engine


def engine_for(module: nn.Module, prefix: str, spec: ModelSpec) -> Engine:
    """Engine over ``module``'s parameters, whose reference key is ``prefix + <module-relative key>``."""
    eng: Optional[Engine] = module.__dict__.get(_KEY)
    if eng is None:
        sd = {prefix + k: v for k, v in module.state_dict(keep_vars=True).items()}
        bad = [k for k, v in sd.items() if not (v.is_cuda and v.dtype == torch.bfloat16)]
        if bad:
            v = sd[bad[0]]
            raise RuntimeError(
                f"simlingo_b200 runs on bf16 CUDA parameters only (no CPU / fp32 fallback): {bad[0]} is {v.dtype} on {v.device}. "
                "Build the model under torch.set_default_dtype(torch.bfloat16) and move it to the GPU, as "
                "team_code/agent_simlingo.py:213-222 does.")
        eng = Engine(sd, spec)
        attach_engine(module, eng)
    tr = module.__dict__.get("_slb_train_engine")
    if tr is not None and eng.generation_fn is None:
        eng.generation_fn = lambda st=tr.store: st.generation
    if eng._packed_version is not None:
        eng.refresh(force=False)
    return eng


def invalidate(module: nn.Module, in_place: bool = False) -> None:
    """Drops the cached engines after the parameters changed behind them.  ``in_place`` (``load_state_dict``: values
    overwritten, storage untouched) keeps the training engine — its flat store still backs the parameters, so optimizers
    and captured graphs built on it stay valid — and only refreshes what it derived from the weights."""
    for m in module.modules():
        m.__dict__.pop(_KEY, None)
        tr = m.__dict__.get("_slb_train_engine")
        if tr is None:
            continue
        if in_place and tr.still_attached():
            if m is tr.root:
                tr.weights_changed()
        else:
            m.__dict__.pop("_slb_train_engine", None)


def grad_mode(*tensors: Tensor) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


# ---- sub-module entry points ---------------------------------------------------------------------
def extract_feature(chat_model: nn.Module, pixel_values: Tensor) -> Tensor:
    from .spec import VIT_PREFIX
    prefix = VIT_PREFIX[: -len("vision_model.")]
    if grad_mode(pixel_values, chat_model.mlp1[1].weight):
        from . import training
        return training.extract_feature(chat_model, pixel_values)
    eng = engine_for(chat_model, prefix, chat_model.spec)
    px = pixel_values.to(torch.bfloat16).contiguous()
    T = px.shape[0]
    # .clone(): small tile counts come out of a CUDA graph's static buffer
    return eng.extract_feature_auto(px).clone().view(T, chat_model.spec.tokens_per_tile, chat_model.spec.llm_hidden)


def embedding_lookup(weight: Tensor, ids: Tensor) -> Tensor:
    if not (weight.is_cuda and weight.dtype == torch.bfloat16):
        raise RuntimeError("simlingo_b200: embedding table must be a bf16 CUDA tensor (no CPU fallback)")
    flat = ids.reshape(-1).to(device=weight.device, dtype=torch.int64).contiguous()
    out = lib.gather_rows(weight.detach(), flat)
    return out.view(*ids.shape, weight.shape[1])


def lm_head(weight: Tensor, x: Tensor) -> Tensor:
    if grad_mode(x):
        from . import training
        return training.linear_frozen(x, weight)
    flat = x.reshape(-1, x.shape[-1]).to(torch.bfloat16).contiguous()
    out = lib.gemm(flat, weight.detach(), out_fp32=True)
    return out.view(*x.shape[:-1], weight.shape[0])


def llm_forward(causal_lm: nn.Module, inputs_embeds: Tensor, attention_mask: Optional[Tensor], want_logits: bool,
                logits_dtype: Optional[torch.dtype] = None) -> Tuple[Tensor, Optional[Tensor]]:
    """Teacher-forced Qwen2 pass (positions = arange over the padded sequence, causal + key-padding mask).
    Returns (features after the final RMSNorm [B,L,H], logits [B,L,V] in the model dtype or None)."""
    spec = causal_lm.spec
    if grad_mode(inputs_embeds) or (torch.is_grad_enabled() and any(p.requires_grad for p in causal_lm.parameters())
                                    and causal_lm.training):
        from . import training
        return training.llm_forward(causal_lm, inputs_embeds, attention_mask, want_logits)
    eng = engine_for(causal_lm, LLM_PREFIX, spec)
    B, L, D = inputs_embeds.shape
    mask = None if attention_mask is None else attention_mask.to(torch.bool)
    feats, _ = eng.forward_model(inputs_embeds.to(torch.bfloat16).contiguous(), mask)
    logits = None
    if want_logits:
        logits = eng.logits(feats.reshape(B * L, D)).view(B, L, spec.vocab).to(logits_dtype or inputs_embeds.dtype)
    return feats, logits
