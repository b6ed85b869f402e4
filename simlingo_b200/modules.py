"""Parameter containers that reproduce the *module tree* (and therefore the ``state_dict`` keys) of what the
reference obtains from ``AutoModel.from_pretrained('OpenGVLab/InternVL2-1B', trust_remote_code=True)`` plus
``peft.get_peft_model(..., target_modules='all-linear')`` - neither of which is importable offline.

The containers hold parameters only; the math is done by ``simlingo_b200.engine.Engine`` (inference) and
``simlingo_b200.training`` (autograd) on the hand-written kernels.  Attribute names follow UPSTREAM
``modeling_intern_vit.py`` / ``modeling_internvl_chat.py`` / ``transformers.models.qwen2`` / ``peft.tuners.lora``.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional

import torch
from torch import Tensor, nn

from .spec import INTERNVL2_1B, LLM_PREFIX, ModelSpec

_VARIANTS: Dict[str, ModelSpec] = {"OpenGVLab/InternVL2-1B": INTERNVL2_1B}


def register_variant(name: str, spec: ModelSpec) -> None:
    """Lets tests instantiate the drop-in classes on a reduced ``ModelSpec`` under their own variant name (the
    name must contain 'internvl2' to pass the reference's variant check, vlm.py:22)."""
    _VARIANTS[name] = spec


def spec_for_variant(name: str) -> ModelSpec:
    if name in _VARIANTS:
        return _VARIANTS[name]
    for k, v in _VARIANTS.items():
        if k.lower() == name.lower():
            return v
    raise ValueError(f"simlingo_b200 only implements {sorted(_VARIANTS)}; got variant {name!r}")


def _normal(*shape, std=0.02):
    return nn.Parameter(torch.empty(*shape).normal_(0.0, std))


def _linear(i: int, o: int, bias: bool = True) -> nn.Linear:
    lin = nn.Linear(i, o, bias=bias)
    with torch.no_grad():
        lin.weight.normal_(0.0, 0.02)
        if bias:
            lin.bias.zero_()
    return lin


# ------------------------------------------------------------------------------------------------
# InternViT-300M + mlp1
# ------------------------------------------------------------------------------------------------
class InternVisionEmbeddings(nn.Module):
    def __init__(self, s: ModelSpec):
        super().__init__()
        self.class_embedding = _normal(1, 1, s.vit_hidden, std=1.0)
        self.patch_embedding = nn.Conv2d(3, s.vit_hidden, kernel_size=s.patch, stride=s.patch)
        self.position_embedding = _normal(1, s.vit_tokens, s.vit_hidden, std=1.0)


class _InternAttention(nn.Module):
    def __init__(self, s: ModelSpec):
        super().__init__()
        self.qkv = _linear(s.vit_hidden, 3 * s.vit_hidden)
        self.proj = _linear(s.vit_hidden, s.vit_hidden)


class _InternMLP(nn.Module):
    def __init__(self, s: ModelSpec):
        super().__init__()
        self.fc1 = _linear(s.vit_hidden, s.vit_mlp)
        self.fc2 = _linear(s.vit_mlp, s.vit_hidden)


class InternVisionEncoderLayer(nn.Module):
    def __init__(self, s: ModelSpec):
        super().__init__()
        self.attn = _InternAttention(s)
        self.mlp = _InternMLP(s)
        self.norm1 = nn.LayerNorm(s.vit_hidden, eps=s.vit_eps)
        self.norm2 = nn.LayerNorm(s.vit_hidden, eps=s.vit_eps)
        self.ls1 = nn.Parameter(torch.ones(s.vit_hidden))  # initializer_factor = 1.0
        self.ls2 = nn.Parameter(torch.ones(s.vit_hidden))


class InternVisionEncoder(nn.Module):
    def __init__(self, s: ModelSpec):
        super().__init__()
        self.layers = nn.ModuleList([InternVisionEncoderLayer(s) for _ in range(s.vit_layers)])
        self.gradient_checkpointing = True


class InternVisionModel(nn.Module):
    def __init__(self, s: ModelSpec):
        super().__init__()
        self.embeddings = InternVisionEmbeddings(s)
        self.encoder = InternVisionEncoder(s)


class InternVLChatModel(nn.Module):
    """What ``LingoInternVLModel.model`` is in the reference after ``language_model`` has been nulled
    (vlm.py:30-31): ``vision_model`` + ``mlp1`` + ``extract_feature``."""

    def __init__(self, s: ModelSpec):
        super().__init__()
        self.spec = s
        self.vision_model = InternVisionModel(s)
        self.mlp1 = nn.Sequential(nn.LayerNorm(s.proj_in, eps=s.proj_eps), _linear(s.proj_in, s.llm_hidden), nn.GELU(),
                                  _linear(s.llm_hidden, s.llm_hidden))
        self.language_model = None
        self.config = SimpleNamespace(output_attentions=False, output_hidden_states=False, use_return_dict=True,
                                      downsample_ratio=s.downsample, ps_version="v2", select_layer=-1)
        self.downsample_ratio, self.ps_version, self.select_layer = s.downsample, "v2", -1

    def extract_feature(self, pixel_values: Tensor) -> Tensor:
        """[T,3,448,448] -> [T,256,896]: ViT -> drop CLS -> pixel_shuffle(0.5) -> mlp1 (UPSTREAM
        ``InternVLChatModel.extract_feature``)."""
        from . import runtime
        return runtime.extract_feature(self, pixel_values)

    def forward(self, *a, **k):
        raise NotImplementedError("the chat model is only used through extract_feature in SimLingo")


# ------------------------------------------------------------------------------------------------
# Qwen2-0.5B + PEFT-LoRA containers
# ------------------------------------------------------------------------------------------------
class LoraLinear(nn.Module):
    """``peft.tuners.lora.Linear`` key layout: base_layer / lora_A.default / lora_B.default."""

    def __init__(self, i: int, o: int, bias: bool, r: int, alpha: int, dropout: float):
        super().__init__()
        self.base_layer = _linear(i, o, bias)
        self.lora_dropout = nn.ModuleDict({"default": nn.Dropout(dropout)})
        self.lora_A = nn.ModuleDict({"default": nn.Linear(i, r, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, o, bias=False)})
        nn.init.zeros_(self.lora_B["default"].weight)  # PEFT: B starts at zero
        self.scaling = {"default": alpha / r}
        self.r = {"default": r}
        self.in_features, self.out_features = i, o
        for p in self.base_layer.parameters():
            p.requires_grad = False


class Qwen2RMSNorm(nn.Module):
    def __init__(self, d: int, eps: float):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.variance_epsilon = eps


class _Qwen2Attention(nn.Module):
    def __init__(self, s: ModelSpec, mk):
        super().__init__()
        d, kv = s.llm_hidden, s.kv_dim
        self.q_proj, self.k_proj, self.v_proj = mk(d, s.llm_heads * s.head_dim, True), mk(d, kv, True), mk(d, kv, True)
        self.o_proj = mk(s.llm_heads * s.head_dim, d, False)


class _Qwen2MLP(nn.Module):
    def __init__(self, s: ModelSpec, mk):
        super().__init__()
        self.gate_proj, self.up_proj = mk(s.llm_hidden, s.llm_mlp, False), mk(s.llm_hidden, s.llm_mlp, False)
        self.down_proj = mk(s.llm_mlp, s.llm_hidden, False)


class Qwen2DecoderLayer(nn.Module):
    def __init__(self, s: ModelSpec, mk):
        super().__init__()
        self.self_attn = _Qwen2Attention(s, mk)
        self.mlp = _Qwen2MLP(s, mk)
        self.input_layernorm = Qwen2RMSNorm(s.llm_hidden, s.rms_eps)
        self.post_attention_layernorm = Qwen2RMSNorm(s.llm_hidden, s.rms_eps)


class TokenEmbedding(nn.Module):
    """nn.Embedding stand-in whose lookup runs in ``slb_gather_rows`` (ids are clamped like adaptors.py:256)."""

    def __init__(self, n: int, d: int):
        super().__init__()
        self.num_embeddings, self.embedding_dim = n, d
        self.weight = _normal(n, d)

    def forward(self, ids: Tensor) -> Tensor:
        from . import runtime
        return runtime.embedding_lookup(self.weight, ids)


class LMHead(nn.Module):
    """bias-free Linear whose matmul runs in the tcgen05 GEMM (fp32 logits cast to the weight dtype)."""

    def __init__(self, d: int, n: int):
        super().__init__()
        self.in_features, self.out_features = d, n
        self.weight = _normal(n, d)

    def forward(self, x: Tensor) -> Tensor:
        from . import runtime
        return runtime.lm_head(self.weight, x)


class Qwen2Model(nn.Module):
    def __init__(self, s: ModelSpec, mk):
        super().__init__()
        self.embed_tokens = TokenEmbedding(s.vocab, s.llm_hidden)
        self.layers = nn.ModuleList([Qwen2DecoderLayer(s, mk) for _ in range(s.llm_layers)])
        self.norm = Qwen2RMSNorm(s.llm_hidden, s.rms_eps)


class CausalLMOutput(dict):
    """Minimal ``CausalLMOutputWithPast`` look-alike: ``out.hidden_states[-1]``, ``out.logits`` and ``out[0]``."""

    def __init__(self, logits, hidden_states):
        super().__init__(logits=logits, hidden_states=hidden_states)
        self.logits, self.hidden_states = logits, hidden_states

    def __getitem__(self, k):
        if isinstance(k, int):
            return (self.logits, self.hidden_states)[k]
        return super().__getitem__(k)


class Qwen2ForCausalLM(nn.Module):
    def __init__(self, s: ModelSpec, lora: bool):
        super().__init__()
        self.spec = s
        mk = (lambda i, o, b: LoraLinear(i, o, b, s.lora_r, s.lora_alpha, s.lora_dropout)) if lora else _linear
        self.model = Qwen2Model(s, mk)
        self.lm_head = LMHead(s.llm_hidden, s.vocab)
        self.config = SimpleNamespace(vocab_size=s.vocab, hidden_size=s.llm_hidden, max_position_embeddings=32768,
                                      num_hidden_layers=s.llm_layers, num_attention_heads=s.llm_heads,
                                      num_key_value_heads=s.llm_kv_heads, rope_theta=s.rope_theta)

    @property
    def base_model(self):  # HF: the decoder stack without the head (llm.py:91 reads base_model.embed_tokens)
        return self.model

    @property
    def dtype(self):
        return self.lm_head.weight.dtype

    def forward(self, inputs_embeds: Tensor = None, attention_mask: Optional[Tensor] = None, output_hidden_states: bool = True,
                position_ids: Optional[Tensor] = None, return_dict: bool = True, **_):
        """Teacher-forced pass; returns post-final-norm features as ``hidden_states[-1]`` and full-vocabulary logits
        (reference llm.py:133-141 / driving.py:217-223)."""
        if position_ids is not None:
            raise NotImplementedError("the reference always passes position_ids=None (arange over the padded sequence)")
        from . import runtime
        feats, logits = runtime.llm_forward(self, inputs_embeds, attention_mask, want_logits=True)
        return CausalLMOutput(logits, (feats,))


class _Forwarding(nn.Module):
    _inner_name = "model"

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            if name == self._inner_name:
                raise
            return getattr(super().__getattr__(self._inner_name), name)

    def forward(self, *a, **k):
        return getattr(self, self._inner_name)(*a, **k)


class LoraModel(_Forwarding):
    """``peft.tuners.lora.LoraModel``: holds the adapted network as ``.model``."""

    def __init__(self, model: Qwen2ForCausalLM):
        super().__init__()
        self.model = model


class PeftModelForCausalLM(_Forwarding):
    """``peft.PeftModel``: holds the ``LoraModel`` as ``.base_model`` => keys ``base_model.model.<hf key>``."""
    _inner_name = "base_model"

    def __init__(self, model: Qwen2ForCausalLM):
        super().__init__()
        self.base_model = LoraModel(model)
        for n, p in model.named_parameters():
            p.requires_grad = (".lora_A." in n) or (".lora_B." in n)

    def print_trainable_parameters(self):
        tr = sum(p.numel() for p in self.parameters() if p.requires_grad)
        al = sum(p.numel() for p in self.parameters())
        print(f"trainable params: {tr:,d} || all params: {al:,d} || trainable%: {100 * tr / al:.4f}")
