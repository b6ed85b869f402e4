"""Static description of the SimLingo / InternVL2-1B hot path.

The reference builds the model with ``AutoModel.from_pretrained('OpenGVLab/InternVL2-1B',
trust_remote_code=True)`` (reference ``simlingo_training/models/encoder/internvl2_model.py:9`` and
``simlingo_training/models/language_model/llm.py:88``).  Neither the Hub nor its cache is reachable
offline, so the architecture constants live here ("Model constants").

Everything that depends on *names* (the ``state_dict`` key schema of
``simlingo_training.models.driving.DrivingModel``) is table driven from this one file so a mismatch
against a real checkpoint can be corrected in a single edit.
"""
from __future__ import annotations

import dataclasses
import hashlib
import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch


@dataclasses.dataclass(frozen=True)
class ModelSpec:
    # ---- InternViT-300M (UPSTREAM config.json: vision_config) ----
    vit_hidden: int = 1024
    vit_layers: int = 24
    vit_heads: int = 16
    vit_mlp: int = 4096
    patch: int = 14
    image: int = 448
    vit_eps: float = 1e-6
    # ---- projector (pixel_shuffle 0.5 'v2' + mlp1) ----
    downsample: float = 0.5
    proj_eps: float = 1e-5
    # ---- Qwen2-0.5B (UPSTREAM config.json: llm_config) ----
    llm_hidden: int = 896
    llm_layers: int = 24
    llm_heads: int = 14
    llm_kv_heads: int = 2
    head_dim: int = 64
    llm_mlp: int = 4864
    vocab: int = 151655
    rope_theta: float = 1.0e6
    rms_eps: float = 1e-6
    # ---- LoRA (reference config: lora_r=32, lora_alpha=64, lora_dropout=0.1, all-linear) ----
    lora_r: int = 32
    lora_alpha: int = 64
    lora_dropout: float = 0.1
    # ---- driving adaptor (reference adaptors.py:96-136) ----
    n_route: int = 20
    n_speed: int = 10
    head_mlp: int = 256
    wp_hidden: int = 256
    wp_hidden2: int = 512
    # ---- tokenizer facts (UPSTREAM tokenizer; treated as config) ----
    eos_id: int = 151645          # <|im_end|>
    img_start_id: int = 151646    # <img>
    img_end_id: int = 151647      # </img>
    img_context_id: int = 151648  # <IMG_CONTEXT>
    first_added_id: int = 151655  # first of SimLingo's 8 added special tokens
    target_point_id: int = 151662 # <TARGET_POINT>
    tiles_per_frame: int = 2

    # derived ---------------------------------------------------------------------------------
    @property
    def n_patches(self) -> int:
        return (self.image // self.patch) ** 2

    @property
    def vit_tokens(self) -> int:
        return self.n_patches + 1

    @property
    def grid(self) -> int:
        return self.image // self.patch

    @property
    def patch_k(self) -> int:
        return 3 * self.patch * self.patch

    @property
    def proj_in(self) -> int:
        return self.vit_hidden * int(1 / self.downsample) ** 2

    @property
    def tokens_per_tile(self) -> int:
        return int(self.n_patches * self.downsample * self.downsample)

    @property
    def n_queries(self) -> int:
        return self.n_route + self.n_speed

    @property
    def lora_scale(self) -> float:
        return self.lora_alpha / self.lora_r

    @property
    def kv_dim(self) -> int:
        return self.llm_kv_heads * self.head_dim

    @property
    def qkv_dim(self) -> int:
        return self.llm_heads * self.head_dim + 2 * self.kv_dim


INTERNVL2_1B = ModelSpec()


def tiny_spec(vit_layers: int = 2, llm_layers: int = 2, vocab: int = 4096) -> ModelSpec:
    """Same widths / head sizes as InternVL2-1B (so every CUDA kernel sees its real inner
    dimensions) but fewer layers and a small vocabulary, for CPU-oracle parity tests that must
    finish in seconds.  Special token ids are re-based just above the shrunken vocabulary in the
    same relative order as the real tokenizer."""
    base = vocab - 10
    return dataclasses.replace(
        INTERNVL2_1B,
        vit_layers=vit_layers,
        llm_layers=llm_layers,
        vocab=vocab,
        eos_id=base,
        img_start_id=base + 1,
        img_end_id=base + 2,
        img_context_id=base + 3,
        first_added_id=vocab,
        target_point_id=vocab + 7,
    )


# ------------------------------------------------------------------------------------------------
# state_dict schema.  name -> (shape, kind)
#   kind in {"matrix", "bias", "norm_w", "norm_b", "ls", "lora_a", "lora_b", "embed", "query",
#            "cls", "pos", "lm_head", "conv"}
# ------------------------------------------------------------------------------------------------
VIT_PREFIX = "vision_model.image_encoder.model.vision_model."
MLP1_PREFIX = "vision_model.image_encoder.model.mlp1."
LLM_PREFIX = "language_model.model.base_model.model."


def _lora_linear(d: "OrderedDict[str, Tuple[tuple, str]]", prefix: str, out_f: int, in_f: int,
                 bias: bool, r: int) -> None:
    d[prefix + "base_layer.weight"] = ((out_f, in_f), "matrix")
    if bias:
        d[prefix + "base_layer.bias"] = ((out_f,), "bias")
    d[prefix + "lora_A.default.weight"] = ((r, in_f), "lora_a")
    d[prefix + "lora_B.default.weight"] = ((out_f, r), "lora_b")


def state_dict_schema(spec: ModelSpec = INTERNVL2_1B, with_aliases: bool = True
                      ) -> "OrderedDict[str, Tuple[tuple, str]]":
    """Ordered ``name -> (shape, kind)`` table of ``DrivingModel.state_dict()``.

    ``with_aliases`` adds the three duplicate names the reference's module aliasing produces
    (``llm.py:91`` registers ``embed_tokens`` a second time on the causal-LM module and
    ``adaptors.py:227-229`` registers ``embed_tokens`` / ``lm_head`` on the language adaptor)."""
    H, L = spec.vit_hidden, spec.vit_layers
    d: "OrderedDict[str, Tuple[tuple, str]]" = OrderedDict()
    e = VIT_PREFIX + "embeddings."
    d[e + "class_embedding"] = ((1, 1, H), "cls")
    d[e + "position_embedding"] = ((1, spec.vit_tokens, H), "pos")
    d[e + "patch_embedding.weight"] = ((H, 3, spec.patch, spec.patch), "conv")
    d[e + "patch_embedding.bias"] = ((H,), "bias")
    for i in range(L):
        p = f"{VIT_PREFIX}encoder.layers.{i}."
        d[p + "ls1"] = ((H,), "ls")
        d[p + "ls2"] = ((H,), "ls")
        d[p + "attn.qkv.weight"] = ((3 * H, H), "matrix")
        d[p + "attn.qkv.bias"] = ((3 * H,), "bias")
        d[p + "attn.proj.weight"] = ((H, H), "matrix")
        d[p + "attn.proj.bias"] = ((H,), "bias")
        d[p + "mlp.fc1.weight"] = ((spec.vit_mlp, H), "matrix")
        d[p + "mlp.fc1.bias"] = ((spec.vit_mlp,), "bias")
        d[p + "mlp.fc2.weight"] = ((H, spec.vit_mlp), "matrix")
        d[p + "mlp.fc2.bias"] = ((H,), "bias")
        d[p + "norm1.weight"] = ((H,), "norm_w")
        d[p + "norm1.bias"] = ((H,), "norm_b")
        d[p + "norm2.weight"] = ((H,), "norm_w")
        d[p + "norm2.bias"] = ((H,), "norm_b")
    D = spec.llm_hidden
    d[MLP1_PREFIX + "0.weight"] = ((spec.proj_in,), "norm_w")
    d[MLP1_PREFIX + "0.bias"] = ((spec.proj_in,), "norm_b")
    d[MLP1_PREFIX + "1.weight"] = ((D, spec.proj_in), "matrix")
    d[MLP1_PREFIX + "1.bias"] = ((D,), "bias")
    d[MLP1_PREFIX + "3.weight"] = ((D, D), "matrix")
    d[MLP1_PREFIX + "3.bias"] = ((D,), "bias")

    m = LLM_PREFIX + "model."
    d[m + "embed_tokens.weight"] = ((spec.vocab, D), "embed")
    r = spec.lora_r
    for i in range(spec.llm_layers):
        p = f"{m}layers.{i}."
        _lora_linear(d, p + "self_attn.q_proj.", spec.llm_heads * spec.head_dim, D, True, r)
        _lora_linear(d, p + "self_attn.k_proj.", spec.kv_dim, D, True, r)
        _lora_linear(d, p + "self_attn.v_proj.", spec.kv_dim, D, True, r)
        _lora_linear(d, p + "self_attn.o_proj.", D, spec.llm_heads * spec.head_dim, False, r)
        _lora_linear(d, p + "mlp.gate_proj.", spec.llm_mlp, D, False, r)
        _lora_linear(d, p + "mlp.up_proj.", spec.llm_mlp, D, False, r)
        _lora_linear(d, p + "mlp.down_proj.", D, spec.llm_mlp, False, r)
        d[p + "input_layernorm.weight"] = ((D,), "norm_w")
        d[p + "post_attention_layernorm.weight"] = ((D,), "norm_w")
    d[m + "norm.weight"] = ((D,), "norm_w")
    d[LLM_PREFIX + "lm_head.weight"] = ((spec.vocab, D), "lm_head")
    if with_aliases:
        d[LLM_PREFIX + "embed_tokens.weight"] = ((spec.vocab, D), "alias:" + m + "embed_tokens.weight")
        d["adaptors.language.embed_tokens.weight"] = ((spec.vocab, D), "alias:" + m + "embed_tokens.weight")
        d["adaptors.language.lm_head.weight"] = ((spec.vocab, D), "alias:" + LLM_PREFIX + "lm_head.weight")

    a = "adaptors.driving."
    d[a + "query_embeds_wps"] = ((1, spec.n_route, D), "query")
    d[a + "query_embeds_speed"] = ((1, spec.n_speed, D), "query")
    hm = spec.head_mlp
    d[a + "route_head.0.weight"] = ((2 * hm, D), "matrix")
    d[a + "route_head.0.bias"] = ((2 * hm,), "bias")
    d[a + "route_head.2.weight"] = ((hm, 2 * hm), "matrix")
    d[a + "route_head.2.bias"] = ((hm,), "bias")
    d[a + "route_head.4.weight"] = ((2, hm), "matrix")
    d[a + "speed_wps_head.0.weight"] = ((hm, D), "matrix")
    d[a + "speed_wps_head.0.bias"] = ((hm,), "bias")
    d[a + "speed_wps_head.2.weight"] = ((2, hm), "matrix")
    w = "wp_encoder.mlp."
    d[w + "0.weight"] = ((spec.wp_hidden, 2), "wp_in")
    d[w + "0.bias"] = ((spec.wp_hidden,), "bias")
    d[w + "2.weight"] = ((spec.wp_hidden2, spec.wp_hidden), "matrix")
    d[w + "2.bias"] = ((spec.wp_hidden2,), "bias")
    d[w + "4.weight"] = ((D, spec.wp_hidden2), "matrix")
    d[w + "4.bias"] = ((D,), "bias")
    return d


def trainable(name: str) -> bool:
    """Reference semantics (SURVEY 8a note 5): ViT + mlp1 + LoRA A/B + heads/queries + wp_encoder
    train; Qwen2 base weights, embeddings and lm_head are frozen (PEFT marks every non-LoRA
    parameter of the wrapped model ``requires_grad=False``)."""
    if name.startswith(LLM_PREFIX) or name.startswith("adaptors.language."):
        return ".lora_A." in name or ".lora_B." in name
    return True


# ------------------------------------------------------------------------------------------------
# Deterministic synthetic weights (no checkpoint offline; SURVEY 8d "Weights")
# ------------------------------------------------------------------------------------------------
def _gen_for(name: str, seed: int) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:8], "little") & 0x7FFF_FFFF_FFFF_FFFF)
    return g


LMHEAD_SHIFT = 7919  # planted next-token walk: argmax after token t is (t + LMHEAD_SHIFT) % vocab


def init_state_dict(spec: ModelSpec = INTERNVL2_1B, seed: int = 0, dtype=torch.float32,
                    planted_lm_head: bool = True, with_aliases: bool = False
                    ) -> "OrderedDict[str, torch.Tensor]":
    """Random-init weights keyed by state_dict name; every value is exactly representable in
    bf16 so the fp32 oracle and the bf16 CUDA path start from identical numbers.

    Scales: matrices N(0, 0.02), biases N(0, 0.02), norm weights U(0.8, 1.2), norm
    biases N(0, 0.02), ls1/ls2 U(0.05, 0.2), LoRA A and B N(0, 0.02) with B non-zero so the LoRA
    branch is exercised, queries 0.02 N(0,1) (reference adaptors.py:112,129).

    ``planted_lm_head``: token embeddings are N(0, 1) and
    ``lm_head[(t + LMHEAD_SHIFT) % V] = embed[t] + N(0, 0.25)``.  With unit-scale embeddings the
    Qwen2 residual stream keeps >= 40 % of its norm in the last input token's embedding (each of
    the 48 sub-layers adds ~0.3 |e| of roughly orthogonal output), so greedy decoding walks
    ``t -> t + LMHEAD_SHIFT`` with a top-1/top-2 margin far above bf16 noise - needed because
    with i.i.d. random logits the expected margin over 151 655 classes is ~0.2 sigma and token
    identity would be a coin toss (SURVEY 7.2 "Greedy-token identity")."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    schema = state_dict_schema(spec, with_aliases=with_aliases)
    for name, (shape, kind) in schema.items():
        if kind.startswith("alias:"):
            out[name] = out[kind[len("alias:"):]]
            continue
        g = _gen_for(name, seed)
        if kind in ("matrix", "conv", "bias", "norm_b", "lora_a", "lora_b", "query", "cls", "pos"):
            t = torch.randn(shape, generator=g) * 0.02
        elif kind == "embed":
            t = torch.randn(shape, generator=g) * (1.0 if planted_lm_head else 0.02)
        elif kind == "wp_in":
            t = torch.randn(shape, generator=g) * 0.05
        elif kind == "norm_w":
            t = torch.rand(shape, generator=g) * 0.4 + 0.8
        elif kind == "ls":
            t = torch.rand(shape, generator=g) * 0.15 + 0.05
        elif kind == "lm_head":
            emb = out[LLM_PREFIX + "model.embed_tokens.weight"].float()
            if planted_lm_head:
                t = torch.roll(emb, LMHEAD_SHIFT, 0) + torch.randn(shape, generator=g) * 0.25
            else:
                t = torch.randn(shape, generator=g) * 0.02
        else:
            raise KeyError(kind)
        out[name] = t.to(torch.bfloat16).to(dtype)
    return out


# ------------------------------------------------------------------------------------------------
# Synthetic inputs (SURVEY 8d "Synthetic inputs")
# ------------------------------------------------------------------------------------------------
def synth_prompt_ids(spec: ModelSpec, batch: int, seed: int, n_text: int = 22, answer_len: int = 0,
                     last_token: int | None = None) -> torch.Tensor:
    """``[B, L]`` int64: 3 template ids + <img> + 512 x <IMG_CONTEXT> + </img> + n_text text ids
    + 2 x <TARGET_POINT> + 4 template ids (+ answer_len answer ids for training prompts).
    L = 545 for the driving prompt at n_text=22."""
    g = torch.Generator().manual_seed(seed)
    n_img = spec.tokens_per_tile * spec.tiles_per_frame
    hi = min(spec.vocab - 12, 151643)
    rows = []
    for _ in range(batch):
        head = torch.randint(0, hi, (3,), generator=g)
        text = torch.randint(0, hi, (n_text,), generator=g)
        tail = torch.randint(0, hi, (4,), generator=g)
        if last_token is not None and answer_len == 0:
            tail[-1] = last_token
        ans = torch.randint(0, hi, (answer_len,), generator=g)
        if last_token is not None and answer_len > 0:
            ans[-1] = last_token
        row = torch.cat([
            head, torch.tensor([spec.img_start_id]),
            torch.full((n_img,), spec.img_context_id), torch.tensor([spec.img_end_id]),
            text, torch.full((2,), spec.target_point_id), tail, ans,
        ])
        rows.append(row)
    return torch.stack(rows).long()


def synth_frames(spec: ModelSpec, batch: int, seed: int, dtype=torch.float32) -> torch.Tensor:
    """``camera_images [B, 1, NP, 3, 448, 448]`` ~ N(0,1) clipped to the ImageNet-normalised
    range, rounded through bf16 (the agent casts to bf16 at agent_simlingo.py:751)."""
    g = torch.Generator().manual_seed(seed + 101)
    x = torch.randn((batch, 1, spec.tiles_per_frame, 3, spec.image, spec.image), generator=g)
    return x.clamp_(-2.2, 2.7).to(torch.bfloat16).to(dtype)


def synth_camera(h: int, w: int, seed: int):
    """Deterministic uint8 [3, h, w] camera-like test image (numpy): smooth gradients + edges + noise, so that a
    resampler sees clipping and ringing.  h, w = 359, 1024 is the agent's cropped front camera
    (512 - 512 * 4.8 // 16 rows of a 1024 x 512 frame, reference agent_simlingo.py:470)."""
    import numpy as np
    rs = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([127 + 120 * np.sin(x / 37.0 + c) * np.cos(y / 23.0 - c) for c in range(3)])
    img += (((x // 16 + y // 16) % 2) * 90 - 45)[None]
    img += rs.randn(3, h, w) * 20
    return np.clip(np.round(img), 0, 255).astype(np.uint8)


def synth_placeholders(spec: ModelSpec, batch: int, seed: int):
    """``list[dict[token_id -> ndarray[2,2]]]`` as built at agent_simlingo.py:566-580."""
    g = torch.Generator().manual_seed(seed + 202)
    out = []
    for _ in range(batch):
        v = (torch.randn((2, 2), generator=g) * 10.0).to(torch.bfloat16).float().numpy()
        out.append({spec.target_point_id: v})
    return out


def synth_labels(spec: ModelSpec, batch: int, seed: int):
    """(waypoints [B,10,2], path [B,20,2]) as cumulative sums of positive steps."""
    g = torch.Generator().manual_seed(seed + 303)
    wps = (torch.rand((batch, spec.n_speed, 2), generator=g) * 1.5).cumsum(1)
    path = (torch.rand((batch, spec.n_route, 2), generator=g) * 1.0).cumsum(1)
    return wps, path


def flops_vit(spec: ModelSpec, tiles: int) -> float:
    N, D, H, d, I, LY = spec.vit_tokens, spec.vit_hidden, spec.vit_heads, 64, spec.vit_mlp, spec.vit_layers
    M = tiles * N
    patch = 2 * tiles * spec.n_patches * spec.patch_k * D
    lin = 2 * M * (3 * D * D + D * D + 2 * D * I)
    att = tiles * H * 4 * N * N * d
    return float(patch + LY * (lin + att))


def flops_proj(spec: ModelSpec, tiles: int) -> float:
    return float(2 * tiles * spec.tokens_per_tile * (spec.proj_in * spec.llm_hidden + spec.llm_hidden ** 2))


def flops_llm(spec: ModelSpec, L: int, B: int = 1, logits_rows: int = 0, lora: bool = True) -> float:
    D, KV, I, LY, H, d, V, r = (spec.llm_hidden, spec.kv_dim, spec.llm_mlp, spec.llm_layers, spec.llm_heads,
                                spec.head_dim, spec.vocab, spec.lora_r)
    M = B * L
    lin = 2 * M * (2 * D * D + 2 * D * KV + 3 * D * I)
    lo = 2 * M * r * ((D + D) + 2 * (D + KV) + (D + D) + 2 * (D + I) + (I + D)) if lora else 0
    att = B * H * 4 * L * L * d * 0.5
    return float(LY * (lin + lo + att) + 2 * logits_rows * D * V)


def flops_frame(spec: ModelSpec = INTERNVL2_1B, L: int = 575) -> float:
    """One teacher-forced frame: 1.898 TF at the 1B config."""
    return flops_vit(spec, 2) + flops_proj(spec, 2) + flops_llm(spec, L)
