"""ORACLE - test infrastructure, not product code.

fp32 CPU restatement (plain PyTorch, no custom kernels) of the SimLingo VLA hot path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; the product (``simlingo_b200`` / ``simlingo_training``) never does.

PARITY PINNING.  The reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md sections 4 and 8c: "parity unpinned"), and its arithmetic lives in third-party code
that is absent from ``/root/reference`` (HF-Hub ``OpenGVLab/InternVL2-1B`` remote code at an
unpinned revision, ``transformers==4.46.3``, ``peft==0.13.2``).  The restatement below is
therefore pinned against outputs of the REFERENCE'S OWN CODE run in the authoring container and against the independent
implementations that *are* present in this image:

* ``driving.py`` / ``adaptors.py`` / ``internvl2_model.py`` / ``llm.py`` / ``utils.py`` of the reference, imported unmodified by
  ``tests/golden/make_golden.py`` / ``make_golden_grads.py`` -> ``tests/golden/reference_run.pt`` (forward, greedy tokens,
  losses) and ``reference_grads.pt`` (gradients) -> ``tests/test_oracle_pinning.py``.
* InternViT-300M + pixel-shuffle + ``mlp1``  <-> transformers' HF-native InternVL port (``InternVLVisionModel``,
  ``InternVLModel.pixel_shuffle``, ``InternVLMultiModalProjector``; weights mapped key by key)
  (``test_oracle_vit_and_projector_match_hf_internvl_port``).

* Qwen2 decoder  <-> ``transformers.Qwen2ForCausalLM`` 5.5.0, eager attention
  (``tests/test_oracle_pinning.py::test_llm_matches_hf_qwen2``) - this is the very class the
  reference calls at ``llm.py:133-141`` (other version).
* projector ``pixel_shuffle`` <-> the verbatim upstream algorithm restated from the
  look-alike ``vllm/model_executor/models/internvl.py:657-672`` against the closed form used here.
* LoRA <-> merged-weight identity ``W + (alpha/r) B A``.
* The in-repo logic (adaptors, placeholder substitution, greedy loop, losses) is restated line by
  line from the reference files cited in each docstring.

Every function takes ``sd``: a dict keyed by the reference ``state_dict`` names
(``simlingo_b200.spec.state_dict_schema``) holding fp32 tensors.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from simlingo_b200.spec import LLM_PREFIX, MLP1_PREFIX, VIT_PREFIX, ModelSpec

SD = Dict[str, Tensor]


# ================================================================================================
# InternViT-300M  (UPSTREAM modeling_intern_vit.py; look-alike
#   /opt/prime-rl/.venv/lib/python3.12/site-packages/vllm/model_executor/models/intern_vit.py)
# ================================================================================================
def vit_embeddings(sd: SD, spec: ModelSpec, pixel_values: Tensor) -> Tensor:
    """intern_vit.py:103-115: Conv2d(3,1024,k=14,s=14) -> flatten -> cat CLS -> + position
    embedding (the bicubic interpolate to the same 32x32 grid is the identity, SURVEY 8a note 9)."""
    e = VIT_PREFIX + "embeddings."
    x = F.conv2d(pixel_values, sd[e + "patch_embedding.weight"], sd[e + "patch_embedding.bias"],
                 stride=spec.patch)
    x = x.flatten(2).transpose(1, 2)
    cls = sd[e + "class_embedding"].expand(x.size(0), 1, -1)
    x = torch.cat([cls, x], dim=1)
    return x + sd[e + "position_embedding"]


def vit_layer(sd: SD, spec: ModelSpec, i: int, x: Tensor) -> Tensor:
    """intern_vit.py:341-350: x + attn(norm1(x))*ls1 ; x + mlp(norm2(x))*ls2.
    Attention (:237-247): qkv Linear -> (3, heads, d) split -> softmax(q k^T / sqrt(d)) v -> proj.
    MLP (:279-284): fc1 -> erf-GELU -> fc2."""
    p = f"{VIT_PREFIX}encoder.layers.{i}."
    B, N, C = x.shape
    H = spec.vit_heads
    h = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], spec.vit_eps)
    qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    att = (q * (C // H) ** -0.5) @ k.transpose(-2, -1)
    att = att.softmax(dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, N, C)
    o = F.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    x = x + o * sd[p + "ls1"]
    h = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], spec.vit_eps)
    h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    h = F.gelu(h)
    h = F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + h * sd[p + "ls2"]


def vit_forward(sd: SD, spec: ModelSpec, pixel_values: Tensor, collect: Optional[list] = None) -> Tensor:
    x = vit_embeddings(sd, spec, pixel_values)
    for i in range(spec.vit_layers):
        x = vit_layer(sd, spec, i, x)
        if collect is not None:
            collect.append(x)
    return x


def pixel_shuffle_upstream(x: Tensor, scale_factor: float = 0.5) -> Tensor:
    """Verbatim algorithm of UPSTREAM InternVLChatModel.pixel_shuffle, ps_version 'v2'
    (look-alike internvl.py:657-672).  x: [n, w, h, c]."""
    n, w, h, c = x.size()
    x = x.view(n, w, int(h * scale_factor), int(c / scale_factor))
    x = x.permute(0, 2, 1, 3).contiguous()
    x = x.view(n, int(h * scale_factor), int(w * scale_factor), int(c / (scale_factor * scale_factor)))
    x = x.permute(0, 2, 1, 3).contiguous()
    return x


def pixel_shuffle_closed_form(tokens: Tensor, grid: int) -> Tensor:
    """SURVEY 8a note 8: out[t = (grid/2)*i + j, q*C + c] = in[(2i + q//2)*grid + (2j + q%2), c].
    tokens: [n, grid*grid, C] -> [n, grid*grid/4, 4C]."""
    n, _, C = tokens.shape
    g2 = grid // 2
    x = tokens.view(n, g2, 2, g2, 2, C)            # (i, qi, j, qj)
    x = x.permute(0, 1, 3, 2, 4, 5)                # (i, j, qi, qj, C)
    return x.reshape(n, g2 * g2, 4 * C)


def extract_feature(sd: SD, spec: ModelSpec, pixel_values: Tensor, collect: Optional[list] = None) -> Tensor:
    """UPSTREAM InternVLChatModel.extract_feature (look-alike internvl.py:674-684):
    ViT -> drop CLS -> pixel_shuffle(0.5) -> mlp1 (LN(4096) -> Linear -> GELU -> Linear)."""
    x = vit_forward(sd, spec, pixel_values, collect)[:, 1:, :]
    x = pixel_shuffle_closed_form(x, spec.grid)
    x = F.layer_norm(x, (spec.proj_in,), sd[MLP1_PREFIX + "0.weight"], sd[MLP1_PREFIX + "0.bias"], spec.proj_eps)
    x = F.linear(x, sd[MLP1_PREFIX + "1.weight"], sd[MLP1_PREFIX + "1.bias"])
    x = F.gelu(x)
    return F.linear(x, sd[MLP1_PREFIX + "3.weight"], sd[MLP1_PREFIX + "3.bias"])


# ================================================================================================
# Qwen2-0.5B + LoRA  (transformers modeling_qwen2.py; PEFT LoRA semantics, SURVEY 8a note 5)
# ================================================================================================
def lora_linear(sd: SD, prefix: str, x: Tensor, scale: float, dropout_p: float = 0.0,
                training: bool = False) -> Tensor:
    """PEFT lora.Linear.forward: base(x) + scale * B(A(dropout(x))); scale = alpha / r."""
    y = F.linear(x, sd[prefix + "base_layer.weight"], sd.get(prefix + "base_layer.bias"))
    a = sd.get(prefix + "lora_A.default.weight")
    if a is None:
        return y
    xd = F.dropout(x, dropout_p, training) if (training and dropout_p > 0) else x
    return y + scale * F.linear(F.linear(xd, a), sd[prefix + "lora_B.default.weight"])


def rms_norm(x: Tensor, w: Tensor, eps: float) -> Tensor:
    """modeling_qwen2.py:258-263 (fp32 variance, cast, then * weight)."""
    v = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x.float() * torch.rsqrt(v + eps)).to(x.dtype)


def rope_cos_sin(spec: ModelSpec, positions: Tensor) -> Tuple[Tensor, Tensor]:
    """modeling_qwen2.py:102-123: inv_freq = theta^(-2i/d); emb = cat(freqs, freqs)."""
    d = spec.head_dim
    inv = 1.0 / (spec.rope_theta ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
    fr = positions.float()[..., None] * inv
    emb = torch.cat([fr, fr], dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x: Tensor) -> Tensor:
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def llm_layer(sd: SD, spec: ModelSpec, i: int, x: Tensor, cos: Tensor, sin: Tensor, bias_mask: Tensor,
              training: bool = False) -> Tensor:
    """modeling_qwen2.py:206-246 (decoder layer), :161-204 (attention, eager path :35-48 repeat_kv)."""
    p = f"{LLM_PREFIX}model.layers.{i}."
    B, L, D = x.shape
    H, KV, d = spec.llm_heads, spec.llm_kv_heads, spec.head_dim
    sc, dp = spec.lora_scale, spec.lora_dropout
    h = rms_norm(x, sd[p + "input_layernorm.weight"], spec.rms_eps)
    q = lora_linear(sd, p + "self_attn.q_proj.", h, sc, dp, training).view(B, L, H, d).transpose(1, 2)
    k = lora_linear(sd, p + "self_attn.k_proj.", h, sc, dp, training).view(B, L, KV, d).transpose(1, 2)
    v = lora_linear(sd, p + "self_attn.v_proj.", h, sc, dp, training).view(B, L, KV, d).transpose(1, 2)
    c, s = cos[:, None], sin[:, None]
    q = q * c + rotate_half(q) * s
    k = k * c + rotate_half(k) * s
    k = k.repeat_interleave(H // KV, dim=1)
    v = v.repeat_interleave(H // KV, dim=1)
    att = (q @ k.transpose(2, 3)) * (d ** -0.5) + bias_mask
    att = att.softmax(dim=-1, dtype=torch.float32).to(q.dtype)
    o = (att @ v).transpose(1, 2).reshape(B, L, H * d)
    x = x + lora_linear(sd, p + "self_attn.o_proj.", o, sc, dp, training)
    h = rms_norm(x, sd[p + "post_attention_layernorm.weight"], spec.rms_eps)
    g = lora_linear(sd, p + "mlp.gate_proj.", h, sc, dp, training)
    u = lora_linear(sd, p + "mlp.up_proj.", h, sc, dp, training)
    return x + lora_linear(sd, p + "mlp.down_proj.", F.silu(g) * u, sc, dp, training)


def llm_forward(sd: SD, spec: ModelSpec, inputs_embeds: Tensor, attention_mask: Optional[Tensor] = None,
                logits_rows: Optional[str] = "all", training: bool = False,
                collect: Optional[list] = None) -> Tuple[Tensor, Optional[Tensor]]:
    """``LLM.forward`` (reference llm.py:126-143): HF ``Qwen2ForCausalLM(inputs_embeds=...,
    attention_mask=..., output_hidden_states=True, position_ids=None)`` -> (hidden_states[-1]
    = post-final-norm features, logits).  ``position_ids=None`` => arange over the padded
    sequence regardless of the mask (SURVEY 8a note 1).  The 2-D mask is a key-padding mask
    combined with the causal mask, as HF builds it."""
    B, L, _ = inputs_embeds.shape
    pos = torch.arange(L)[None].expand(B, L)
    cos, sin = rope_cos_sin(spec, pos)
    neg = torch.finfo(torch.float32).min
    causal = torch.full((L, L), neg).triu(1)[None, None].expand(B, 1, L, L).clone()
    if attention_mask is not None:
        pad = ~attention_mask.bool()
        causal = causal.masked_fill(pad[:, None, None, :], neg)
    x = inputs_embeds
    for i in range(spec.llm_layers):
        x = llm_layer(sd, spec, i, x, cos, sin, causal, training)
        if collect is not None:
            collect.append(x)
    feats = rms_norm(x, sd[LLM_PREFIX + "model.norm.weight"], spec.rms_eps)
    logits = F.linear(feats, sd[LLM_PREFIX + "lm_head.weight"]) if logits_rows == "all" else None
    return feats, logits


def greedy_sample(sd: SD, spec: ModelSpec, input_embeds: Tensor, max_new_tokens: int,
                  eos_token_id: Optional[int], attention_mask: Tensor,
                  margins: Optional[list] = None) -> Tuple[Tensor, Tensor]:
    """``LLM.greedy_sample`` (reference llm.py:178-250), temperature<=0 branch: full re-forward
    of the growing sequence every step (no KV cache), logits of the last row through
    ``F.linear(last_hidden, lm_head.weight)``, argmax, append ``F.embedding(next_token)``
    (the EOS embedding is appended too), EOS pre-filled ``sampled_tokens``, early exit when every
    row has produced EOS."""
    emb_w = sd[LLM_PREFIX + "model.embed_tokens.weight"]
    head_w = sd[LLM_PREFIX + "lm_head.weight"]
    B = input_embeds.size(0)
    sampled = torch.empty((B, max_new_tokens), dtype=torch.long)
    if eos_token_id is not None:
        sampled.fill_(eos_token_id)
    incomplete = torch.ones(B, dtype=torch.bool)
    attention_mask = attention_mask.clone()
    for i in range(max_new_tokens):
        feats, _ = llm_forward(sd, spec, input_embeds, attention_mask, logits_rows=None)
        logits = F.linear(feats[:, -1], head_w)
        if margins is not None:
            top2 = logits.topk(2, dim=-1).values
            margins.append((top2[:, 0] - top2[:, 1]).min().item())
        nxt = logits.argmax(dim=-1)
        input_embeds = torch.cat([input_embeds, F.embedding(nxt.unsqueeze(1), emb_w)], dim=1)
        attention_mask = torch.cat([attention_mask, torch.ones((B, 1), dtype=attention_mask.dtype)], dim=1)
        sampled[incomplete, i] = nxt[incomplete]
        if eos_token_id is not None:
            incomplete = sampled[:, i] != eos_token_id
            if not incomplete.any():
                sampled = sampled[:, : i + 1]
                break
    return sampled, input_embeds


# ================================================================================================
# Adaptors (reference simlingo_training/models/adaptors/adaptors.py)
# ================================================================================================
def mlp_seq(sd: SD, prefix: str, idxs: List[int], x: Tensor, act) -> Tensor:
    for n, i in enumerate(idxs):
        x = F.linear(x, sd[f"{prefix}{i}.weight"], sd.get(f"{prefix}{i}.bias"))
        if n + 1 < len(idxs):
            x = act(x)
    return x


def wp_encoder(sd: SD, x: Tensor) -> Tensor:
    """WaypointInputAdaptor (adaptors.py:64-93): Linear(2,256) ReLU Linear(256,512) ReLU Linear(512,896)."""
    return mlp_seq(sd, "wp_encoder.mlp.", [0, 2, 4], x, F.relu)


def driving_predictions(sd: SD, spec: ModelSpec, features: Tensor) -> Dict[str, Tensor]:
    """DrivingAdaptor.get_predictions (adaptors.py:163-180): route head on the first 20 query
    features, speed-waypoint head on the next 10, each followed by ``.cumsum(1)``."""
    r = mlp_seq(sd, "adaptors.driving.route_head.", [0, 2, 4], features[:, : spec.n_route], F.silu).cumsum(1)
    s = mlp_seq(sd, "adaptors.driving.speed_wps_head.", [0, 2],
                features[:, spec.n_route: spec.n_route + spec.n_speed], F.silu).cumsum(1)
    return {"route": r, "speed_wps": s}


def language_embed(sd: SD, spec: ModelSpec, ids: Tensor) -> Tensor:
    """LanguageAdaptor.forward (adaptors.py:256): embed_tokens(ids.clamp(0, V-1))."""
    return F.embedding(ids.clamp(min=0, max=spec.vocab - 1), sd[LLM_PREFIX + "model.embed_tokens.weight"])


def driving_queries(sd: SD, batch: int) -> Tensor:
    """DrivingAdaptor.forward (adaptors.py:139-161): cat(route queries, speed queries).expand(B)."""
    q = torch.cat([sd["adaptors.driving.query_embeds_wps"], sd["adaptors.driving.query_embeds_speed"]], dim=1)
    return q.expand(batch, -1, -1)


def adaptor_list_forward(sd: SD, spec: ModelSpec, ids: Tensor, ids_valid: Tensor, ids_mask: Tensor) -> Dict:
    """AdaptorList.forward (adaptors.py:301-331): concat [language | driving], stable "valid
    first" permutation."""
    B = ids.size(0)
    lang = language_embed(sd, spec, ids)
    drv = driving_queries(sd, B)
    inputs = torch.cat([lang, drv], dim=1)
    mask = torch.cat([ids_valid.bool(), torch.ones((B, drv.size(1)), dtype=torch.bool)], dim=1)
    ar = torch.arange(B)[:, None]
    rand_perm = torch.arange(inputs.size(1)).expand(B, -1)
    valid_perm = mask[ar, rand_perm].byte().argsort(dim=-1, descending=True, stable=True)
    perm = rand_perm.gather(1, valid_perm)
    return {
        "language_inputs": lang, "language_inputs_mask": ids_valid.bool(), "language__ids": ids,
        "language__ids_mask": ids_mask, "driving_inputs": drv,
        "inputs": inputs[ar, perm], "inputs_mask": mask[ar, perm], "perm": perm,
        "split_sizes": torch.as_tensor([lang.size(1), drv.size(1)]),
    }


def replace_placeholder_tokens(sd: SD, spec: ModelSpec, ad: Dict, pixel_values: Tensor,
                               placeholder_values: List[dict]) -> Dict:
    """LingoInternVLModel.replace_placeholder_tokens (reference internvl2_model.py:17-144):
    (i) wp_encoder embeddings overwrite the run starting at the first occurrence of each added
    special id (:54-91); (ii) extract_feature over [B*NP,3,448,448] (:102-114); (iii) rows with
    id == <IMG_CONTEXT> are replaced by the ViT embeddings in order (:119-131); (iv) the
    language part is copied into the permuted ``inputs`` from each row's start index (:138-142)."""
    emb = ad["language_inputs"].clone()
    ids = ad["language__ids"]
    B, L = ids.shape
    special = sorted(set(ids[ids >= spec.first_added_id].tolist()))
    if special and len(placeholder_values) > 0:
        for b in range(B):
            for sid in special:
                pos = (ids[b] == sid).nonzero()
                if pos.numel() == 0:
                    continue
                start = int(pos[0])
                if start == 0:  # reference: first_occurrences.nonzero() drops index 0 (:78)
                    continue
                coords = torch.as_tensor(placeholder_values[b][sid], dtype=torch.float32)
                w = wp_encoder(sd, coords.unsqueeze(0)).squeeze(0)
                emb[b, start:start + coords.size(0)] = w
    if pixel_values is not None and L != 1 and pixel_values.size(0) > 0:
        BS, T, NP, C, H, W = pixel_values.shape
        assert T == 1
        vit = extract_feature(sd, spec, pixel_values.reshape(BS * NP, C, H, W)).reshape(-1, emb.size(-1))
        flat = emb.reshape(B * L, -1).clone()
        sel = ids.reshape(B * L) == spec.img_context_id
        flat[sel] = flat[sel] * 0.0 + vit
        emb = flat.reshape(B, L, -1)
    ad = dict(ad)
    ad["language_inputs"] = emb
    inputs = ad["inputs"].clone()
    start_id = ad["perm"][:, 0]
    for b, i in enumerate(start_id.tolist()):
        inputs[b, : L - i] = emb[b, i:]
    ad["inputs"] = inputs
    return ad


# ================================================================================================
# DrivingModel (reference simlingo_training/models/driving.py)
# ================================================================================================
def driving_forward(sd: SD, spec: ModelSpec, camera_images: Tensor, phrase_ids: Tensor, phrase_valid: Tensor,
                    placeholder_values: List[dict], max_new_tokens: int = 100,
                    eos_token_id: Optional[int] = None, margins: Optional[list] = None):
    """``DrivingModel.forward`` inference path (driving.py:104-187): per batch item greedy decode
    (no cache), then one more full pass over [prompt + generated | 30 queries] *without*
    attention mask (:154-156), heads on the last 30 features.
    Returns (speed_wps [B,10,2], route [B,20,2], list of sampled token tensors)."""
    B = phrase_ids.size(0)
    ad = adaptor_list_forward(sd, spec, phrase_ids, phrase_valid, torch.zeros_like(phrase_valid))
    ad = replace_placeholder_tokens(sd, spec, ad, camera_images, placeholder_values)
    speed, route, toks = [], [], []
    for b in range(B):
        emb = ad["language_inputs"][b:b + 1]
        mask = ad["language_inputs_mask"][b:b + 1]
        sampled, emb_after = greedy_sample(sd, spec, emb, max_new_tokens, eos_token_id, mask, margins)
        cat = torch.cat([emb_after, ad["driving_inputs"][b:b + 1]], dim=1)
        feats, _ = llm_forward(sd, spec, cat, None, logits_rows=None)
        pred = driving_predictions(sd, spec, feats[:, -spec.n_queries:])
        speed.append(pred["speed_wps"])
        route.append(pred["route"])
        toks.append(sampled[0])
    return torch.cat(speed), torch.cat(route), toks


def forward_model(sd: SD, spec: ModelSpec, ad: Dict, camera_images: Tensor, placeholder_values: List[dict],
                  training: bool = False, logits: bool = True):
    """``DrivingModel.forward_model`` (driving.py:190-233): teacher-forced single pass."""
    ad = replace_placeholder_tokens(sd, spec, ad, camera_images, placeholder_values)
    feats, lg = llm_forward(sd, spec, ad["inputs"], ad["inputs_mask"], "all" if logits else None, training)
    return ad, feats, lg


def split_outputs(ad: Dict, outputs: Tensor) -> Tuple[Tensor, Tensor]:
    """AdaptorList.split_outputs_by_adaptor (adaptors.py:357-370)."""
    inv = ad["perm"].argsort(-1)
    ar = torch.arange(inv.size(0))[:, None]
    outputs = outputs[ar, inv]
    a, b = [int(x) for x in ad["split_sizes"]]
    return outputs[:, :a], outputs[:, a:a + b]


def forward_loss(sd: SD, spec: ModelSpec, camera_images: Tensor, phrase_ids: Tensor, phrase_valid: Tensor,
                 loss_masking: Tensor, placeholder_values: List[dict], waypoints: Tensor, path: Tensor,
                 training: bool = False):
    """``DrivingModel.forward_loss`` (driving.py:236-261) + ``AdaptorList.compute_loss``
    (adaptors.py:333-355) + ``LanguageAdaptor.compute_loss`` (:259-274) +
    ``DrivingAdaptor.compute_loss`` (:183-221) + ``summarise_losses`` (models/utils.py:7-41).
    Returns (loss, dict of per-term averages, dict of predictions)."""
    ad = adaptor_list_forward(sd, spec, phrase_ids, phrase_valid, loss_masking)
    ad, feats, logits = forward_model(sd, spec, ad, camera_images, placeholder_values, training)
    lang_logits, _ = split_outputs(ad, logits)
    _, drv_feats = split_outputs(ad, feats)
    labels = torch.where(loss_masking.bool(), phrase_ids, -1)[:, 1:]
    ce = F.cross_entropy(lang_logits[:, :-1].flatten(0, -2), labels.flatten(), ignore_index=-1,
                         reduction="none").view_as(labels)
    pred = driving_predictions(sd, spec, drv_feats)
    l_route = F.smooth_l1_loss(pred["route"], path, reduction="none").sum(-1)
    l_speed = F.smooth_l1_loss(pred["speed_wps"], waypoints[:, : spec.n_route + 1], reduction="none").sum(-1)
    terms = {
        "language_loss": (ce, labels.ne(-1)),
        "route_loss": (l_route, torch.ones_like(l_route, dtype=torch.long)),
        "speed_wps_loss": (l_speed, torch.ones_like(l_speed, dtype=torch.long)),
    }
    avgs = {k: torch.where(n.sum() > 0, v.sum() / n.sum(), torch.zeros(())) for k, (v, n) in terms.items()}
    loss = torch.stack(list(avgs.values())).sum()
    return loss, avgs, pred


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, beta1: float = 0.9,
               beta2: float = 0.999, eps: float = 1e-8, wd: float = 0.1):
    """torch.optim.AdamW single-tensor update (decoupled decay applied first), as configured at
    driving.py:718-724 (one param group, wd on everything)."""
    p = p * (1 - lr * wd)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


def one_cycle(step: int, total_steps: int, max_lr: float, pct_start: float = 0.05, div_factor: float = 25.0,
              final_div_factor: float = 1e4, base_momentum: float = 0.85, max_momentum: float = 0.95):
    """torch.optim.lr_scheduler.OneCycleLR (cos anneal, two phases) as used at driving.py:729-731 -> (lr, beta1).
    The reference keeps the default ``cycle_momentum=True``: the scheduler overwrites AdamW's beta1 every step,
    annealing it 0.95 -> 0.85 while the lr warms up and back to 0.95 while it decays."""
    initial = max_lr / div_factor
    min_lr = initial / final_div_factor
    up_end = float(pct_start * total_steps) - 1
    down_end = total_steps - 1

    def cos(a, b, pct):
        return b + (a - b) / 2.0 * (math.cos(math.pi * pct) + 1)

    if step <= up_end:
        pct = step / up_end if up_end > 0 else 1.0
        return cos(initial, max_lr, pct), cos(max_momentum, base_momentum, pct)
    pct = (step - up_end) / (down_end - up_end)
    return cos(max_lr, min_lr, pct), cos(base_momentum, max_momentum, pct)


def one_cycle_lr(step: int, total_steps: int, max_lr: float, pct_start: float = 0.05) -> float:
    return one_cycle(step, total_steps, max_lr, pct_start)[0]
