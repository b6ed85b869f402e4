"""Language model wrapper (drop-in for reference ``simlingo_training/models/language_model/llm.py``): Qwen2-0.5B from
InternVL2-1B with PEFT-style LoRA on all linears, teacher-forced ``forward`` and ``greedy_sample``.

``greedy_sample`` keeps the reference's observable behaviour (EOS-prefilled ``sampled_tokens``, the embedding of
every sampled token - EOS included - appended to ``input_embeds``, early exit once all rows hit EOS) but runs on a
KV cache instead of re-forwarding the whole growing sequence for every token (reference :217-235)."""
from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from simlingo_b200 import lib as _lib
from simlingo_b200 import runtime as _rt
from simlingo_b200.modules import PeftModelForCausalLM, Qwen2ForCausalLM, spec_for_variant
from simlingo_b200.spec import LLM_PREFIX


class LLM(nn.Module):
    def __init__(self, **cfg):
        super().__init__()
        for key, value in cfg.items():
            setattr(self, key, value)
        if "internvl" not in self.variant.lower():
            raise ValueError(f"Carefull: Variant {self.variant} not tested.")
        self.spec = spec_for_variant(self.variant)
        lora = bool(getattr(self, "lora", False))
        if lora:
            import dataclasses
            want = dict(lora_r=int(getattr(self, "lora_r", self.spec.lora_r)), lora_alpha=getattr(self, "lora_alpha", self.spec.lora_alpha),
                        lora_dropout=float(getattr(self, "lora_dropout", self.spec.lora_dropout)))
            if any(getattr(self.spec, k) != v for k, v in want.items()):   # every LoRA hyper-parameter of the config is honoured (llm.py:106-118)
                self.spec = dataclasses.replace(self.spec, **want)
        if not lora:
            raise NotImplementedError("simlingo_b200 implements the released configuration (lora=True, all-linear)")
        self.model = Qwen2ForCausalLM(self.spec, lora=True)
        # alias used by LanguageAdaptor; registers the table a second time under `...model.embed_tokens`
        self.model.embed_tokens = self.model.base_model.embed_tokens
        print("Using PEFT model")
        self.model = PeftModelForCausalLM(self.model)
        self.model.print_trainable_parameters()
        self.vocab_size = self.model.config.vocab_size
        self.hidden_size = self.model.config.hidden_size
        self.max_position_embeddings = self.model.config.max_position_embeddings

    def _causal_lm(self):
        return self.model.base_model.model

    def forward(self, embeddings: Tensor, attention_mask: Tensor = None, return_dict: bool = True,
                position_ids: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """-> (features = last hidden state after the final norm, logits) as reference :126-143."""
        out = self.model(inputs_embeds=embeddings, attention_mask=attention_mask, output_hidden_states=True,
                         position_ids=position_ids, return_dict=return_dict)
        return out.hidden_states[-1], out[0]

    def sample_categorical(self, logits: Tensor, temperature: float = 0.0, top_k: Optional[int] = None,
                           top_p: Optional[float] = None, restrict_tokens: Optional[Tuple[int, int]] = None):
        """argmax for temperature <= 0 (the only branch the reference exercises), else top-k / nucleus sampling."""
        if restrict_tokens is not None:
            lo, n = restrict_tokens
            logits[..., :lo] = -float("inf")
            logits[..., lo + n:] = -float("inf")
        if temperature <= 0.0:
            if logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 2:
                return _lib.argmax(logits.contiguous())
            return logits.argmax(dim=-1, keepdim=False)
        if top_k is not None:
            kth = torch.topk(logits, min(top_k, logits.size(-1))).values.select(-1, -1).unsqueeze(-1)
            logits = torch.where(logits < kth, -float("inf"), logits)
        logits = logits / max(temperature, 1e-9)
        if top_p is not None:
            srt, order = torch.sort(logits, descending=True, dim=-1)
            drop = (torch.softmax(srt, dim=-1).cumsum(dim=-1) > top_p).roll(shifts=1, dims=-1)
            drop[..., 0] = False
            logits[drop.gather(-1, order.argsort(-1))] = -float("inf")
        return torch.multinomial(logits.softmax(dim=-1), 1).squeeze(-1)

    @torch.no_grad()
    def greedy_sample(self, input_embeds: Tensor, inputs_mask: Optional[Tensor] = None, max_new_tokens: int = 100,
                      temperature: float = 0.0, top_k: Optional[int] = None, top_p: Optional[float] = None,
                      eos_token_id: Optional[int] = None, cache_offset: int = 0, input_embed_matrix: Optional[Tensor] = None,
                      logit_matrix: Optional[Tensor] = None, restrict_tokens: Optional[Tuple[int, int]] = None,
                      attention_mask=None, position_ids=None) -> Tuple[Tensor, Tensor]:
        """-> (sampled_tokens [B, <=max_new_tokens] int64, input_embeds [B, L+G, H])."""
        if input_embed_matrix is None or logit_matrix is None:
            raise ValueError("No input embeddings / logit matrix available; please provide input_embed_matrix and logit_matrix.")
        if position_ids is not None:
            raise NotImplementedError("position_ids must be None (reference always uses arange)")
        eng = _rt.engine_for(self._causal_lm(), LLM_PREFIX, self.spec)
        B, L, D = input_embeds.shape
        dev = input_embeds.device
        cache = eng.new_cache(B, L + max_new_tokens)
        kv_valid = None
        if attention_mask is not None and not bool(attention_mask.bool().all()):
            kv_valid = torch.ones((B, cache[0].shape[3]), device=dev, dtype=torch.uint8)
            kv_valid[:, :L] = attention_mask.to(torch.uint8)
        sampled = torch.empty((B, max_new_tokens), device=dev, dtype=torch.long)
        if eos_token_id is not None:
            sampled.fill_(eos_token_id)
        incomplete = torch.ones(B, dtype=torch.bool, device=dev)
        x = eng.llm_chunk(input_embeds.reshape(B * L, D).to(torch.bfloat16).clone(), B, L, 0, cache, kv_valid)
        last = eng.final_norm(x.view(B, L, D)[:, -1].contiguous())
        appended = []
        for i in range(max_new_tokens):
            logits = _lib.gemm(last, logit_matrix.detach(), out_fp32=True)
            nxt = self.sample_categorical(logits, temperature=temperature, top_k=top_k, top_p=top_p, restrict_tokens=restrict_tokens)
            e = _lib.gather_rows(input_embed_matrix.detach(), nxt)
            appended.append(e)
            sampled[incomplete, i] = nxt[incomplete]
            if eos_token_id is not None:
                incomplete = sampled[:, i] != eos_token_id
                if not incomplete.any():
                    sampled = sampled[:, : i + 1]
                    break
            if i + 1 < max_new_tokens:
                last = eng.final_norm(eng.llm_chunk(e.clone(), B, 1, L + i, cache, kv_valid))
        out_embeds = torch.cat([input_embeds] + [a.view(B, 1, D).to(input_embeds.dtype) for a in appended], dim=1)
        return sampled, out_embeds
