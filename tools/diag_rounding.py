"""Which bf16 roundings of the Qwen2 stack cost how much accuracy?  fp32 emulation of the decoder on the GPU (torch ops, TF32
off) with bf16 rounding switched on for selected intermediates, against the same emulation without any rounding.
Diagnostic only.   python tools/diag_rounding.py > gpurun_out/diag_rounding.log"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as O  # noqa: E402
from simlingo_b200.spec import INTERNVL2_1B as SPEC, LLM_PREFIX, init_state_dict  # noqa: E402
from tests.helpers import make_case_inputs  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = "cuda"
sd = init_state_dict(SPEC, seed=0)
case = make_case_inputs(SPEC, 1, seed=71)
with torch.no_grad():
    ad = O.adaptor_list_forward(sd, SPEC, case["ids"], case["valid"], case["loss_masking"])
    ad = O.replace_placeholder_tokens(sd, SPEC, ad, case["frames"], case["placeholders"])
    feats_cpu, _ = O.llm_forward(sd, SPEC, ad["inputs"], ad["inputs_mask"], None)
x0 = ad["inputs"].to(dev)
W = {k: v.to(dev) for k, v in sd.items() if k.startswith(LLM_PREFIX)}
bf = lambda t: t.to(torch.bfloat16).float()


def forward(active, fold=False):
    r = lambda name, t: bf(t) if name in active else t
    B, L, D = x0.shape
    H, KV, d = SPEC.llm_heads, SPEC.llm_kv_heads, SPEC.head_dim
    pos = torch.arange(L, device=dev)[None]
    inv = 1.0 / (SPEC.rope_theta ** (torch.arange(0, d, 2, dtype=torch.float32, device=dev) / d))
    fr = pos.float()[..., None] * inv
    emb = torch.cat([fr, fr], -1)
    cos, sin = emb.cos()[:, None], emb.sin()[:, None]
    mask = torch.full((L, L), float("-inf"), device=dev).triu(1)
    rot = lambda t: torch.cat([-t[..., d // 2:], t[..., :d // 2]], -1)

    def lin(p, x):
        w, b = W[p + "base_layer.weight"], W.get(p + "base_layer.bias")
        a, bb = W[p + "lora_A.default.weight"], W[p + "lora_B.default.weight"]
        if fold:
            return F.linear(x, bf(w + 2.0 * bb @ a), b)
        return F.linear(x, w, b) + 2.0 * F.linear(r("lora_t", F.linear(x, a)), bb)

    def norm(x, w):
        return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + SPEC.rms_eps))
    x = r("resid", x0)
    for i in range(SPEC.llm_layers):
        p = f"{LLM_PREFIX}model.layers.{i}."
        h = r("h", norm(x, W[p + "input_layernorm.weight"]))
        q = r("qkv", lin(p + "self_attn.q_proj.", h)).view(B, L, H, d).transpose(1, 2)
        k = r("qkv", lin(p + "self_attn.k_proj.", h)).view(B, L, KV, d).transpose(1, 2)
        v = r("qkv", lin(p + "self_attn.v_proj.", h)).view(B, L, KV, d).transpose(1, 2)
        q, k = r("rope", q * cos + rot(q) * sin), r("rope", k * cos + rot(k) * sin)
        k, v = k.repeat_interleave(H // KV, 1), v.repeat_interleave(H // KV, 1)
        s = (q @ k.transpose(2, 3)) * d ** -0.5 + mask
        pr = s.softmax(-1)
        if "p" in active:   # flash style: un-normalised exp rounded to bf16, fp32 row sum
            m = s.max(-1, keepdim=True).values
            e = (s - m).exp()
            o = (bf(e) @ v) / e.sum(-1, keepdim=True)
        else:
            o = pr @ v
        o = r("att", o.transpose(1, 2).reshape(B, L, H * d))
        x = r("resid", x + lin(p + "self_attn.o_proj.", o))
        h = r("h", norm(x, W[p + "post_attention_layernorm.weight"]))
        a = r("act", F.silu(lin(p + "mlp.gate_proj.", h)) * lin(p + "mlp.up_proj.", h))
        x = r("resid", x + lin(p + "mlp.down_proj.", a))
    return norm(x, W[LLM_PREFIX + "model.norm.weight"])


rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
rms = lambda a, b: float((a - b).norm() / b.norm())
with torch.no_grad():
    ref = forward(set())
    print(f"emulation vs CPU oracle: max {rel(ref.cpu(), feats_cpu):.2e}")
    q = slice(-30, None)
    for name, active, fold in [("fold only", set(), True), ("resid only", {"resid"}, False), ("h only", {"h"}, False), ("qkv+rope only", {"qkv", "rope"}, False),
                               ("p only", {"p"}, False), ("att only", {"att"}, False), ("act only", {"act"}, False), ("lora_t only", {"lora_t"}, False),
                               ("all but resid, exact LoRA", {"h", "qkv", "rope", "p", "att", "act", "lora_t"}, False),
                               ("all, exact LoRA", {"resid", "h", "qkv", "rope", "p", "att", "act", "lora_t"}, False),
                               ("all but resid, folded", {"h", "qkv", "rope", "p", "att", "act"}, True),
                               ("all, folded (= the engine)", {"resid", "h", "qkv", "rope", "p", "att", "act"}, True)]:
        f = forward(active, fold)
        print(f"{name:34s}: all rows max {rel(f, ref):.4f} rms {rms(f, ref):.4f} | query rows max {rel(f[:, q], ref[:, q]):.4f} rms {rms(f[:, q], ref[:, q]):.4f}")
