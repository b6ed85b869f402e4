"""End-to-end rounding budget: fp32 emulation of ViT -> projector -> Qwen2 -> heads on the GPU with selectable bf16 roundings,
error of route / speed_wps (max-abs over the tensor's largest magnitude) over several samples.  Diagnostic only."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as O  # noqa: E402
from simlingo_b200.spec import INTERNVL2_1B as SPEC, LLM_PREFIX, MLP1_PREFIX, VIT_PREFIX, init_state_dict  # noqa: E402
from tests.helpers import make_case_inputs  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = "cuda"
sd = init_state_dict(SPEC, seed=0)
W = {k: v.to(dev) for k, v in sd.items()}
bf = lambda t: t.to(torch.bfloat16).float()
NS = int(os.environ.get("NS", "6"))
cases = [make_case_inputs(SPEC, 1, seed=100 + i) for i in range(NS)]


def vit(px, A):
    r = lambda n, t: bf(t) if n in A else t
    e = VIT_PREFIX + "embeddings."
    x = F.conv2d(px, W[e + "patch_embedding.weight"], W[e + "patch_embedding.bias"], stride=14).flatten(2).transpose(1, 2)
    x = torch.cat([W[e + "class_embedding"].expand(x.size(0), 1, -1), x], 1) + W[e + "position_embedding"]
    x = r("vresid", x)
    B, N, C = x.shape
    for i in range(SPEC.vit_layers):
        p = f"{VIT_PREFIX}encoder.layers.{i}."
        h = r("vh", F.layer_norm(x, (C,), W[p + "norm1.weight"], W[p + "norm1.bias"], SPEC.vit_eps))
        qkv = r("vqkv", F.linear(h, W[p + "attn.qkv.weight"], W[p + "attn.qkv.bias"])).reshape(B, N, 3, 16, 64).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        s = (q * 0.125) @ k.transpose(-2, -1)
        if "vp" in A:
            m = s.max(-1, keepdim=True).values
            ee = (s - m).exp()
            o = (bf(ee) @ v) / ee.sum(-1, keepdim=True)
        else:
            o = s.softmax(-1) @ v
        o = r("vatt", o.transpose(1, 2).reshape(B, N, C))
        x = r("vresid", x + F.linear(o, W[p + "attn.proj.weight"], W[p + "attn.proj.bias"]) * W[p + "ls1"])
        h = r("vh", F.layer_norm(x, (C,), W[p + "norm2.weight"], W[p + "norm2.bias"], SPEC.vit_eps))
        h = r("vact", F.gelu(F.linear(h, W[p + "mlp.fc1.weight"], W[p + "mlp.fc1.bias"])))
        x = r("vresid", x + F.linear(h, W[p + "mlp.fc2.weight"], W[p + "mlp.fc2.bias"]) * W[p + "ls2"])
    x = O.pixel_shuffle_closed_form(x[:, 1:], SPEC.grid)
    x = r("vh", F.layer_norm(x, (SPEC.proj_in,), W[MLP1_PREFIX + "0.weight"], W[MLP1_PREFIX + "0.bias"], SPEC.proj_eps))
    x = r("vact", F.gelu(F.linear(x, W[MLP1_PREFIX + "1.weight"], W[MLP1_PREFIX + "1.bias"])))
    return r("vout", F.linear(x, W[MLP1_PREFIX + "3.weight"], W[MLP1_PREFIX + "3.bias"]))


def llm(x0, A, fold):
    r = lambda n, t: bf(t) if n in A else t
    B, L, D = x0.shape
    H, KV, d = 14, 2, 64
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
    norm = lambda x, w: w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + SPEC.rms_eps))
    x = r("resid", x0)
    for i in range(SPEC.llm_layers):
        p = f"{LLM_PREFIX}model.layers.{i}."
        h = r("h", norm(x, W[p + "input_layernorm.weight"]))
        q = r("qkv", lin(p + "self_attn.q_proj.", h)).view(B, L, H, d).transpose(1, 2)
        k = r("qkv", lin(p + "self_attn.k_proj.", h)).view(B, L, KV, d).transpose(1, 2)
        v = r("qkv", lin(p + "self_attn.v_proj.", h)).view(B, L, KV, d).transpose(1, 2)
        q, k = r("qkv", q * cos + rot(q) * sin), r("qkv", k * cos + rot(k) * sin)
        k, v = k.repeat_interleave(H // KV, 1), v.repeat_interleave(H // KV, 1)
        s = (q @ k.transpose(2, 3)) * d ** -0.5 + mask
        if "p" in A:
            m = s.max(-1, keepdim=True).values
            e = (s - m).exp()
            o = (bf(e) @ v) / e.sum(-1, keepdim=True)
        else:
            o = s.softmax(-1) @ v
        o = r("att", o.transpose(1, 2).reshape(B, L, H * d))
        x = r("resid", x + lin(p + "self_attn.o_proj.", o))
        h = r("h", norm(x, W[p + "post_attention_layernorm.weight"]))
        a = r("act", F.silu(lin(p + "mlp.gate_proj.", h)) * lin(p + "mlp.up_proj.", h))
        x = r("resid", x + lin(p + "mlp.down_proj.", a))
    return r("feat", norm(x, W[LLM_PREFIX + "model.norm.weight"]))


def run(case, A, fold):
    with torch.no_grad():
        ad = O.adaptor_list_forward(sd, SPEC, case["ids"], case["valid"], case["loss_masking"])
        emb = ad["language_inputs"].to(dev).clone()
        ids = case["ids"].to(dev)
        # <TARGET_POINT> rows (wp_encoder) and image rows
        coords = torch.as_tensor(case["placeholders"][0][SPEC.target_point_id], dtype=torch.float32, device=dev)
        wp = coords
        for j, act in ((0, True), (2, True), (4, False)):
            wp = F.linear(wp, W[f"wp_encoder.mlp.{j}.weight"], W[f"wp_encoder.mlp.{j}.bias"])
            wp = F.relu(wp) if act else wp
        start = int((ids[0] == SPEC.target_point_id).nonzero()[0])
        emb[0, start:start + 2] = wp
        v = vit(case["frames"].reshape(2, 3, 448, 448).to(dev), A).reshape(-1, SPEC.llm_hidden)
        emb[0, ids[0] == SPEC.img_context_id] = v
        if "embed" in A:
            emb = bf(emb)
        x0 = torch.cat([emb, W["adaptors.driving.query_embeds_wps"], W["adaptors.driving.query_embeds_speed"]], 1)
        f = llm(x0, A, fold)[:, -30:]
        pred = O.driving_predictions(W, SPEC, f)
    return pred["route"], pred["speed_wps"]


rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
refs = [run(c, set(), False) for c in cases]
with torch.no_grad():
    c0 = cases[0]
    sp_cpu, rt_cpu, _ = None, None, None
    ad = O.adaptor_list_forward(sd, SPEC, c0["ids"], c0["valid"], c0["loss_masking"])
    ad, feats, _ = O.forward_model(sd, SPEC, ad, c0["frames"], c0["placeholders"], logits=False)
    pr = O.driving_predictions(sd, SPEC, O.split_outputs(ad, feats)[1])
    print(f"emulation vs CPU oracle: route {rel(refs[0][0].cpu(), pr['route']):.2e} speed {rel(refs[0][1].cpu(), pr['speed_wps']):.2e}")
VIT_ALL = {"vresid", "vh", "vqkv", "vp", "vatt", "vact", "vout"}
LLM_ALL = {"h", "qkv", "p", "att", "act", "feat", "embed"}
variants = [
    ("engine now: ViT all bf16, LLM fp32 resid, folded", VIT_ALL | LLM_ALL, True),
    ("+ exact LoRA", VIT_ALL | LLM_ALL | {"lora_t"}, False),
    ("+ ViT fp32 resid (folded)", (VIT_ALL - {"vresid"}) | LLM_ALL, True),
    ("+ ViT fp32 resid + exact LoRA", (VIT_ALL - {"vresid"}) | LLM_ALL | {"lora_t"}, False),
    ("r01 engine: everything bf16, folded", VIT_ALL | LLM_ALL | {"resid"}, True),
    ("ViT roundings only", VIT_ALL, False),
    ("ViT vresid only", {"vresid"}, False),
    ("LLM non-resid roundings only, exact LoRA", LLM_ALL | {"lora_t"}, False),
    ("fold only", set(), True),
]
for name, A, fold in variants:
    errs = [(rel(r_[0], ref[0]), rel(r_[1], ref[1])) for r_, ref in ((run(c, A, fold), ref) for c, ref in zip(cases, refs))]
    rt = [e[0] for e in errs]
    sp = [e[1] for e in errs]
    print(f"{name:52s}: route max {max(rt):.4f} mean {sum(rt)/len(rt):.4f} | speed max {max(sp):.4f} mean {sum(sp)/len(sp):.4f}")
