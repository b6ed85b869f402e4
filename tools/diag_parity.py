"""Where does the bf16-vs-fp32 error of the full-depth model come from?  Per-stage max-abs error over the tensor's largest
magnitude (the north_star's measure): ViT layers, projector, decoder layers (fed with the ORACLE's inputs, so each stage
is judged on its own), with the LoRA weights folded (inference engine) and un-merged (training engine, eval mode).
    python tools/diag_parity.py > gpurun_out/diag_parity.log"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as O  # noqa: E402  (diagnostic tool, not product code)
from simlingo_b200.spec import INTERNVL2_1B as SPEC, init_state_dict  # noqa: E402
from tests.helpers import build_drop_in_model, make_case_inputs  # noqa: E402

rel = lambda a, b: float((a.float().cpu() - b).abs().max() / b.abs().max())
rms = lambda a, b: float((a.float().cpu() - b).norm() / b.norm())
sd = init_state_dict(SPEC, seed=0)
model = build_drop_in_model(SPEC, "OpenGVLab/InternVL2-1B").eval()
eng = model._engine()
case = make_case_inputs(SPEC, 1, seed=71)
px = case["frames"].reshape(2, 3, 448, 448)
ref_layers, got_layers = [], []
with torch.no_grad():
    vit_ref = O.extract_feature(sd, SPEC, px, ref_layers)
    eng.vit(px.to("cuda", torch.bfloat16), got_layers)
    vit_got = eng.extract_feature(px.to("cuda", torch.bfloat16))
for i in (0, 5, 11, 17, 23):
    print(f"vit layer {i:2d}: max {rel(got_layers[i].view(2, 1025, 1024), ref_layers[i]):.4f} rms {rms(got_layers[i].view(2, 1025, 1024), ref_layers[i]):.4f}")
print(f"vit_embeds  : max {rel(vit_got.view(2, 256, 896), vit_ref):.4f} rms {rms(vit_got.view(2, 256, 896), vit_ref):.4f}")
with torch.no_grad():
    ad = O.adaptor_list_forward(sd, SPEC, case["ids"], case["valid"], case["loss_masking"])
    ad = O.replace_placeholder_tokens(sd, SPEC, ad, case["frames"], case["placeholders"])
    ref_l = []
    feats_ref, _ = O.llm_forward(sd, SPEC, ad["inputs"], ad["inputs_mask"], None, collect=ref_l)
    inputs = ad["inputs"].to("cuda", torch.bfloat16)     # the oracle's exact inputs, rounded once
    B, Lt, D = inputs.shape
    got_l = []
    x = eng.llm_chunk(inputs.reshape(B * Lt, D).clone(), B, Lt, 0, eng.new_cache(B, Lt), None, collect=got_l)
    feats = eng.final_norm(x).view(B, Lt, D)
for i in (0, 5, 11, 17, 23):
    print(f"llm layer {i:2d} (folded LoRA): max {rel(got_l[i].view(B, Lt, D), ref_l[i]):.4f} rms {rms(got_l[i].view(B, Lt, D), ref_l[i]):.4f}   query rows: "
          f"max {rel(got_l[i].view(B, Lt, D)[:, -30:], ref_l[i][:, -30:]):.4f}")
print(f"features (folded): max {rel(feats, feats_ref):.4f} rms {rms(feats, feats_ref):.4f}; query rows max {rel(feats[:, -30:], feats_ref[:, -30:]):.4f} rms {rms(feats[:, -30:], feats_ref[:, -30:]):.4f}")
# un-merged LoRA: the training engine in eval mode (no dropout), forward only
from simlingo_b200 import training  # noqa: E402
store = model.param_store()
teng = model.__dict__["_slb_train_engine"]
teng.graphs_enabled = False
with torch.no_grad():
    f2, saved = teng.llm_forward(inputs.clone(), None, dropout=False)
print(f"features (un-merged LoRA, training kernels): max {rel(f2, feats_ref):.4f} rms {rms(f2, feats_ref):.4f}; query rows max {rel(f2[:, -30:], feats_ref[:, -30:]):.4f} "
      f"rms {rms(f2[:, -30:], feats_ref[:, -30:]):.4f}")
for i in (0, 5, 11, 17, 23):
    xo = saved["layers"][i + 1]["x"] if i + 1 < SPEC.llm_layers else saved["xf"]
    print(f"llm layer {i:2d} (un-merged)  : max {rel(xo.view(B, Lt, D), ref_l[i]):.4f} rms {rms(xo.view(B, Lt, D), ref_l[i]):.4f}")
