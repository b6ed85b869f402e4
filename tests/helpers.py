"""Shared builders for tests: synthetic DrivingInput tuples and a duck-typed tokenizer (no tokenizer files offline)."""
import torch

from simlingo_b200.spec import LMHEAD_SHIFT, synth_frames, synth_labels, synth_placeholders, synth_prompt_ids


class StubTokenizer:
    """The tokenizer facts the model code touches (SURVEY 8b "duck-typed collaborators")."""

    def __init__(self, spec):
        self.spec = spec
        self.eos_token_id = spec.eos_id
        self.additional_special_tokens_ids = list(range(spec.first_added_id, spec.first_added_id + 8))
        self.added_tokens_encoder = {"<|im_end|>": spec.eos_id}

    def convert_tokens_to_ids(self, tok):
        return {"<IMG_CONTEXT>": self.spec.img_context_id}[tok]

    def batch_decode(self, tokens, skip_special_tokens=True):
        return [" ".join(str(int(t)) for t in row) for row in tokens]


def make_case_inputs(spec, B, seed, G_list=None, answer_len=0, pad_rows=()):
    """Same construction as tests/golden/make_golden.py::make_input (kept in sync by the golden tests)."""
    ids = synth_prompt_ids(spec, B, seed, answer_len=answer_len)
    if G_list is not None:
        for b, G in enumerate(G_list):
            ids[b, -1] = (spec.eos_id - G * LMHEAD_SHIFT) % spec.vocab
    valid = torch.ones_like(ids, dtype=torch.bool)
    for b, n in pad_rows:
        valid[b, :n] = False
    lm = torch.zeros_like(valid)
    if answer_len:
        lm[:, -answer_len:] = True
    return dict(ids=ids, valid=valid, loss_masking=lm, frames=synth_frames(spec, B, seed),
                placeholders=synth_placeholders(spec, B, seed), labels=synth_labels(spec, B, seed))


def to_driving_input(case, device=None, dtype=None):
    from simlingo_training.utils.custom_types import DrivingInput, LanguageLabel
    mv = (lambda t: t.to(device)) if device is not None else (lambda t: t)
    fr = case["frames"]
    if dtype is not None:
        fr = fr.to(dtype)
    label = LanguageLabel(mv(case["ids"]), mv(case["valid"]), mv(case["valid"].clone()), case["placeholders"],
                          [""] * case["ids"].shape[0], mv(case["loss_masking"]))
    z = mv(torch.zeros(case["ids"].shape[0], 1))
    return DrivingInput(mv(fr), z, z, z, z, z, label, label)


def build_drop_in_model(spec, name, seed=0, device="cuda", freeze=False, lora_dropout=0.1):
    """DrivingModel (drop-in mirror) on ``spec`` with the deterministic synthetic weights, bf16 on ``device``."""
    from simlingo_b200.modules import register_variant
    from simlingo_b200.spec import init_state_dict
    from simlingo_training.models.driving import DrivingModel
    register_variant(name, spec)
    cfg = dict(vision_model=dict(_target_="simlingo_training.models.encoder.vlm.VLMEncoderModel", variant=name, embed_dim=512, freeze=freeze),
               language_model=dict(_target_="simlingo_training.models.language_model.llm.LLM", variant=name, lora=True, lora_alpha=64,
                                   lora_r=32, lora_dropout=lora_dropout),
               lr=3e-5, weight_decay=0.1, betas=(0.9, 0.999), pct_start=0.05, speed_wps_mode="2d", predict_route_as_wps=True)
    torch.set_default_dtype(torch.bfloat16)
    try:
        m = DrivingModel(cfg_data_module={"use_global_img": False}, processor=StubTokenizer(spec), cache_dir=None, **cfg)
    finally:
        torch.set_default_dtype(torch.float32)
    m.load_state_dict(init_state_dict(spec, seed=seed, with_aliases=True), strict=True)
    return m.to(device)


def to_driving_example(case, device="cuda", dtype=torch.bfloat16):
    from simlingo_training.utils.custom_types import DrivingExample, DrivingLabel
    di = to_driving_input(case, device, dtype)
    wps, path = case["labels"]
    B = case["ids"].shape[0]
    return DrivingExample(di, DrivingLabel(wps.to(device), path.to(device), di.prompt, torch.zeros(1)), ["x"] * B)
