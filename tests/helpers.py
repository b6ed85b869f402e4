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


class StubChatTokenizer:
    """Deterministic stand-in for the InternVL2 tokenizer's batch interface (no tokenizer files offline): splits on the
    chat / image special tokens, newlines and whitespace, maps pieces to ids by a stable hash; pads to the longest row on
    ``padding_side``.  Enough to exercise the token-level wire format (role-marker search, loss mask, padding)."""
    SPECIALS = {"<|im_start|>": 151644, "<|im_end|>": 151645, "<img>": 151646, "</img>": 151647, "<IMG_CONTEXT>": 151648,
                "<TARGET_POINT>": 151662, "\n": 198}

    def __init__(self, padding_side="left"):
        import re
        self.padding_side = padding_side
        self.pad_token_id = 151643
        self.eos_token_id = 151645
        self._split = re.compile("(" + "|".join(re.escape(k) for k in self.SPECIALS) + r"|\s+)")

    def convert_tokens_to_ids(self, tok):
        return self.SPECIALS[tok]

    def _encode(self, text):
        import zlib
        out = []
        for piece in self._split.split(text):
            if not piece or (piece.isspace() and piece != "\n"):
                continue
            out.append(self.SPECIALS[piece] if piece in self.SPECIALS else zlib.crc32(piece.encode()) % 150000)
        return out

    def __call__(self, text, padding=False, return_tensors=None, add_special_tokens=True):
        if isinstance(text, str):
            return {"input_ids": self._encode(text)}
        rows = [self._encode(t) for t in text]
        width = max(len(r) for r in rows)
        pad = lambda r: ([self.pad_token_id] * (width - len(r)) + r) if self.padding_side == "left" else (r + [self.pad_token_id] * (width - len(r)))
        ids = torch.tensor([pad(r) for r in rows], dtype=torch.long)
        return {"input_ids": ids, "attention_mask": (ids != self.pad_token_id).long()}


def make_dataset_outputs(n, seed, h=359, w=1024):
    """``DatasetOutput`` samples shaped like the reference's driving dataset items (dataset_base.py / dataset_driving.py):
    one user + one assistant turn, a <TARGET_POINT> placeholder pair, 11 waypoints, 20 route points."""
    import numpy as np
    from simlingo_b200.spec import synth_camera
    from simlingo_training.utils.custom_types import DatasetOutput
    rs = np.random.RandomState(seed)
    out = []
    for i in range(n):
        q = "Current speed: %.1f m/s. Target waypoint: <TARGET_POINT><TARGET_POINT>. " % rs.uniform(0, 10) + " ".join(["Predict"] + ["the"] * int(rs.randint(0, 4)) + ["waypoints."])
        a = "Waypoints: " + " ".join("w%d" % int(rs.randint(0, 50)) for _ in range(int(rs.randint(2, 7))))
        conv = [{"role": "user", "content": [{"type": "text", "text": q}]}, {"role": "assistant", "content": [{"type": "text", "text": a}]}]
        out.append(DatasetOutput(
            conversation=conv, answer=[conv[1]], image_ff=synth_camera(h, w, seed * 100 + i)[None], image_ff_org_size=np.array([512, 1024]),
            waypoints=rs.rand(11, 2).astype(np.float32).cumsum(0), waypoints_1d=rs.rand(11, 2).astype(np.float32), path=rs.rand(20, 2).astype(np.float32).cumsum(0),
            target_points=rs.randn(2).astype(np.float32) * 10, speed=float(rs.uniform(0, 10)),
            placeholder_values={"<TARGET_POINT>": rs.randn(2, 2).astype(np.float32) * 10}, measurement_path=f"route_{seed}/measurements/{i:04d}.json.gz",
            dataset="driving", qa_templates=("q?", "a."), eval_infos={"k": i}))
    return out
