"""Batch assembly / tokenisation wire format against the golden fixture produced by the REFERENCE's own
``dl_collate_fn`` / ``get_chat_tokens`` / ``get_custom_chat_template`` (tests/golden/make_golden_collate.py).

CPU: token ids, validity, loss mask, placeholder dicts, labels, run ids and calibration matrices are integer / exact float
work - bit-exact.  GPU: the raw uint8 frames our ``Collator`` ships become, on the device, exactly the reference's Pillow
tiles (SHA-256 over the uint8 tiles the normalisation was applied to), and a collated batch runs through ``DrivingModel``."""
import hashlib
import os

import numpy as np
import pytest
import torch

from tests.helpers import StubChatTokenizer, make_dataset_outputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "collate.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


def test_chat_tokens_and_loss_mask_match_reference(golden):
    from simlingo_training.utils.internvl2_utils import CONV_TEMPLATES, get_chat_tokens
    roles = CONV_TEMPLATES["internlm2-chat"]["roles"]
    assert len(golden["chat"]) == 2
    for case in golden["chat"]:
        got = get_chat_tokens(StubChatTokenizer(case["side"]), case["prompts"], *roles)
        assert torch.equal(got["phrase_ids"], case["phrase_ids"]) and torch.equal(got["phrase_valid"], case["phrase_valid"])
        assert torch.equal(got["loss_masking"], case["loss_masking"]) and torch.equal(got["phrase_mask"], case["phrase_valid"])
        assert got["loss_masking"].dtype == torch.bool and got["language_string"] == case["prompts"]
        # the mask starts AT the assistant role marker and runs to the end of the row (right padding included, as the reference)
        n_role = len(StubChatTokenizer()("<|im_start|>assistant\n")["input_ids"])
        n_pad = int((~got["phrase_valid"][2]).sum()) if case["side"] == "right" else 0
        assert int(got["loss_masking"][2].sum()) == n_role + n_pad


def test_loss_mask_multi_round_and_asserts():
    from simlingo_training.utils.internvl2_utils import get_assistant_loss_mask
    ids = torch.zeros((2, 12), dtype=torch.long)
    m = get_assistant_loss_mask([[0, 6], [1]], [[3, 9], [4]], ids)
    assert m[0].tolist() == [False] * 3 + [True] * 3 + [False] * 3 + [True] * 3
    assert m[1].tolist() == [False] * 4 + [True] * 8
    with pytest.raises(AssertionError):
        get_assistant_loss_mask([[5]], [[2]], ids[:1])
    with pytest.raises(AssertionError):
        get_assistant_loss_mask([[0, 4]], [[2]], ids[:1])


@pytest.mark.parametrize("idx", [0, 1])
def test_collate_matches_reference_run(golden, idx):
    from simlingo_training.dataloader.datamodule import Collator
    case = golden["collate"][idx]
    tok = StubChatTokenizer(case["side"])
    ex = Collator(tok, "OpenGVLab/InternVL2-1B", use_global_img=False, predict=case["predict"])(make_dataset_outputs(case["n"], case["seed"]))
    di, dl = ex.driving_input, ex.driving_label
    for name, lab in (("prompt", di.prompt), ("prompt_inference", di.prompt_inference)):
        ref = case[name]
        for k in ("phrase_ids", "phrase_valid", "phrase_mask", "loss_masking"):
            assert torch.equal(getattr(lab, k), ref[k]), (name, k)
        assert lab.language_string == ref["language_string"]
        assert [sorted(d) for d in lab.placeholder_values] == [sorted(d) for d in ref["placeholder_values"]]
        for got, want in zip(lab.placeholder_values, ref["placeholder_values"]):
            assert all(np.array_equal(np.asarray(got[k]), want[k]) for k in want)
    for k in ("image_sizes", "camera_intrinsics", "camera_extrinsics", "vehicle_speed", "target_point"):
        assert torch.equal(getattr(di, k), case[k]), k
    assert torch.equal(dl.waypoints, case["waypoints"]) and torch.equal(dl.path, case["path"]) and torch.equal(dl.image_ff_org, case["image_ff_org"])
    assert dl.answer.language_string == case["answer_strings"] and dl.eval_infos == case["eval_infos"] and ex.qa_templates == case["qa_templates"]
    assert torch.equal(ex.run_id, case["run_id"]) and ex.run_id.dtype == torch.uint8
    # wire-format difference (documented): raw uint8 frames instead of float32 tiles
    assert di.camera_images.dtype == torch.uint8 and tuple(di.camera_images.shape) == (case["n"], 1, 3, 359, 1024)
    from simlingo_training.models.driving import decode_uint8
    assert decode_uint8(ex.run_id)[0].startswith("route_")


def test_encode_uint8_round_trip():
    from simlingo_training.dataloader.datamodule import encode_uint8
    from simlingo_training.models.driving import decode_uint8
    s = ["a/b/c.json.gz", "", "x" * 40]
    enc = encode_uint8(s, 64)
    assert enc.shape == (3, 64) and enc.dtype == torch.uint8 and decode_uint8(enc) == s
    with pytest.raises(AssertionError):
        encode_uint8(["y" * 65], 64)


@pytest.mark.gpu
def test_collated_frames_become_the_reference_tiles_on_the_gpu(golden):
    from simlingo_b200.preprocess import preprocess_frames
    from simlingo_training.dataloader.datamodule import Collator
    case = golden["collate"][0]
    ex = Collator(StubChatTokenizer(case["side"]))(make_dataset_outputs(case["n"], case["seed"]))
    cam = ex.driving_input.camera_images
    tiles = preprocess_frames(cam.view(-1, 3, 359, 1024).cuda()).view(case["n"], 1, 2, 3, 448, 448).float().cpu()
    mean = torch.tensor((0.485, 0.456, 0.406)).view(1, 1, 1, 3, 1, 1)
    std = torch.tensor((0.229, 0.224, 0.225)).view(1, 1, 1, 3, 1, 1)
    ref_u8 = None  # the reference's tiles are k/255 normalised in fp32; ours are the same values rounded once to bf16
    u8 = torch.round((tiles * std + mean) * 255.0).clamp(0, 255).to(torch.uint8)
    assert tuple(tiles.shape) == case["camera_shape"]
    assert hashlib.sha256(np.ascontiguousarray(u8.numpy()).tobytes()).hexdigest() == case["camera_sha256"]


@pytest.mark.gpu
def test_collated_batch_runs_through_the_model():
    """DrivingModel accepts the collated wire format directly: uint8 frames are tiled / normalised on the device, the prompt
    with 512 <IMG_CONTEXT> tokens + <TARGET_POINT> placeholders goes through forward_loss and forward."""
    from simlingo_b200.spec import tiny_spec
    from simlingo_training.dataloader.datamodule import Collator
    from tests.helpers import build_drop_in_model
    spec = tiny_spec(2, 2, 4096)

    class Tok(StubChatTokenizer):   # ids inside the tiny vocabulary, special ids where the tiny spec expects them
        def __init__(self):
            super().__init__("left")
            self.SPECIALS = dict(StubChatTokenizer.SPECIALS)
            self.SPECIALS.update({"<|im_end|>": spec.eos_id, "<img>": spec.img_start_id, "</img>": spec.img_end_id, "<IMG_CONTEXT>": spec.img_context_id,
                                  "<TARGET_POINT>": spec.target_point_id, "<|im_start|>": spec.eos_id - 1, "\n": 198})
            self.pad_token_id = spec.eos_id - 2
            self.eos_token_id = spec.eos_id
            self.additional_special_tokens_ids = list(range(spec.first_added_id, spec.first_added_id + 8))
            self.added_tokens_encoder = {"<|im_end|>": spec.eos_id}

        def _encode(self, text):
            special = set(self.SPECIALS.values())
            return [t if t in special else t % 3000 for t in super()._encode(text)]

        def batch_decode(self, tokens, skip_special_tokens=True):
            return [" ".join(str(int(t)) for t in row) for row in tokens]

    tok = Tok()
    ex = Collator(tok)(make_dataset_outputs(2, 3))
    to = lambda t: t.cuda() if torch.is_tensor(t) else t
    mv = lambda nt: type(nt)(*[mv(v) if hasattr(v, "_fields") else to(v) for v in nt])
    ex = mv(ex)
    model = build_drop_in_model(spec, "internvl2-tiny-collate").eval()
    model.processor = tok
    model.tokenizer = tok
    model.vision_model.image_encoder.processor = tok
    ex = ex._replace(driving_label=ex.driving_label._replace(waypoints=ex.driving_label.waypoints[:, :10]))
    with torch.no_grad():
        out, _ = model.forward_loss(ex)
    assert torch.isfinite(out.loss) and float(out.loss_counts["language_loss"].sum()) == float(ex.driving_input.prompt.loss_masking[:, 1:].sum())
    sp, rt, lang = model(ex)
    assert sp.shape == (2, 10, 2) and rt.shape == (2, 20, 2) and len(lang) == 2
