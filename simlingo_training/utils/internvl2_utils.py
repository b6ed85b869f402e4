"""Drop-in for the reference's ``simlingo_training/utils/internvl2_utils.py``.

Two halves, as in the reference:

* image pre-processing directly in front of the model: ``preprocess_image_batch`` etc. run on the GPU
  (``simlingo_b200.preprocess``, bit-exact against the reference's Pillow path);
* the tokenisation wire format of training / evaluation batches: chat template -> token ids ->
  ``phrase_valid`` / ``loss_masking`` (reference :29-175).  Pure host-side integer work on KB-sized tensors; the tokenizer
  is whatever object the caller hands in (duck-typed like HF's: ``tokenizer(list[str], padding=True, return_tensors="pt",
  add_special_tokens=False)["input_ids"]``, ``tokenizer(str)["input_ids"]``, ``pad_token_id``).

The upstream pieces this file leans on cannot be fetched offline: ``conversation.py`` of the HF-Hub repo
``OpenGVLab/InternVL2-1B`` (revision unpinned by the reference, :111-120).  Its ``internlm2-chat`` template is restated in
``CONV_TEMPLATES`` (system template, roles, separator, MPT separator style) - parity of that restatement is UNPINNED; the
token-level logic (``get_chat_tokens`` / ``get_assistant_loss_mask``) is pinned against the reference's own functions
(``tests/golden/make_golden_collate.py``)."""
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from simlingo_b200.preprocess import preprocess_frames, preprocess_image_batch, tile_grid  # noqa: F401
from simlingo_b200.spec import INTERNVL2_1B

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)

IMG_START_TOKEN, IMG_END_TOKEN, IMG_CONTEXT_TOKEN, IMG_TOKEN = "<img>", "</img>", "<IMG_CONTEXT>", "<image>"

# UPSTREAM conversation.py (MPT separator style: system + sep, then role + message + sep per turn, a bare role for an open turn)
CONV_TEMPLATES = {
    "internlm2-chat": dict(
        system_template="<|im_start|>system\n{system_message}",
        system_message="你是由上海人工智能实验室联合商汤科技开发的书生多模态大模型，英文名叫InternVL, 是一个有用无害的人工智能助手。",
        roles=("<|im_start|>user\n", "<|im_start|>assistant\n"),
        sep="<|im_end|>",
    ),
}


def get_num_image_tokens_per_patch(encoder_variant: str) -> int:
    """(image_size // patch_size)^2 * downsample_ratio^2 = 256 for InternVL2-1B; the reference reads it from the Hub
    config (:21-27), here it comes from the static architecture description."""
    from simlingo_b200.modules import spec_for_variant
    try:
        return spec_for_variant(encoder_variant).tokens_per_tile
    except Exception:
        return INTERNVL2_1B.tokens_per_tile


def get_assistant_loss_mask(user_starts: Sequence[Sequence[int]], assistant_starts: Sequence[Sequence[int]],
                            prompt_tokenized_ids: torch.Tensor) -> torch.Tensor:
    """bool [B, L], True where the loss is taken: from every assistant-turn start (the role tokens included) up to the
    token before the next user turn, the last turn running to the end of the row (reference :29-47)."""
    n_rows, seq_len = prompt_tokenized_ids.shape
    mask = torch.zeros((n_rows, seq_len), dtype=torch.bool)
    for row, (users, assistants) in enumerate(zip(user_starts, assistant_starts)):
        assert users[0] < assistants[0], "First user start should be before first assistant start"
        assert len(users) == len(assistants), "Number of user and assistant starts should be the same"
        stops = [u - 1 for u in users[1:]] + [seq_len - 1]
        for begin, stop in zip(assistants, stops):
            mask[row, begin:stop + 1] = True
    return mask


def _find_all(ids: torch.Tensor, pattern: torch.Tensor) -> List[List[int]]:
    """start indices of every occurrence of ``pattern`` in each row of ``ids``"""
    n = pattern.numel()
    if ids.shape[1] < n:
        return [[] for _ in range(ids.shape[0])]
    hit = (ids.unfold(1, n, 1) == pattern).all(dim=2)
    return [torch.nonzero(h, as_tuple=True)[0].tolist() for h in hit]


def get_chat_tokens(tokenizer, prompts: List[str], user_start_token_str: str, assistant_start_token_str: str) -> Dict:
    """Tokenises a batch of chat prompts with padding and derives the loss mask from the positions of the user / assistant
    role markers (reference :50-92).  Multi-round rows are handled per row (the reference sizes its per-row lists by the
    number of matches and therefore only works for the single-round conversations SimLingo uses; identical there)."""
    ids = tokenizer(prompts, padding=True, return_tensors="pt", add_special_tokens=False)["input_ids"]
    valid = ids != tokenizer.pad_token_id
    user_pat = torch.tensor(tokenizer(user_start_token_str)["input_ids"])
    assistant_pat = torch.tensor(tokenizer(assistant_start_token_str)["input_ids"])
    loss_mask = get_assistant_loss_mask(_find_all(ids, user_pat), _find_all(ids, assistant_pat), ids)
    return {"phrase_ids": ids, "phrase_valid": valid, "phrase_mask": valid, "language_string": prompts, "loss_masking": loss_mask}


def _mpt_prompt(tpl: dict, messages: List[Tuple[str, Optional[str]]]) -> str:
    out = tpl["system_template"].format(system_message=tpl["system_message"]) + tpl["sep"]
    for role, text in messages:
        out += role + text + tpl["sep"] if text else role
    return out


def get_custom_chat_template(conversations: List[List[Dict]], tokenizer, encoder_variant: str, num_image_tokens_total: int,
                             cache_root_dir: str = "pretrained") -> Tuple[Dict, Dict]:
    """(conversation batch, question-only batch) in the ``internlm2-chat`` format with the system prompt removed and
    ``<image>`` replaced by ``<img>`` + num_image_tokens_total x ``<IMG_CONTEXT>`` + ``</img>`` (reference :95-175).
    ``cache_root_dir`` is accepted for signature compatibility (the reference downloads conversation.py there)."""
    tpl = CONV_TEMPLATES["internlm2-chat"]
    user_role, assistant_role = tpl["roles"]
    system_prompt = tpl["system_template"].format(system_message=tpl["system_message"]) + tpl["sep"]
    image_span = IMG_START_TOKEN + IMG_CONTEXT_TOKEN * num_image_tokens_total + IMG_END_TOKEN
    full, question = [], []
    for conv in conversations:
        assert len(conv) == 2, "For question and answer templates only two turn conversation (user + assistant) is supported. During training is should work but is not checked!!"
        assert conv[0]["role"] == "user", "First turn should be user as this should be the question."
        turns = []
        for i, part in enumerate(conv):
            text = part["content"][0]["text"]
            if part["role"] == "assistant":
                turns.append((assistant_role, text))
            elif part["role"] == "user":
                if i == 0 and IMG_TOKEN not in text:
                    text = f"{IMG_TOKEN}\n" + text
                turns.append((user_role, text))
            else:
                raise ValueError(f"Role {part['role']} not supported")
        first = conv[0]["content"][0]["text"]
        if IMG_TOKEN not in first:
            first = f"{IMG_TOKEN}\n" + first
        for store, messages in ((full, turns), (question, [(user_role, first), (assistant_role, None)])):
            prompt = _mpt_prompt(tpl, messages).replace(system_prompt, "")   # drop the system prompt to save tokens
            store.append(prompt.replace(IMG_TOKEN, image_span, 1))
    return (get_chat_tokens(tokenizer, full, user_role, assistant_role),
            get_chat_tokens(tokenizer, question, user_role, assistant_role))
