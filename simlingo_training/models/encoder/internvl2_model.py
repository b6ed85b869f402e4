"""InternVL2 image encoder wrapper + placeholder substitution.

Drop-in for reference ``simlingo_training/models/encoder/internvl2_model.py`` (``LingoInternVLModel``).  The
reference loads ``AutoModel.from_pretrained(variant, trust_remote_code=True)`` from the HF Hub; here the same
module tree is built from the static InternVL2-1B description (``simlingo_b200.spec``) and weights arrive through
``load_state_dict`` (the SimLingo checkpoint holds every tensor)."""
from typing import List, Optional

import torch
from torch import nn

from simlingo_b200 import lib as _lib
from simlingo_b200.modules import InternVLChatModel, spec_for_variant


class LingoInternVLModel(nn.Module):
    def __init__(self, variant, *args, **kwargs):
        super().__init__()
        self.spec = spec_for_variant(variant)
        self.model = InternVLChatModel(self.spec)
        self.num_embeddings = self.spec.vocab
        self.use_global_img = None
        self.processor = None

    def _tokenizer(self):
        proc = self.processor
        return proc.tokenizer if "tokenizer" in proc.__dict__ else proc

    def replace_placeholder_tokens(
        self,
        adaptor_dict=None,
        pixel_values: torch.FloatTensor = None,
        inputs_embeds: Optional[torch.FloatTensor] = None,
        output_attentions: Optional[bool] = None,
        output_hidden_states: Optional[bool] = None,
        return_dict: Optional[bool] = None,
        placeholder_values: Optional[List[dict]] = None,
        wp_encoder: Optional[nn.Module] = None,
    ):
        """Overwrites, in ``adaptor_dict['language_inputs']``:
          * the run starting at the first occurrence of every added special id that has ``placeholder_values``
            with ``wp_encoder(coords)`` (reference :54-91), and
          * every ``<IMG_CONTEXT>`` row with the projected ViT features of the frame's tiles, in order (:102-131);
        then copies each row's language part into the permuted ``adaptor_dict['inputs']`` (:138-142)."""
        self.tokenizer = self._tokenizer()
        self.img_context_token_id = self.tokenizer.convert_tokens_to_ids("<IMG_CONTEXT>")
        if inputs_embeds is not None:
            return adaptor_dict
        inputs_embeds = adaptor_dict["language_inputs"]
        input_ids = adaptor_dict["language__ids"]
        batch, seq_len = input_ids.shape
        hidden = inputs_embeds.shape[-1]

        # ---- waypoint placeholders ------------------------------------------------------------------------
        first_added = self.tokenizer.additional_special_tokens_ids[0]
        if placeholder_values is not None and len(placeholder_values) > 0:
            ids_host = input_ids.detach().cpu()            # one D2H sync (the reference has two, :55 and :78)
            hits = (ids_host >= first_added).nonzero().tolist()
            first_pos = {}
            for b, l in hits:
                first_pos.setdefault((b, int(ids_host[b, l])), l)
            # as the reference (:80): a special id present in a row (past position 0) without a value for that row is a KeyError
            jobs = [(b, sid, l) for (b, sid), l in sorted(first_pos.items()) if l != 0]
            if jobs:
                enc_dtype = wp_encoder.mlp[0].weight.dtype
                coords = [torch.as_tensor(placeholder_values[b][sid], dtype=torch.float32).reshape(-1, 2) for b, sid, _ in jobs]
                lengths = [c.shape[0] for c in coords]
                for (b, sid, l), n in zip(jobs, lengths):
                    if l + n > seq_len:   # the reference's slice assignment (:91) raises on the size mismatch
                        raise RuntimeError(f"placeholder run of token {sid} in row {b} does not fit: {l} + {n} > {seq_len}")
                flat = torch.cat(coords).to(device=input_ids.device).to(enc_dtype)
                wp = wp_encoder(flat.unsqueeze(0)).squeeze(0).to(inputs_embeds.dtype)
                rows = torch.cat([torch.arange(l, l + n) + b * seq_len for (b, _, l), n in zip(jobs, lengths)]).to(input_ids.device)
                inputs_embeds = _scatter_rows(inputs_embeds.reshape(batch * seq_len, hidden), rows, wp).view(batch, seq_len, hidden)

        # ---- image features -------------------------------------------------------------------------------
        # host-visible indices first (each is a device sync): taken before the vision tower is queued, so the CPU never
        # waits on the heavy kernels and can queue the whole forward + loss + backward behind them
        starts = adaptor_dict["perm"][:, 0].tolist()
        if pixel_values is not None and seq_len != 1 and pixel_values.size(0) > 0:
            BS, T, NP, C, H, W = pixel_values.shape
            assert T == 1, "Only one frame is supported for now"
            rows = (input_ids.reshape(-1) == self.img_context_token_id).nonzero().squeeze(1)
            vit_embeds = self.model.extract_feature(pixel_values.reshape(BS * NP, C, H, W)).reshape(-1, hidden)
            if rows.numel() != vit_embeds.shape[0]:
                print(f"warning: {rows.numel()} <IMG_CONTEXT> tokens but {vit_embeds.shape[0]} image features")
                n = min(rows.numel(), vit_embeds.shape[0])
                rows, vit_embeds = rows[:n], vit_embeds[:n]
            inputs_embeds = _scatter_rows(inputs_embeds.reshape(batch * seq_len, hidden), rows,
                                          vit_embeds.to(inputs_embeds.dtype)).view(batch, seq_len, hidden)

        adaptor_dict["language_inputs"] = inputs_embeds
        # language tokens sit at the front of the permuted stream; a left-padded row starts at its first valid token
        full = adaptor_dict["inputs"]
        if not (torch.is_grad_enabled() and inputs_embeds.requires_grad):
            for b, i in enumerate(starts):
                full[b, : seq_len - i] = inputs_embeds[b, i:]
        else:
            pieces = []
            for b, i in enumerate(starts):
                pieces.append(torch.cat([inputs_embeds[b, i:], full[b, seq_len - i:]], 0))
            full = torch.stack(pieces)
        adaptor_dict["inputs"] = full
        return adaptor_dict


def _scatter_rows(dst2d: torch.Tensor, rows: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    """dst[rows[i]] = src[i].  In-place CUDA scatter without autograd; functional index_put under autograd so the
    gradient reaches ``src`` (ViT features / waypoint embeddings) - the reference keeps that path alive with
    ``x * 0.0 + vit_embeds`` (:124)."""
    if torch.is_grad_enabled() and (src.requires_grad or dst2d.requires_grad):
        return dst2d.index_put((rows,), src)
    return _lib.scatter_rows(dst2d, rows, src)
