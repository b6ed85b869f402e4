"""Checkpoint I/O around the hot path: what reference ``simlingo_training/train.py:104-111`` does before
training / evaluation starts,

    state_dict = get_fp32_state_dict_from_zero_checkpoint(cfg.checkpoint)   # a DeepSpeed ZeRO directory
    state_dict = torch.load(cfg.checkpoint, map_location="cpu")             # or a single file
    model.load_state_dict(state_dict)

without DeepSpeed installed.  The consolidation of a ZeRO stage-1/2 directory is third-party arithmetic (``deepspeed==0.16.2``
pinned in the reference's ``environment.yaml``; ``deepspeed/utils/zero_to_fp32.py``), restated here from its published
algorithm: per optimizer parameter group the ranks' flat fp32 partitions are concatenated in rank order and cut into the
parameters listed (in order) by ``param_shapes``; the tail of every group is padding up to a multiple of ``2 * world_size``;
frozen parameters travel as ``frozen_param_fragments`` of rank 0's model-states file, buffers inside ``module``; tied
parameters are re-linked through ``shared_params``.  DeepSpeed is absent from this image, so this restatement is UNPINNED
against the real implementation (stated in DESIGN.md); ``write_zero2_checkpoint`` produces the same layout for round-trip tests.

Host-side file I/O only - nothing here touches the GPU; ``DrivingModel.load_state_dict`` then copies into the (flat) bf16
parameters in place."""
from __future__ import annotations

import glob
import math
import os
import re
from collections import OrderedDict
from typing import Dict, List, Optional

import torch

_PREFIXES = ("_forward_module.", "module.")   # Lightning's DeepSpeed wrapper / DDP-style wrappers


def _strip(name: str) -> str:
    for p in _PREFIXES:
        while name.startswith(p):
            name = name[len(p):]
    return name


def _natural(path: str):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", os.path.basename(path))]


def load_zero_checkpoint(checkpoint_dir: str, tag: Optional[str] = None) -> "OrderedDict[str, torch.Tensor]":
    """``get_fp32_state_dict_from_zero_checkpoint(checkpoint_dir, tag)`` for ZeRO stage 1 / 2 (the reference trains with
    ``deepspeed_stage_2``, config.py:299): fp32 ``state_dict`` on the CPU keyed by the module's parameter names."""
    if tag is None:
        latest = os.path.join(checkpoint_dir, "latest")
        if not os.path.isfile(latest):
            raise ValueError(f"Unable to find 'latest' file at {latest}")
        tag = open(latest).read().strip()
    ds_dir = os.path.join(checkpoint_dir, tag)
    if not os.path.isdir(ds_dir):
        raise FileNotFoundError(f"Directory '{ds_dir}' doesn't exist")
    optim_files = sorted(glob.glob(os.path.join(ds_dir, "*_optim_states.pt")), key=_natural)
    if not optim_files:
        raise FileNotFoundError(f"can't find *_optim_states.pt files in directory '{ds_dir}'")
    optim = [torch.load(f, map_location="cpu", weights_only=False)["optimizer_state_dict"] for f in optim_files]
    stage = optim[0]["zero_stage"]
    if stage not in (1, 2):
        raise NotImplementedError(f"ZeRO stage {stage} checkpoints are not handled (the reference trains with stage 2)")
    world = optim[0]["partition_count"]
    world = max(world) if isinstance(world, (list, tuple)) else int(world)
    if world != len(optim_files):
        raise ValueError(f"Expected {world} of '*_optim_states.pt' under '{ds_dir}' but found {len(optim_files)} files")
    model_files = sorted(glob.glob(os.path.join(ds_dir, "*_model_states.pt")), key=_natural)
    if not model_files:
        raise FileNotFoundError(f"can't find *_model_states.pt files in directory '{ds_dir}'")
    ms = torch.load(model_files[0], map_location="cpu", weights_only=False)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    module = ms.get("module") or {}
    for name in ms.get("buffer_names", []):
        if name in module:
            out[name] = module[name].float()
    for name, frag in (ms.get("frozen_param_fragments") or {}).items():
        shape = ms["frozen_param_shapes"][name]
        out[name] = frag.float().reshape(tuple(shape))
    groups = [o["single_partition_of_fp32_groups"] for o in optim]
    param_shapes = ms["param_shapes"]
    if isinstance(param_shapes, dict):
        param_shapes = [param_shapes]
    if len(param_shapes) != len(groups[0]):
        raise ValueError(f"{len(param_shapes)} parameter groups in the model states but {len(groups[0])} in the optimizer states")
    align = 2 * world
    for gi, shapes in enumerate(param_shapes):
        flat = torch.cat([g[gi].reshape(-1) for g in groups], 0)
        offset = 0
        for name, shape in shapes.items():
            n = int(math.prod(tuple(shape)))
            if offset + n > flat.numel():
                raise ValueError(f"parameter {name} runs past the end of group {gi}: {offset} + {n} > {flat.numel()}")
            out[name] = flat.narrow(0, offset, n).view(tuple(shape)).clone()
            offset += n
        if align * math.ceil(offset / align) != align * math.ceil(flat.numel() / align):
            raise ValueError(f"consumed {offset} numels out of {flat.numel()} in group {gi} - something is wrong")
    for pair in (ms.get("shared_params") or {}).items() if isinstance(ms.get("shared_params"), dict) else (ms.get("shared_params") or []):
        if pair[1] in out:
            out[pair[0]] = out[pair[1]]
    return OrderedDict((_strip(k), v) for k, v in out.items())


def load_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    """What ``train.py:104-111`` feeds to ``model.load_state_dict``: a ZeRO directory is consolidated, a single file is
    ``torch.load``-ed (a Lightning checkpoint's ``state_dict`` entry is unwrapped)."""
    if os.path.isdir(path):
        return load_zero_checkpoint(path)
    sd = torch.load(path, map_location="cpu", weights_only=False)
    if isinstance(sd, dict) and "state_dict" in sd and all(isinstance(k, str) for k in sd["state_dict"]):
        sd = sd["state_dict"]
    return OrderedDict((_strip(k), v) for k, v in sd.items())


def write_zero2_checkpoint(model: torch.nn.Module, checkpoint_dir: str, world: int, tag: str = "checkpoint", prefix: str = "_forward_module.") -> None:
    """Writes ``model``'s parameters in the ZeRO-2 directory layout (one parameter group; trainable parameters as flat fp32
    partitions padded to ``2 * world``, frozen ones as fragments) - the inverse of ``load_zero_checkpoint``, used by the
    round-trip tests and to hand a checkpoint to the reference's ``train.py`` / ``eval.py``."""
    ds_dir = os.path.join(checkpoint_dir, tag)
    os.makedirs(ds_dir, exist_ok=True)
    seen: Dict[int, str] = {}
    shapes: "OrderedDict[str, torch.Size]" = OrderedDict()
    frozen_shapes, frozen_frags, shared, flat = OrderedDict(), OrderedDict(), {}, []
    for name, p in model.named_parameters(remove_duplicate=False):
        key = prefix + name
        if id(p) in seen:
            shared[key] = seen[id(p)]
            continue
        seen[id(p)] = key
        if p.requires_grad:
            shapes[key] = p.shape
            flat.append(p.detach().float().cpu().reshape(-1))
        else:
            frozen_shapes[key] = p.shape
            frozen_frags[key] = p.detach().cpu().reshape(-1).clone()
    full = torch.cat(flat) if flat else torch.zeros(0)
    align = 2 * world
    padded = align * math.ceil(full.numel() / align)
    full = torch.cat([full, torch.zeros(padded - full.numel())])
    part = padded // world
    buffers = {prefix + n: b.detach().cpu() for n, b in model.named_buffers()}
    torch.save({"module": dict(buffers), "buffer_names": list(buffers), "param_shapes": [shapes], "frozen_param_shapes": frozen_shapes,
                "frozen_param_fragments": frozen_frags, "shared_params": shared, "ds_version": "0.16.2"},
               os.path.join(ds_dir, "mp_rank_00_model_states.pt"))
    for r in range(world):
        torch.save({"optimizer_state_dict": {"zero_stage": 2, "partition_count": world,
                                             "single_partition_of_fp32_groups": [full[r * part:(r + 1) * part].clone()]}, "ds_version": "0.16.2"},
                   os.path.join(ds_dir, f"zero_pp_rank_{r}_mp_rank_00_optim_states.pt"))
    with open(os.path.join(checkpoint_dir, "latest"), "w") as f:
        f.write(tag)
