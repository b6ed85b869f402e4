"""Generates ``tests/golden/collate.pt`` by running the REFERENCE's own batch-assembly code in this container:

* ``simlingo_training/utils/internvl2_utils.py::get_chat_tokens`` / ``get_assistant_loss_mask`` / ``get_custom_chat_template``
  (:29-175) and
* ``simlingo_training/dataloader/datamodule.py::DataModule.dl_collate_fn`` (:310-443), called unbound on a namespace that
  carries the attributes it reads from ``self``

imported unmodified; ``hydra``, ``pytorch_lightning`` and ``line_profiler`` are stubbed (they do no arithmetic here).  The one
upstream piece that cannot be fetched offline is ``conversation.py`` of the HF-Hub repo OpenGVLab/InternVL2-1B, which the
reference executes from its cache directory (:111-120): a stand-in with the ``internlm2-chat`` template (system template, roles,
``<|im_end|>`` separator, MPT style) is written to a temp dir - so the TEMPLATE STRING is unpinned, everything downstream of it
(tokenised ids, validity, loss mask, placeholder dicts, labels, run ids, calibration matrices, Pillow tiles) is the reference's.
The tokenizer is ``tests.helpers.StubChatTokenizer`` (no tokenizer files offline).

    python tests/golden/make_golden_collate.py
"""
import hashlib
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from tests.helpers import StubChatTokenizer, make_dataset_outputs  # noqa: E402

CONVERSATION_PY = '''
class Conversation:
    def __init__(self):
        self.system_template = "<|im_start|>system\\n{system_message}"
        self.system_message = "你是由上海人工智能实验室联合商汤科技开发的书生多模态大模型，英文名叫InternVL, 是一个有用无害的人工智能助手。"
        self.roles = ("<|im_start|>user\\n", "<|im_start|>assistant\\n")
        self.sep = "<|im_end|>"
        self.messages = []
    def append_message(self, role, message):
        self.messages.append([role, message])
    def get_prompt(self):
        ret = self.system_template.format(system_message=self.system_message) + self.sep
        for role, message in self.messages:
            ret += role + message + self.sep if message else role
        return ret
def get_conv_template(name):
    assert name == "internlm2-chat"
    return Conversation()
'''

tmp = tempfile.mkdtemp()
os.makedirs(os.path.join(tmp, "InternVL2-1B"))
open(os.path.join(tmp, "InternVL2-1B", "conversation.py"), "w").write(CONVERSATION_PY)


def stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


hydra = stub("hydra", main=lambda **kw: (lambda f: f))
hydra.utils = stub("hydra.utils", to_absolute_path=lambda p: p, instantiate=None)
stub("line_profiler", profile=lambda f: f)
pl = stub("pytorch_lightning", LightningDataModule=object)
# the reference's package is importable as a namespace from /root/reference, but our drop-in has the same name: load by path
ref_root = "/root/reference/simlingo_training"


def load(modname, path):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


for k in [k for k in sys.modules if k.startswith("simlingo_training")]:
    del sys.modules[k]
pkg = stub("simlingo_training"); pkg.__path__ = [ref_root]
stub("simlingo_training.utils").__path__ = [ref_root + "/utils"]
stub("simlingo_training.dataloader").__path__ = [ref_root + "/dataloader"]
load("simlingo_training.utils.custom_types", ref_root + "/utils/custom_types.py")
ref_utils = load("simlingo_training.utils.internvl2_utils", ref_root + "/utils/internvl2_utils.py")
load("simlingo_training.utils.projection", ref_root + "/utils/projection.py")
ref_dm = load("simlingo_training.dataloader.datamodule", ref_root + "/dataloader/datamodule.py")

out = {"chat": [], "collate": []}
roles = ("<|im_start|>user\n", "<|im_start|>assistant\n")
for side in ("left", "right"):
    tok = StubChatTokenizer(side)
    prompts = ["<|im_start|>user\n<img><IMG_CONTEXT><IMG_CONTEXT></img>\nWhere to?<|im_end|><|im_start|>assistant\nStraight ahead now<|im_end|>",
               "<|im_start|>user\nStop?<|im_end|><|im_start|>assistant\nNo<|im_end|>",
               "<|im_start|>user\nA much longer question about the road ahead ?<|im_end|><|im_start|>assistant\n"]
    r = ref_utils.get_chat_tokens(tok, prompts, *roles)
    out["chat"].append(dict(side=side, prompts=prompts, phrase_ids=r["phrase_ids"], phrase_valid=r["phrase_valid"], loss_masking=r["loss_masking"]))

for side, predict, n, seed in (("left", False, 3, 1), ("left", True, 2, 2)):
    tok = StubChatTokenizer(side)
    fake_self = types.SimpleNamespace(NUM_IMAGE_PATCHES=2, IMAGES_TO_CONSIDER=["image_ff"], encoder_variant="OpenGVLab/InternVL2-1B", use_global_img=False,
                                      tokenizer=tok, num_image_tokens_total=512, base_dataset=types.SimpleNamespace(use_1d_wps=False), predict=predict)
    data = make_dataset_outputs(n, seed)
    cwd = os.getcwd()
    os.chdir(tmp)   # get_custom_chat_template looks for pretrained/<variant>/conversation.py relative to the working directory
    os.makedirs("pretrained", exist_ok=True)
    if not os.path.exists("pretrained/InternVL2-1B"):
        os.symlink(os.path.join(tmp, "InternVL2-1B"), "pretrained/InternVL2-1B")
    try:
        ex = ref_dm.DataModule.dl_collate_fn(fake_self, data)
    finally:
        os.chdir(cwd)
    di, dl = ex.driving_input, ex.driving_label
    cam = di.camera_images   # [B, 1, 2, 3, 448, 448] float32 from Pillow + torchvision
    mean = torch.tensor(ref_utils.IMAGENET_MEAN).view(1, 1, 1, 3, 1, 1)
    std = torch.tensor(ref_utils.IMAGENET_STD).view(1, 1, 1, 3, 1, 1)
    u8 = torch.round((cam * std + mean) * 255.0).to(torch.uint8)
    rec = dict(side=side, predict=predict, n=n, seed=seed, camera_shape=tuple(cam.shape),
               camera_sha256=hashlib.sha256(np.ascontiguousarray(u8.numpy()).tobytes()).hexdigest(), image_sizes=di.image_sizes,
               camera_intrinsics=di.camera_intrinsics, camera_extrinsics=di.camera_extrinsics, vehicle_speed=di.vehicle_speed, target_point=di.target_point,
               waypoints=dl.waypoints, path=dl.path, image_ff_org=dl.image_ff_org, eval_infos=dl.eval_infos, run_id=ex.run_id, qa_templates=ex.qa_templates,
               answer_strings=dl.answer.language_string)
    for name, lab in (("prompt", di.prompt), ("prompt_inference", di.prompt_inference)):
        rec[name] = dict(phrase_ids=lab.phrase_ids, phrase_valid=lab.phrase_valid, phrase_mask=lab.phrase_mask, loss_masking=lab.loss_masking,
                         language_string=lab.language_string, placeholder_values=[{k: np.asarray(v) for k, v in d.items()} for d in lab.placeholder_values])
    out["collate"].append(rec)
    print(side, predict, tuple(di.prompt.phrase_ids.shape), int(di.prompt.loss_masking.sum()), tuple(cam.shape))
torch.save(out, os.path.join(os.path.dirname(__file__), "collate.pt"))
