"""GPU check of the optimizer / store ordering guard: two training steps, then a model re-load behind the optimizer must
make ``FusedAdamW.step()`` raise.  Prints one line."""
import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simlingo_b200.spec import tiny_spec
from simlingo_b200.optim import FusedAdamW
from tests.helpers import build_drop_in_model, make_case_inputs, to_driving_example
spec = tiny_spec(2, 2, 4096)
m = build_drop_in_model(spec, "internvl2-tiny-chk").eval()
ex = to_driving_example(make_case_inputs(spec, 2, seed=5, answer_len=16, pad_rows=[(1, 3)]))
opt = FusedAdamW(list(m.parameters()), m.param_store(), lr=3e-3, weight_decay=0.1, max_grad_norm=0.3)
for _ in range(2):
    opt.zero_grad(); m.forward_loss(ex)[0].loss.backward(); opt.step()
torch.cuda.synchronize()
sd = opt.state_dict()
m.load_state_dict({k: v.clone() for k, v in m.state_dict().items()})
m.forward_loss(ex)[0].loss.backward()
try:
    opt.step(); print("NOT DETECTED")
except RuntimeError as e:
    print("step ok x2; orphan detected:", str(e)[:60])
