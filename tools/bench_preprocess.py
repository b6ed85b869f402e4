"""GPU pre-processing vs the reference formulation (PIL + torchvision on the host), agent frame 359x1024."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import preprocess as P
from simlingo_b200.preprocess import preprocess_frames
img = P.synth_camera(359, 1024, 1)
x = torch.from_numpy(img)[None].cuda()
for _ in range(5): preprocess_frames(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): preprocess_frames(x)
e1.record(); torch.cuda.synchronize()
print(f"GPU preprocess_frames (1 frame, device resident): {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")
xh = torch.from_numpy(img)[None].pin_memory()
t0 = time.perf_counter()
for _ in range(50):
    out = preprocess_frames(xh.cuda(non_blocking=True)); torch.cuda.synchronize()
print(f"GPU incl. H2D of the uint8 frame + sync: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms")
try:
    from PIL import Image
    import torchvision.transforms as T
    from torchvision.transforms.functional import InterpolationMode
    tf = T.Compose([T.Resize((448, 448), interpolation=InterpolationMode.BICUBIC), T.ToTensor(), T.Normalize(P.IMAGENET_MEAN, P.IMAGENET_STD)])
    t0 = time.perf_counter()
    for _ in range(10):
        im = Image.fromarray(np.transpose(img, (1, 2, 0))).resize((896, 448))
        tiles = torch.stack([tf(im.crop((i * 448, 0, (i + 1) * 448, 448))) for i in range(2)])
        dev = tiles.to(torch.bfloat16).cuda(); torch.cuda.synchronize()
    print(f"host PIL + torchvision (reference formulation) incl. H2D: {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms")
except Exception as e:
    print("PIL path unavailable:", e)
