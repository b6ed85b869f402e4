"""Timeline of the persistent decode kernel (csrc/decode.cu): every CTA stamps its SM clock at each grid barrier's entry and exit
(debug hook slb_debug_set_trace); prints, per phase type, the critical path (slowest CTA's work between two barriers) and the
barrier's own latency (shortest wait = the last arriver's), for the agent (batch 1) and language (batch 32) cases and for the
tuning switches in SLB_DECODE_FLAGS.      python tools/trace_decode.py [flags ...]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as Bn
from simlingo_b200 import lib, spec as S

NB = 320   # kTraceBarriers
spec = S.INTERNVL2_1B
dev = torch.device("cuda", 0)
model = Bn.build_model(spec, dev)
eng = model._engine()
L = lib.load()
grid = L.slb_num_sms()
FS = 32   # kFineStamps
buf = torch.zeros(8192 + grid * NB * 2 + 16 + grid * FS, dtype=torch.int64, device=dev)
L.slb_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
PH = ["1 norm+qkv", "2 attention", "3 o-proj", "4 norm+gate|up", "5 down"]


def analyse(tag, n_layers):
    torch.cuda.synchronize()
    t = buf[8192:8192 + grid * NB * 2].view(grid, NB, 2).cpu().double()
    st = buf[8192 + grid * NB * 2:][:4].cpu().tolist()
    per_tok = n_layers * 5 + 2
    ghz = (st[3] - st[1]) / max(st[2] - st[0], 1)
    entry, exit_ = t[:, :, 0], t[:, :, 1]
    nb = int((entry[0] > 0).sum())
    print(f"--- {tag}: kernel {1e-3 * (st[2] - st[0]):.1f} us, SM clock {ghz:.2f} GHz, {nb} barriers traced ({nb / per_tok:.1f} tokens)")
    if nb < 2 * per_tok:
        print("    fewer than two tokens traced")
        return
    work = (entry[:, 1:nb] - exit_[:, :nb - 1]) / ghz * 1e-3    # us, barrier i (i >= 1)
    wait = (exit_[:, 1:nb] - entry[:, 1:nb]) / ghz * 1e-3
    i0 = per_tok - 1   # second token: barriers per_tok .. 2 per_tok - 1  -> indices in work[] shifted by one
    tot = 0.0
    rows = []
    for ph in range(5):
        idx = [i0 + l * 5 + ph for l in range(n_layers)]
        w, q = work[:, idx], wait[:, idx]
        rows.append((PH[ph], w.max(0).values.mean().item(), w.mean().item(), q.min(0).values.mean().item(), q[0].mean().item()))
        tot += (w.max(0).values + q.min(0).values).sum().item()
    for name, j in (("lm_head + arg-max", i0 + n_layers * 5), ("sample (CTA 0)", i0 + n_layers * 5 + 1)):
        w, q = work[:, j], wait[:, j]
        rows.append((name, w.max().item(), w.mean().item(), q.min().item(), q[0].item()))
        tot += w.max().item() + q.min().item()
    print(f"    {'phase':20s} {'slowest CTA':>12s} {'mean CTA':>10s} {'barrier':>9s} {'CTA0 wait':>10s}   (us; per layer for phases 1-5)")
    for r in rows:
        print(f"    {r[0]:20s} {r[1]:12.2f} {r[2]:10.2f} {r[3]:9.2f} {r[4]:10.2f}")
    print(f"    critical path of the second token: {tot:.1f} us")
    # stamps inside the phases of (second token, second layer): the CTA that was slowest in each phase
    f = buf[8192 + grid * NB * 2 + 16:][:grid * FS].view(grid, FS).cpu().double()
    us = lambda a, b: (b - a) / ghz * 1e-3

    def show(name, ids, labels, cta):
        r = f[cta]
        parts = [f"{lb} {us(r[i0], r[i1]):.2f}" for (i0, i1), lb in zip(ids, labels) if r[i0] > 0 and r[i1] > 0]
        print(f"    {name} (CTA {cta}): " + ", ".join(parts))
    att = (f[:, 8] - f[:, 0]).clamp_min(0) * (f[:, 8] > 0)
    show("attention", [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 8)],
         ["q load + sincos", "rotate + K/V wait", "scores", "softmax", "PV", "partial store", "counter", "merge"], int(att.argmax()))
    print(f"      merged by this CTA: {int(f[int(att.argmax()), 9])}; CTAs with an item: {int((f[:, 8] > 0).sum())}")
    g4 = (f[:, 19] if (f[:, 19] > 0).any() else f[:, 15]) - f[:, 10]
    c4 = int(((f[:, 19].clamp_min(0) + f[:, 15]) - f[:, 10]).argmax())
    show("norm + gate|up", [(10, 11), (11, 12), (12, 13), (13, 14), (14, 15), (15, 16), (16, 17), (17, 18), (18, 19)],
         ["norm", "weights wait r0", "MMA r0", "stage next", "reduce r0", "weights wait r1", "MMA r1", "stage next", "reduce r1"], c4)
    c1 = int((f[:, 25] - f[:, 20]).argmax())
    show("norm + qkv", [(20, 21), (21, 22), (22, 23), (23, 24), (24, 25)], ["norm", "weights wait", "MMA", "stage next", "reduce"], c1)


def timed(fn, n=3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


flags_list = [int(a) for a in sys.argv[1:]] or [0, 1, 2]
ex1 = Bn.make_example(Bn.host_agent_batch(spec, 1, 99, 25), dev)
ex1b = Bn.make_example(Bn.host_agent_batch(spec, 1, 99, 1), dev)
hb = Bn.host_agent_batch(spec, 32, 500, None)
ids, fr, vd = hb["ids"].to(dev), hb["frames"].to(dev), hb["valid"].to(dev)
lang = lambda: eng.driving_forward(fr, ids, vd, hb["placeholders"], max_new_tokens=16, eos_token_id=None, ids_cpu=hb["ids"])
for mega in ((False, True) if os.environ.get("TRACE_CHAIN", "0") == "1" else (True,)):
    for flags in (flags_list if mega else [0]):
        os.environ["SLB_DECODE_FLAGS"] = str(flags)
        eng.decode_mega = mega
        for k in [k for k in eng._graphs if k[0] == "gen"]:   # re-capture the generation graphs (the ViT graph keeps the pool alive)
            eng._graphs.pop(k)
        for _ in range(3):
            model(ex1); model(ex1b)
        t25, t1 = timed(lambda: model(ex1)), timed(lambda: model(ex1b))
        print(f"=== mega={mega} flags={flags}: agent G=25 {t25:.2f} ms, G=1 {t1:.2f} ms -> {(t25 - t1) / 24 * 1e3:.0f} us per decoded token")
        if mega:
            buf.zero_(); model(ex1); analyse(f"agent batch 1 flags={flags}", spec.llm_layers)
        for _ in range(3):
            lang()
        t16 = timed(lang)
        print(f"=== mega={mega} flags={flags}: language batch 32, 16 tokens {t16:.2f} ms")
        if mega:
            buf.zero_(); lang(); analyse(f"language batch 32 flags={flags}", spec.llm_layers)
L.slb_debug_set_trace(None)
