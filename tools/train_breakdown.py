"""CUDA-event breakdown of one eager training step (batch given on the command line), aggregated per C-ABI entry."""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import camels_diffusion_model_b200 as cdm
from camels_diffusion_model_b200 import _lib as L, train as TR, diffusion as D

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
torch.manual_seed(0)
model = cdm.ContextUnet(1, 128, 6, 64).to(dev).train()
sched = D.make_schedule(1500)
step = TR.GraphedTrainStep(model, B, 1500, sched[2], lr=1e-5, use_graph=False)
g = torch.Generator().manual_seed(0)
x, p = torch.rand(B, 1, 64, 64, generator=g).to(dev), torch.rand(B, 6, generator=g).to(dev)
t = torch.randint(1, 1501, (B,), generator=g).to(dev)
sc = torch.rand(256, generator=g) * 2 - 1
for _ in range(2): step(x, p, t=t, shortcut=sc)
torch.cuda.synchronize()
names = [n for n in dir(L) if callable(getattr(L, n)) and getattr(L.lib(), "cdm_" + n, None) is not None]
recs, orig = [], {}
for n in names:
    fn = getattr(L, n); orig[n] = fn
    def wrap(*a, _fn=fn, _n=n, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = _fn(*a, **k); e1.record()
        tag = _n
        if _n == "gemm_tn": tag += f" taps{k.get('taps', 1)}"
        if _n == "conv3x3": tag += f" H{a[0].shape[1]}"
        recs.append((tag, e0, e1)); return r
    setattr(L, n, wrap)
e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e_all0.record(); step(x, p, t=t, shortcut=sc); e_all1.record(); torch.cuda.synchronize()
for n, fn in orig.items(): setattr(L, n, fn)
agg = collections.OrderedDict()
for tag, e0, e1 in recs:
    agg.setdefault(tag, [0, 0.0]); agg[tag][0] += 1; agg[tag][1] += e0.elapsed_time(e1)
tot = sum(v for _, v in agg.values())
print(f"batch {B}: step {e_all0.elapsed_time(e_all1):.2f} ms, inside C-ABI launches {tot:.2f} ms")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v:9.3f} ms {100 * v / tot:5.1f}%  x{n:3d}  {k}")
