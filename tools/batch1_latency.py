"""Per-step latency of small-batch sampling (BASELINE config 4): CUDA-graph replay, CUDA events.
    python tools/batch1_latency.py            (CDM_PDL=0 disables programmatic dependent launch for an A/B run)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import camels_diffusion_model_b200 as cdm
from camels_diffusion_model_b200 import diffusion as D

dev = torch.device("cuda")
T, NCF = 1500, 6
torch.manual_seed(0)
model = cdm.ContextUnet(1, 128, NCF, 64).to(dev).eval()
sched = D.make_schedule(T)
g = torch.Generator().manual_seed(0)
for B, gw in ((1, 0.0), (1, 2.0), (4, 2.0), (8, 2.0), (30, 0.0), (128, 2.0)):
    tab = D.draw_shortcut_table(T, 2 if gw > 0 else 1, 128)
    run = D._SamplerRun(model, torch.randn(B, 1, 64, 64, generator=g).to(dev), torch.rand(B, NCF, generator=g).to(dev),
                        gw, T, sched, shortcut_tab=tab, seed=1)
    run.capture()
    run.run(20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run.run(200); e1.record(); torch.cuda.synchronize()
    print(f"PDL={os.environ.get('CDM_PDL', '1')} batch {B:4d} guide_w {gw:g}: {e0.elapsed_time(e1) / 200:.4f} ms/step", flush=True)
    del run
