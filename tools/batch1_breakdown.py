"""Per-kernel durations of ONE batch-1 CFG sampling step (BASELINE config 4), eager launches so that
`ncu --metrics gpu__time_duration.sum` lists them:
    B1_EAGER_ONLY=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/batch1_breakdown.py
Also prints the CUDA-graph replay time per step for the same state (plain run)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import camels_diffusion_model_b200 as cdm
from camels_diffusion_model_b200 import diffusion as D

dev = torch.device("cuda")
T, NCF = 1500, 6
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.manual_seed(0)
model = cdm.ContextUnet(1, 128, NCF, 64).to(dev).eval()
sched = D.make_schedule(T)
g = torch.Generator().manual_seed(0)
tab = D.draw_shortcut_table(T, 2, 128)
run = D._SamplerRun(model, torch.randn(B, 1, 64, 64, generator=g).to(dev), torch.rand(B, NCF, generator=g).to(dev),
                    2.0, T, sched, shortcut_tab=tab, seed=1)
if os.environ.get("B1_EAGER_ONLY") != "1":   # under ncu: skip the graph replays (6000 launches to profile)
    run.capture()
    run.run(20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run.run(200); e1.record(); torch.cuda.synchronize()
    print(f"graph replay, batch {B} guide_w 2: {e0.elapsed_time(e1) / 200:.4f} ms/step", flush=True)
else:
    run._one_step(); run._one_step()
torch.cuda.synchronize()
print("EAGER-STEPS-BEGIN", flush=True)
for _ in range(3):   # the last 3 x 28 launches of the process = three eager steps
    run._one_step()
torch.cuda.synchronize()
print("EAGER-STEPS-END", flush=True)
