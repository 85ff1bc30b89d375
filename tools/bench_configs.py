"""Secondary measurements for the BASELINE configs that are not the headline bench line (1 GPU):
cfg 4 (batch-1 / batch-25 sampling latency per step), cfg 5 (all-timestep NLL sweep throughput),
cfg 3 (training step, global batch 256 on one GPU).  Prints one JSON object."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import camels_diffusion_model_b200 as cdm
from camels_diffusion_model_b200 import diffusion as D, train as TR

dev = torch.device("cuda")
T, NCF = 1500, 6
torch.manual_seed(0)
model = cdm.ContextUnet(1, 128, NCF, 64).to(dev).eval()
sched = D.make_schedule(T)
out = {}
g = torch.Generator().manual_seed(0)


def ev():
    return torch.cuda.Event(enable_timing=True)


def sampler_ms(B, gw, steps=40):
    tab = D.draw_shortcut_table(T, 2 if gw > 0 else 1, 128)
    run = D._SamplerRun(model, torch.randn(B, 1, 64, 64, generator=g).to(dev), torch.rand(B, NCF, generator=g).to(dev),
                        gw, T, sched, shortcut_tab=tab, seed=1)
    run.capture()
    run.run(5)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record(); run.run(steps); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

for B, gw in ((1, 0.0), (1, 2.0), (5, 2.0), (25, 0.0), (4, 2.0)):
    ms = sampler_ms(B, gw)
    out[f"cfg4_sampling_B{B}_w{gw:g}"] = {"ms_per_step": round(ms, 4), "samples_per_s_1500_steps": round(B / (ms * 1.5), 4),
                                          "forwards_per_step": 2 if gw > 0 else 1}
# cfg 5: all-timestep NLL, 512 maps on one GPU (the per-GPU share of 4096 maps on 8 GPUs)
B = 512
loop = D._EvalLoop(model, torch.rand(B, 1, 64, 64, generator=g), torch.rand(B, NCF, generator=g), T, sched, "one_minus",
                   1.0 / (2 * sched[0].float()), seed=3)
loop.step.fill_(1)
loop.one(); torch.cuda.synchronize()
gph = torch.cuda.CUDAGraph()
with torch.cuda.graph(gph):
    loop.one(); TR.L.step_advance(loop.step, 1)
loop.step.fill_(1)
for _ in range(5): gph.replay()
torch.cuda.synchronize()
e0, e1 = ev(), ev()
K = 40
e0.record()
for _ in range(K): gph.replay()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
out["cfg5_nll_sweep_B512"] = {"ms_per_timestep": round(ms, 3), "image_forwards_per_s": round(B / ms * 1e3, 1),
                              "maps_per_s_all_1500_timesteps": round(B / (ms * 1.5), 3),
                              "tflops": round(19.1785e9 * B / (ms * 1e-3) / 1e12, 1)}
# cfg 3: training step, global batch 256 on ONE GPU (see tools/dp_check.py for the multi-GPU run)
model.train()
opt = TR.FusedAdam(model.parameters(), lr=1e-5)
xb, pb = torch.rand(256, 1, 64, 64, generator=g).to(dev), torch.rand(256, NCF, generator=g).to(dev)
sc = torch.rand(256, generator=g) * 2 - 1
for _ in range(3): TR.training_step(model, opt, xb, pb, T, sched[2], shortcut=sc)
torch.cuda.synchronize()
e0, e1 = ev(), ev()
e0.record()
for _ in range(10): TR.training_step(model, opt, xb, pb, T, sched[2], shortcut=sc)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
out["cfg3_train_step_B256_1gpu"] = {"ms_per_step": round(ms, 2), "img_per_s": round(256 / ms * 1e3, 1),
                                    "tflops_3x_forward": round(3 * 19.1785e9 * 256 / (ms * 1e-3) / 1e12, 1)}
for B in (256, 32):
    torch.manual_seed(0)
    model = cdm.ContextUnet(1, 128, NCF, 64).to(dev).train()
    gstep = TR.GraphedTrainStep(model, B, T, sched[2], lr=1e-5)
    xb, pb = torch.rand(B, 1, 64, 64, generator=g).to(dev), torch.rand(B, NCF, generator=g).to(dev)
    tb = torch.randint(1, T + 1, (B,), generator=g).to(dev)
    for _ in range(3): gstep(xb, pb, t=tb, shortcut=sc)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(10): loss = gstep(xb, pb, t=tb, shortcut=sc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out[f"cfg3_train_step_graph_B{B}_1gpu"] = {"ms_per_step": round(ms, 2), "img_per_s": round(B / ms * 1e3, 1),
                                               "tflops_3x_forward": round(3 * 19.1785e9 * B / (ms * 1e-3) / 1e12, 1),
                                               "loss": round(float(loss), 4)}
    del gstep, model
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
