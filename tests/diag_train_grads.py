"""Diagnostic: per-parameter gradient error of one training step vs the CPU oracle's autograd."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import camels_diffusion_model_b200 as cdm
from oracle import contextunet_oracle as O
from tests._util import NCF, T, cal_sd, load, rel_l2, split_shortcut

g = load("train_step.npz")
sd = cal_sd()
model = cdm.ContextUnet(1, 128, NCF, 64)
model.load_state_dict(sd)
model = model.cuda().train()
b_t, a_t, ab_t = cdm.make_schedule(1500)
x, param, noise, t = T(g["x"]), T(g["param"]), T(g["noise"]), T(g["t"])
x_pert = cdm.perturb_input(x, t, noise, ab_t)
pred = model(x_pert, (t / 1500).cuda(), param.cuda(), shortcut=T(g["shortcut"]))
loss = F.mse_loss(pred, noise.cuda())
loss.backward()
_, _, ab_cpu = O.make_schedule(1500)
_, grads, _ = O.train_step(sd, x, param, t, noise, split_shortcut(g["shortcut"]), 1500, ab_cpu, n_cfeat=NCF)
_, g16, _ = O.train_step(sd, x, param, t, noise, split_shortcut(g["shortcut"]), 1500, ab_cpu, n_cfeat=NCF,
                         emulate_bf16=True)
for name, p in model.named_parameters():
    ref = grads[name]
    print(f"{name:40s} err32 {rel_l2(p.grad, ref):9.3e} err16 {rel_l2(p.grad, g16[name]):9.3e} |ref| {float(ref.norm()):9.3e}  |got| {float(p.grad.norm()):9.3e}")
