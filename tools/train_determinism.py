"""Which gradients of a training step are bit-reproducible?  Runs the same step N times from identical state."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import camels_diffusion_model_b200 as cdm
from camels_diffusion_model_b200 import diffusion as D

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
ref = cdm.ContextUnet(1, 128, 6, 64)
sd = {k: v.clone() for k, v in ref.state_dict().items()}
T = 1500
ab_t = D.make_schedule(T)[2]
g = torch.Generator().manual_seed(1)
x, p = torch.rand(B, 1, 64, 64, generator=g).cuda(), torch.rand(B, 6, generator=g).cuda()
noise = torch.randn(B, 1, 64, 64, generator=g).cuda()
t = torch.randint(1, T + 1, (B,), generator=g)
sc = torch.rand(256, generator=g) * 2 - 1
runs = []
for it in range(4):
    m = cdm.ContextUnet(1, 128, 6, 64)
    m.load_state_dict(sd)
    m = m.cuda().train()
    xp = D.perturb_input(x, t, noise, ab_t)
    pred = m(xp, (t / T).cuda(), p, shortcut=sc)
    loss = F.mse_loss(pred, noise)
    loss.backward()
    torch.cuda.synchronize()
    runs.append((pred.detach().clone(), {k: q.grad.detach().clone() for k, q in m.named_parameters()},
                 {k: b.detach().clone() for k, b in m.named_buffers()}))
    if it == 1:  # disturb the allocator between runs
        junk = [torch.randn(1 << 24, device="cuda") for _ in range(3)]
        del junk
bad = {}
for it in range(1, 4):
    if not torch.equal(runs[0][0], runs[it][0]):
        bad["<forward pred>"] = bad.get("<forward pred>", 0) + 1
    for k in runs[0][1]:
        if not torch.equal(runs[0][1][k], runs[it][1][k]):
            bad[k] = bad.get(k, 0) + 1
    for k in runs[0][2]:
        if not torch.equal(runs[0][2][k], runs[it][2][k]):
            bad["buffer " + k] = bad.get("buffer " + k, 0) + 1
print(f"batch {B}: {len(bad)} of {len(runs[0][1])} gradient tensors differ between identical runs")
for k, n in bad.items():
    d = float((runs[0][1][k] - runs[1][1][k]).abs().max()) if k in runs[0][1] else -1
    print(f"   {k}: differs in {n}/3 repeats, max |diff| run0-run1 {d:.3e}")
