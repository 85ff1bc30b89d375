"""World-size-N worker of tests/test_gpu_multi.py (launched with torchrun, one process per GPU, NCCL).

Data-parallel training (BASELINE config 3; reference step: code/train_diffusion_paper.py:349-366, which the
north_star extends with cross-rank BatchNorm statistics + gradient all-reduce): the N-rank step on a sharded global
batch must reproduce the single-process step on the whole batch.  Every check is an assertion; rank 0 prints
`DP-WORKER OK` at the end."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import camels_diffusion_model_b200 as cdm  # noqa: E402
from camels_diffusion_model_b200 import _lib as L, diffusion as D, parallel as P, train as TR  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
NCF, T = 6, 1500


def all_equal(t):
    g = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(g, t.contiguous())
    return all(torch.equal(g[0], x) for x in g)


# ---- 1. the fused reduce + cross-rank exchange kernel: exact, and bit-identical on every rank
px = P.PeerExchange(dev)
gx = torch.Generator().manual_seed(123)
parts = [torch.randn(37, 300, generator=gx) for _ in range(world)]
out = torch.empty(300, device=dev)
for it in range(20):
    L.xrank_sum((parts[rank] * (it + 1)).to(dev), out, xr=px.args)
    expect = torch.zeros(300, device=dev)
    for r in range(world):  # rank order; each rank's fold computed by the same kernel (world = 1 path)
        loc = torch.empty(300, device=dev)
        L.xrank_sum((parts[r] * (it + 1)).to(dev), loc)
        expect += loc
    assert torch.equal(out, expect), f"xrank_sum differs from the rank-ordered sum (exchange {it})"
    assert all_equal(out), "xrank_sum must be bit-identical on all ranks"
del px

# ---- 2. one sharded training step vs the single-process step on the global batch
torch.manual_seed(0)
ref_model = cdm.ContextUnet(1, 128, NCF, 64)
g = torch.Generator().manual_seed(1)
for k, v in ref_model.state_dict().items():  # non-trivial norm layers
    if k.endswith(".1.weight") and v.dim() == 1:
        v.uniform_(0.5, 1.5, generator=g)
    if k.endswith(".1.bias") and v.dim() == 1:
        v.normal_(0, 0.2, generator=g)
sd = {k: v.clone() for k, v in ref_model.state_dict().items()}
b_t, a_t, ab_t = cdm.make_schedule(T, device=dev)
per = 4
B = per * world
x, prm = torch.rand(B, 1, 64, 64, generator=g), torch.rand(B, NCF, generator=g)
noise, t = torch.randn(B, 1, 64, 64, generator=g), torch.randint(1, T + 1, (B,), generator=g)
sc = torch.rand(256, generator=g) * 2 - 1


def step(xs, ps, ns, ts, dp):
    TR.DATA_PARALLEL = dp
    m = cdm.ContextUnet(1, 128, NCF, 64)
    m.load_state_dict(sd)
    m = m.to(dev).train()
    xp = cdm.perturb_input(xs, ts, ns, ab_t)
    pred = m(xp, (ts / T).to(dev), ps.to(dev), shortcut=sc)
    loss = F.mse_loss(pred, ns.to(dev))
    loss.backward()
    return m, float(loss)


s, e = rank * per, (rank + 1) * per
m_1, loss_1 = step(x, prm, noise, t, False)
for mode in ("nccl", "peer"):
    TR.PEER_EXCHANGE = mode == "peer"
    m_dp, loss_dp = step(x[s:e], prm[s:e], noise[s:e], t[s:e], True)
    if mode == "peer":
        assert TR.PEER is not None, "the peer-memory exchange must be the path that ran"
    lt = torch.tensor([loss_dp], device=dev, dtype=torch.float64)
    dist.all_reduce(lt)
    assert abs(float(lt) / world - loss_1) <= 1e-4 * abs(loss_1) + 1e-6, (mode, float(lt) / world, loss_1)
    for (k, b1), (_, b2) in zip(m_1.named_buffers(), m_dp.named_buffers()):
        if "running" in k:
            assert float((b1 - b2).abs().max()) <= 1e-4, f"[{mode}] {k}: running statistics differ from the global batch"
        else:
            assert torch.equal(b1, b2), k
    if rank == 0 and mode == "nccl":  # diagnostic: batch statistics as seen through the running buffers (momentum 0.1)
        rows = []
        for (k, b1), (_, b2) in zip(m_1.named_buffers(), m_dp.named_buffers()):
            if "running_var" in k:
                d = ((b1 - b2).abs() / (b1 - 0.9).abs().clamp_min(1e-12)).max()  # relative to 0.1 * batch var
                rows.append((float(d), k))
        rows.sort(reverse=True)
        print("DP-DIAG worst relative batch-variance differences (sharded vs single):", [(f"{d:.1e}", k) for d, k in rows[:4]])
    g1, g2 = dict(m_1.named_parameters()), dict(m_dp.named_parameters())
    # gradients nearest the loss are a sharp check of the exchange logic (deeper ones amplify bf16 ReLU-mask flips
    # ~1.3x per layer in this random-init train-mode BatchNorm stack)
    for k, tol in (("out.3.weight", 2e-3), ("out.1.weight", 5e-3), ("out.0.weight", 1e-2),
                   ("up2.model.2.conv2.1.weight", 2e-2)):
        err = float((g1[k].grad - g2[k].grad).norm() / g1[k].grad.norm())
        assert err <= tol, f"[{mode}] grad {k}: sharded vs single-process rel-L2 {err:.2e} > {tol}"
    if rank == 0:
        errs = sorted(((float((g1[k].grad - g2[k].grad).norm() / g1[k].grad.norm()), k) for k in g1
                       if float(g1[k].grad.norm()) > 1e-7), reverse=True)
        print(f"DP-DIAG[{mode}] worst gradients (sharded vs single rel-L2):", [(f"{e:.1e}", k) for e, k in errs[:5]],
              "| out.3.weight", f"{dict((k, e) for e, k in errs)['out.3.weight']:.1e}", flush=True)
    for k, p in g2.items():  # the all-reduced gradients and the updated statistics are the same on every rank
        assert all_equal(p.grad), f"[{mode}] grad {k} differs between ranks"

# ---- 3. the captured step: ranks draw DIFFERENT noise (global-sample-keyed Philox) and stay in lock step
TR.PEER_EXCHANGE, TR.DATA_PARALLEL = True, True
torch.manual_seed(0)
model = cdm.ContextUnet(1, 128, NCF, 64).to(dev).train()
gs = TR.GraphedTrainStep(model, per, T, ab_t, lr=1e-5, seed=11)
torch.manual_seed(5)  # CPU generators in step on every rank: same shortcut draw, disjoint slices of the global t draw
losses = []
for _ in range(3):
    losses.append(float(gs(x[s:e].to(dev), prm[s:e].to(dev))))
gn = [torch.empty_like(gs.noise) for _ in range(world)]
dist.all_gather(gn, gs.noise)
for r in range(1, world):
    assert not torch.equal(gn[0], gn[r]), "two ranks drew the same noise"
gt = [torch.empty_like(gs.t) for _ in range(world)]
dist.all_gather(gt, gs.t)
assert not all(torch.equal(gt[0], v) for v in gt[1:]), "two ranks drew the same timesteps"
for k, p in model.named_parameters():
    assert all_equal(p.detach()), f"parameters diverged between ranks after 3 steps: {k}"
assert all(l == l and l < 10 for l in losses)

# ---- 4. sharded all-timestep NLL + ELBO (BASELINE config 5, tiny): equals the single-process result
model.eval()
Tn = 6
sched = D.make_schedule(Tn, device=dev)
maps, mp = torch.rand(4 * world, 1, 64, 64, generator=g), torch.rand(4 * world, NCF, generator=g)
tabs = torch.rand(Tn + 1, 1, 2, 128, generator=g) * 2 - 1


def ev(dl, sample_offset=0):
    tot_n, tot_e, num = 0.0, 0.0, 0
    w = 1.0 / (2 * sched[0].float())
    w2 = 0.5 * (1.0 / (1.0 - sched[2].float()) - 1.0)
    w2[0] = 0
    for xb, pb in dl:
        loop = D._EvalLoop(model, xb, pb, Tn, sched, "one_minus", w, shortcut_tab=tabs, seed=3, weight_tab2=w2,
                           sample_offset=sample_offset + num)
        acc = loop.sweep_all()
        tot_n += acc.sum().item()
        tot_e += loop.acc2.sum().item()
        num += xb.shape[0]
    return tot_n / num, tot_e / num


nll_s, elbo_s = P.evaluate_sharded(ev, maps, mp, 4, dev)
nll_1, elbo_1 = ev([(maps, mp)])
assert abs(nll_s - nll_1) <= 1e-5 * abs(nll_1) and abs(elbo_s - elbo_1) <= 1e-5 * abs(elbo_1), (nll_s, nll_1, elbo_s, elbo_1)

del gs
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    print(f"DP-WORKER OK world={world}", flush=True)
sys.stdout.flush()
sys.stderr.flush()
# every check has passed and every rank has reached the barrier: leave without the interpreter's teardown (destroying
# the process group under live symmetric-memory mappings and captured NCCL graphs can block at exit)
os._exit(0)
