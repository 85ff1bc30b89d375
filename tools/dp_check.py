"""Data-parallel training check (run under torchrun on >= 2 GPUs):
the 2-rank step on a sharded global batch (cross-rank BatchNorm statistics + gradient all-reduce) must reproduce
the single-process step on the whole batch.   torchrun --nproc-per-node 2 tools/dp_check.py
Also times a training step (BASELINE config 3: global batch 256) and prints img/s."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import torch.nn.functional as F
import camels_diffusion_model_b200 as cdm
from camels_diffusion_model_b200 import train as TR

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
NCF, T = 6, 1500
from camels_diffusion_model_b200 import _lib as L
from camels_diffusion_model_b200.parallel import PeerExchange
# ---- the fused reduce + cross-rank exchange kernel on its own: rank-dependent partials, 50 back-to-back exchanges
px = PeerExchange(dev) if world > 1 else None
gx = torch.Generator().manual_seed(123)
parts = [torch.randn(37, 300, generator=gx) for _ in range(world)]   # every rank knows every rank's partials
out = torch.empty(300, device=dev)
ok = True
for it in range(50):
    mine = (parts[rank] * (it + 1)).to(dev)
    L.xrank_sum(mine, out, xr=px.args if px else None)
    expect = torch.zeros(300, device=dev)
    for r in range(world):  # rank order, each rank's row-sum computed by the same kernel path (world = 1)
        loc = torch.empty(300, device=dev)
        L.xrank_sum((parts[r] * (it + 1)).to(dev), loc)
        expect += loc
    ok = ok and torch.equal(out, expect)
gath = [torch.empty_like(out) for _ in range(world)]
dist.all_gather(gath, out)
same = all(torch.equal(gath[0], g_) for g_ in gath)
if rank == 0:
    print(f"XRANK-SUM world={world}: 50 exchanges exact vs rank-ordered sum: {ok}; bit-identical on all ranks: {same}")
del px
torch.manual_seed(0)
ref_model = cdm.ContextUnet(1, 128, NCF, 64)
g = torch.Generator().manual_seed(1)
for k, v in ref_model.state_dict().items():            # non-trivial norm layers
    if k.endswith(".1.weight") and v.dim() == 1: v.uniform_(0.5, 1.5, generator=g)
    if k.endswith(".1.bias") and v.dim() == 1: v.normal_(0, 0.2, generator=g)
sd = {k: v.clone() for k, v in ref_model.state_dict().items()}
b_t, a_t, ab_t = cdm.make_schedule(T, device=dev)
B = 4 * world
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

s, e = rank * 4, rank * 4 + 4
m_1, loss_1 = step(x, prm, noise, t, False)
for mode in ("nccl", "peer"):
    TR.PEER_EXCHANGE = mode == "peer"
    m_dp, loss_dp = step(x[s:e], prm[s:e], noise[s:e], t[s:e], True)
    worst, key = 0.0, {}
    for (n1, p1), (n2, p2) in zip(m_1.named_parameters(), m_dp.named_parameters()):
        if float(p1.grad.norm()) < 1e-7: continue
        err = float((p1.grad - p2.grad).norm() / p1.grad.norm())
        worst = max(worst, err)
        if n1 in ("out.3.weight", "out.0.weight", "out.1.weight", "up2.model.2.conv2.1.weight", "up2.model.2.conv2.0.weight"):
            key[n1] = f"{err:.2e}"
    bn_err = max(float((b1 - b2).abs().max()) for (k, b1), (_, b2) in zip(m_1.named_buffers(), m_dp.named_buffers())
                 if "running" in k)
    lt = torch.tensor([loss_dp], device=dev)
    dist.all_reduce(lt)
    if rank == 0:
        # random-init train-mode BatchNorm stacks amplify ANY perturbation ~1.3x per layer (gradient explosion at
        # init), so only the layers nearest the loss are a sharp check of the exchange logic; `worst` is informational
        print(f"DP-CHECK[{mode}] well-conditioned gradients (sharded vs single-process rel-L2):", key)
        print(f"DP-CHECK[{mode}] world={world}: worst grad rel-L2 (sharded vs single-process) {worst:.3e}; "
              f"max |running-stat diff| {bn_err:.3e}; mean rank loss {float(lt) / world:.6f} vs global {loss_1:.6f}")
TR.PEER_EXCHANGE = False

# ---- throughput of a training step, BASELINE config 3: global batch 256
TR.DATA_PARALLEL = True
GB = 256
per = GB // world
model = cdm.ContextUnet(1, 128, NCF, 64)
model.load_state_dict(sd)
model = model.to(dev).train()
opt = TR.FusedAdam(model.parameters(), lr=1e-5)
xb, pb = torch.rand(per, 1, 64, 64, generator=g).to(dev), torch.rand(per, NCF, generator=g).to(dev)
for _ in range(3):
    TR.training_step(model, opt, xb, pb, T, ab_t, shortcut=sc)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record()
for _ in range(K):
    loss = TR.training_step(model, opt, xb, pb, T, ab_t, shortcut=sc)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / K], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    tf = 3 * 19.1785e9 * GB / (float(ms) * 1e-3) / 1e12
    print(f"TRAIN-BENCH world={world} global_batch={GB}: {float(ms):.2f} ms/step, {GB / float(ms) * 1e3:.0f} img/s, "
          f"{tf:.0f} TFLOP/s aggregate (3x forward FLOPs), loss {float(loss):.4f}")
# ---- the same step captured as a CUDA graph, 32 and 128 images per GPU; BatchNorm statistics exchanged by NCCL
# all-reduces inside the capture ("nccl") or by the fused reduce + exchange kernel over NVLink peer memory ("peer")
for mode in ("nccl", "peer"):
    TR.PEER_EXCHANGE = mode == "peer"
    for per in (32, GB // world):
        try:
            torch.manual_seed(0)
            model = cdm.ContextUnet(1, 128, NCF, 64)
            model.load_state_dict(sd)
            model = model.to(dev).train()
            gs = TR.GraphedTrainStep(model, per, T, ab_t, lr=1e-5)
            g2 = torch.Generator().manual_seed(77 + rank)
            xb, pb = torch.rand(per, 1, 64, 64, generator=g2).to(dev), torch.rand(per, NCF, generator=g2).to(dev)
            tb = torch.randint(1, T + 1, (per,), generator=g2).to(dev)
            for _ in range(3): gs(xb, pb, t=tb, shortcut=sc)
            torch.cuda.synchronize(); dist.barrier()
            e0.record()
            for _ in range(K): loss = gs(xb, pb, t=tb, shortcut=sc)
            e1.record(); torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / K], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"TRAIN-BENCH-GRAPH[{mode}] world={world} per_gpu={per}: {float(ms):.2f} ms/step, "
                      f"{per * world / float(ms) * 1e3:.0f} img/s, loss {float(loss):.4f}")
            del gs, model
            torch.cuda.empty_cache()
        except Exception as ex:  # noqa: BLE001
            if rank == 0:
                print(f"TRAIN-BENCH-GRAPH[{mode}] failed:", type(ex).__name__, str(ex)[:300])
            break
dist.destroy_process_group()
