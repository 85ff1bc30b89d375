"""BASELINE config 5 on N GPUs: all-timestep NLL + ELBO/BPD of a synthetic map set, sharded over the ranks
(`parallel.evaluate_sharded`: contiguous shards, no data-path collective, one scalar all-reduce per output).
    torchrun --nproc-per-node N tools/nll_sharded.py [n_maps] [timesteps] [batch]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import camels_diffusion_model_b200 as cdm
from camels_diffusion_model_b200 import diffusion as D, parallel as P

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n_maps = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
bs = int(sys.argv[3]) if len(sys.argv) > 3 else 512
torch.manual_seed(0)
model = cdm.ContextUnet(1, 128, 6, 64).to(dev).eval()
b_t, a_t, ab_t = D.make_schedule(T, device=dev)
g = torch.Generator().manual_seed(1)
maps, prm = torch.rand(n_maps, 1, 64, 64, generator=g), torch.rand(n_maps, 6, generator=g)
torch.manual_seed(2)  # identical shortcut draws on every rank
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.time()
# ONE sweep accumulates both weights over the same forwards (BASELINE config 5); sample_offset keys the noise by global map
nll, elbo, bpd = P.evaluate_sharded(
    lambda dl, sample_offset=0: D.calculate_likelihood_and_elbo(model, dl, T, dev, ab_t, b_t, a_t, seed=7,
                                                                sample_offset=sample_offset), maps, prm, bs, dev)
torch.cuda.synchronize()
t_nll = time.time() - t0
if rank == 0:
    fwd = n_maps * T
    print(f"NLL-SHARDED world={world}: {n_maps} maps x {T} timesteps in {t_nll:.2f} s = {fwd / t_nll:.0f} image-forwards/s, "
          f"{19.1785e9 * fwd / t_nll / 1e12:.0f} TFLOP/s aggregate; nll {nll:.4f} elbo {elbo:.5f} bpd {bpd:.3e}")
if world > 1:
    dist.destroy_process_group()
