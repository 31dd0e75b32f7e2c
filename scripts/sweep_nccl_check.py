"""torchrun --nproc-per-node N scripts/sweep_nccl_check.py: a small wavelength sweep sharded over N GPUs (NCCL
all_gather of the per-point n_eff) must reproduce the serial solves.  Used by tests/test_gpu.py when >= 2 GPUs exist."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
from common import strip_problem
from femwell_b200.basis import Basis
from femwell_b200.element import ElementTriP0
from femwell_b200.maxwell.waveguide import compute_modes
from femwell_b200.sweep import wavelength_sweep

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
m, eps = strip_problem(14)
b0 = Basis(m, ElementTriP0(), intorder=4)
wls = np.linspace(1.5, 1.6, 5)
eps_of = lambda wl: eps * (1.0 + 0.02 * (wl - 1.55))
table = wavelength_sweep(b0, eps_of, wls, num_modes=2, order=2)
if rank == 0:
    ref = np.stack([compute_modes(b0, eps_of(wl), wl, num_modes=2, order=2).n_effs for wl in wls])
    err = np.abs(table - ref).max()
    print(f"sweep over {world} ranks: max |n_eff - serial| = {err:.2e}")
    assert table.shape == (5, 2) and err < 1e-10, err
    print("SWEEP_NCCL_OK")
dist.barrier()
dist.destroy_process_group()
