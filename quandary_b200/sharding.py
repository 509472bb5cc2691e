"""Initial-condition sharding and the single per-evaluation exchange.

The reference distributes initial conditions over `np_init` MPI ranks (src/main.cpp:133-175): rank r owns the contiguous
block iinit_global = r*ninit_local + iinit (src/optimproblem.cpp:248), ninit % np_init == 0 is required
(src/main.cpp:150-153), and every evaluation ends with 7 scalar MPI_Allreduce + one of length ndesign
(src/optimproblem.cpp:292-298,454-460,527).  Here: one rank per GPU, and ONE all-reduce (NCCL on GPUs, gloo in the CPU
tests) of the packed buffer [grad(ndesign) | obj_re, obj_im, fid_re, fid_im, penalty, penalty_dpdm, penalty_energy, pad].
"""
import numpy as np

NSCALARS = 8
SCALAR_NAMES = ("obj_cost_re", "obj_cost_im", "fidelity_re", "fidelity_im", "penalty", "penalty_dpdm", "penalty_energy")


def local_ic_range(ninit, rank, world):
    """[first, last) global initial-condition indices of `rank` (src/optimproblem.cpp:248)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    if world > ninit:
        raise ValueError(f"more ranks ({world}) than initial conditions ({ninit}): the reference caps np_init at ninit (src/main.cpp:145)")
    if ninit % world != 0:
        raise ValueError(f"ninit = {ninit} must be divisible by the number of ranks {world} (src/main.cpp:150-153)")
    nloc = ninit // world
    return rank * nloc, (rank + 1) * nloc


def pack(grad, scalars):
    buf = np.zeros(grad.size + NSCALARS)
    buf[: grad.size] = grad
    for i, k in enumerate(SCALAR_NAMES):
        buf[grad.size + i] = scalars.get(k, 0.0)
    return buf


def unpack(buf, ndesign):
    return buf[:ndesign].copy(), {k: float(buf[ndesign + i]) for i, k in enumerate(SCALAR_NAMES)}


def allreduce_(tensor, group=None):
    """In-place SUM over the initial-condition group (no-op without an initialised process group)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor
