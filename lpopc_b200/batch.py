"""Batch sharding of independent OCP instances over the GPUs of one box.

A batch of MPC-style instances (same problem, mesh, tables and sparsity pattern; different
initial states) is the only part of the path that shards: contiguous blocks of ceil(B/G)
instances per rank, everything else replicated, no collective on the evaluation path.
`torch.distributed` (NCCL over NVLink on the GPUs, gloo in the CPU tests) is used only for
(i) agreeing on the shared mesh (broadcast from rank 0, a few KB) and (ii) gathering the
per-instance results (objective / status), 8-16 B per instance.
"""
import numpy as np


def shard_range(nbatch, rank, world):
    """Contiguous block of instances owned by `rank`: [lo, hi)."""
    per = (nbatch + world - 1) // world
    lo = min(nbatch, rank * per)
    return lo, min(nbatch, lo + per)


def broadcast_mesh(meshpoints, nodes, dist=None, device="cpu"):
    """Rank 0's mesh wins (setup broadcast).  Returns (meshpoints, nodes) as numpy arrays."""
    import torch
    mp = np.asarray(meshpoints, dtype=np.float64)
    nd = np.asarray(nodes, dtype=np.int64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return mp, nd
    hdr = torch.tensor([mp.size, nd.size], dtype=torch.int64, device=device)
    dist.broadcast(hdr, 0)
    k1, k2 = int(hdr[0]), int(hdr[1])
    tm = torch.zeros(k1, dtype=torch.float64, device=device)
    tn = torch.zeros(k2, dtype=torch.int64, device=device)
    if dist.get_rank() == 0:
        tm.copy_(torch.from_numpy(mp))
        tn.copy_(torch.from_numpy(nd))
    dist.broadcast(tm, 0)
    dist.broadcast(tn, 0)
    return tm.cpu().numpy(), tn.cpu().numpy()


def gather_results(local, nbatch, dist=None, device="cpu"):
    """All-gather of a per-instance result vector (padded contiguous shards) -> [nbatch]."""
    import torch
    local = np.asarray(local, dtype=np.float64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local.copy()
    world = dist.get_world_size()
    per = (nbatch + world - 1) // world
    buf = torch.zeros(per, dtype=torch.float64, device=device)
    buf[: local.size] = torch.from_numpy(local).to(device)
    out = [torch.zeros(per, dtype=torch.float64, device=device) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat(out).cpu().numpy()[:nbatch]


def mpc_starting_points(op, lgr_points, x0s):
    """[B, n] starting points for MPC instances that differ in the initial state: the problem's
    two-point guess with the initial-state offset ramped linearly to zero over the horizon."""
    ph = op.phases[0]
    N, ns = ph.GetTotalNodes(), len(ph.statemin)
    base = op.guess(lgr_points)
    tau = np.concatenate([np.asarray(lgr_points[0]), [1.0]])
    ramp = 0.5 * (1.0 - tau)
    x0s = np.asarray(x0s, dtype=np.float64)
    X = np.tile(base, (len(x0s), 1))
    nominal = np.array([ph.stateguess[j][0] for j in range(ns)])
    for j in range(ns):
        X[:, j * (N + 1):(j + 1) * (N + 1)] += np.outer(x0s[:, j] - nominal[j], ramp)
    return X


def mpc_bounds(xl, xu, op, x0s):
    """Per-instance variable bounds [B, n] (torch tensors like xl/xu): the initial state of instance b
    is fixed to x0s[b] through the state0 bounds, exactly like examples.quadrotor(x0=...) does for one."""
    import torch
    ph = op.phases[0]
    N, ns = ph.GetTotalNodes(), len(ph.statemin)
    B = len(x0s)
    XL, XU = xl.repeat(B, 1), xu.repeat(B, 1)
    idx = (torch.arange(ns) * (N + 1)).to(XL.device)
    x0t = torch.as_tensor(np.asarray(x0s), dtype=torch.float64).to(XL.device)
    XL[:, idx] = x0t
    XU[:, idx] = x0t
    return XL, XU
