"""Multi-GPU host plumbing: one process per GPU, every rank scans its bp-balanced shard of the padded
genome coordinate (`MerPCR(shard=(rank, world))`) with NO collective on the scan path; the only exchange is the
final hit gather to rank 0, which merges the per-rank sorted hit lists by the reference's order key
(engine.py:434: stable sort of discovery order by pos1) and writes the output.

    python -m merpcr_b200 --gpus 8 sts fa          # spawns 8 ranks (torch.distributed.run) on this node
    torchrun --nproc-per-node 8 -m merpcr_b200 sts fa   # the same, launched by hand

torch.distributed is plumbing only (NCCL on GPU boxes, gloo in the CPU test tier)."""
from __future__ import annotations

import os
import socket
import subprocess
import sys
from typing import List, Optional, Tuple

import numpy as np

ORDER_KEY = ("contig", "pos1", "hash_off", "rec", "rank")   # include/merpcr_b200.h: mpcr_hit order


def env_world() -> Tuple[int, int, int]:
    """(rank, world, local_rank) as torchrun exports them; (0, 1, 0) outside a multi-process launch."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Join the process group torchrun described in the environment (no-op for a single process)."""
    rank, world, local = env_world()
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def active_world(shard: Tuple[int, int]) -> bool:
    """True when this engine's shard is one rank of an initialised process group of the same size."""
    if shard[1] <= 1:
        return False
    try:
        import torch.distributed as dist
    except ImportError:
        return False
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() == shard[1] and \
        dist.get_rank() == shard[0]


def merge_hits(per_rank: List[np.ndarray]) -> np.ndarray:
    """Per-rank hit lists (each already in order) -> one list in the reference's output order.  A position is owned by
    exactly one rank, so there are no duplicates to drop; shard boundaries can interleave by up to one primer length
    (pos1 = seed position - hash_offset), hence a real merge by the full order key."""
    parts = [h for h in per_rank if h is not None and h.size]
    if not parts:
        return per_rank[0][:0] if per_rank and per_rank[0] is not None else np.zeros(0)
    merged = np.concatenate(parts)
    if len(parts) > 1:
        merged = merged[np.lexsort(tuple(merged[k] for k in reversed(ORDER_KEY)))]
    return merged


def gather_hits(hits: np.ndarray, dst: int = 0) -> Tuple[Optional[np.ndarray], int]:
    """The final hit gather: returns (merged hit list on rank `dst` / None elsewhere, total hit count on every rank)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
    raw = np.ascontiguousarray(hits).view(np.uint8).reshape(-1)
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[rank] = raw.size
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM)
    sizes = sizes.cpu().tolist()
    total = sum(sizes) // hits.dtype.itemsize
    # point-to-point copies of the raw records (a few MB at most): no pickling, nothing on the scan path
    merged = None
    if rank == dst:
        parts = []
        for r in range(world):
            if r == rank:
                parts.append(hits)
            elif sizes[r]:
                buf = torch.empty(sizes[r], dtype=torch.uint8, device=dev)
                dist.recv(buf, src=r)
                parts.append(buf.cpu().numpy().view(hits.dtype))
        merged = merge_hits(parts)
    elif raw.size:
        dist.send(torch.from_numpy(raw.copy()).to(dev), dst=dst)
    return merged, int(total)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def launch(n_gpus: int, argv: List[str]) -> int:
    """`python -m merpcr_b200 --gpus N ...`: re-run the command line as N ranks of one node; rank 0 writes the output."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_gpus}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), "-m", "merpcr_b200"] + list(argv)
    env = dict(os.environ)
    env.setdefault("OMP_NUM_THREADS", "1")
    return subprocess.call(cmd, env=env)
