"""Multi-GPU driver: one process per GPU, images sharded by global index, no collective inside a Gibbs
step; one gather of the final ids and scores per call.  torch.distributed is plumbing
(NCCL over NVLink on the B200 box, gloo in CPU tests)."""
from __future__ import annotations

import os
import random
from typing import Callable, List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when not launched by it."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend: str | None = None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous block of the global index for this rank; the first n % world ranks get one extra."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def batches_of(n_images: int, batch_size: int) -> List[range]:
    """run.py's DataLoader(shuffle=False, drop_last=True): consecutive full batches only (run.py:178)."""
    return [range(i, i + batch_size) for i in range(0, n_images - batch_size + 1, batch_size)]


def consume_order_rng(order: str, max_len: int, max_iter: int):
    """Advance the global RNGs exactly as one generate_caption call of the reference would
    (gen_utils.py:110-111 shuffle, :210 random), so that a rank skipping a batch it does not own still draws
    the same visiting orders for the batches it does own as a single process would."""
    if order == "shuffle":
        random.shuffle(list(range(max_len)))
    elif order == "random":
        for _ in range(max_iter * max_len):
            np.random.randint(0, max_len)


def run_sharded(n_samples: int, batches: Sequence[range], run_batch: Callable[[int, int, range], object],
                skip_batch: Callable[[int, int, range], None], rank: int, world: int):
    """Walk (sample, batch) in the reference's order (run.py:180-190).  Batches are dealt to ranks in
    contiguous blocks; `skip_batch` is called for batches owned by other ranks (RNG bookkeeping)."""
    mine = set(shard_range(len(batches), rank, world))
    out = {}
    for s in range(n_samples):
        for bi, idx in enumerate(batches):
            if bi in mine:
                out[(s, bi)] = run_batch(s, bi, idx)
            else:
                skip_batch(s, bi, idx)
    return out


def gather_ids_scores(ids: torch.Tensor, scores: torch.Tensor, world: int):
    """all_gather of the final token ids int32[B_local, L] and scores f32[B_local] (equal B_local on every
    rank) -> ([world*B_local, L], [world*B_local]) on every rank.  The only collective of a call."""
    if world == 1:
        return ids, scores
    ids = ids.contiguous()
    scores = scores.contiguous()
    out_i = torch.empty((world * ids.shape[0],) + tuple(ids.shape[1:]), dtype=ids.dtype, device=ids.device)
    out_s = torch.empty((world * scores.shape[0],) + tuple(scores.shape[1:]), dtype=scores.dtype, device=scores.device)
    dist.all_gather_into_tensor(out_i, ids)
    dist.all_gather_into_tensor(out_s, scores)
    return out_i, out_s


def gather_objects(obj, world: int):
    """Host-side gather of small Python results (caption strings per batch) to every rank, rank order."""
    if world == 1:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def max_over_ranks(value: float, world: int, device=None) -> float:
    if world == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
