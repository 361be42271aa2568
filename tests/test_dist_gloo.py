"""world_size-2 gloo tests of the multi-GPU host logic (sharding by image index, RNG bookkeeping for skipped
batches, the single result gather).  No GPU involved."""
import os
import random
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conzic_b200 import dist as cdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, lr, w = cdist.init("gloo")
    assert (r, w) == (rank, world)
    n_images, bs, n_len, iters, samples = 13, 2, 6, 2, 2
    batches = cdist.batches_of(n_images, bs)            # drop_last: 6 batches
    random.seed(42); np.random.seed(42)
    orders = {}

    def run_batch(s, bi, idx):
        order = list(range(n_len)); random.shuffle(order)   # what shuffle_generation draws (gen_utils.py:110-111)
        orders[(s, bi)] = order
        return [i * 10 + s for i in idx]

    def skip_batch(s, bi, idx):
        cdist.consume_order_rng("shuffle", n_len, iters)

    mine = cdist.run_sharded(samples, batches, run_batch, skip_batch, rank, world)
    # the one collective: fixed-size ids / scores per rank
    ids = torch.full((3, 4), rank, dtype=torch.int32)
    sc = torch.full((3,), float(rank))
    gi, gs = cdist.gather_ids_scores(ids, sc, world)
    allres = cdist.gather_objects((mine, orders), world)
    mx = cdist.max_over_ranks(float(rank + 1), world, "cpu")
    if rank == 0:
        q.put((gi.tolist(), gs.tolist(), allres, mx))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gi, gs, allres, mx = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert gi == [[0] * 4] * 3 + [[1] * 4] * 3 and gs == [0.0] * 3 + [1.0] * 3 and mx == 2.0
    # single-process reference walk: same RNG, every batch run
    n_images, bs, n_len, samples = 13, 2, 6, 2
    batches = cdist.batches_of(n_images, bs)
    assert [list(b) for b in batches] == [[0, 1], [2, 3], [4, 5], [6, 7], [8, 9], [10, 11]]
    random.seed(42)
    want = {}
    for s in range(samples):
        for bi, idx in enumerate(batches):
            o = list(range(n_len)); random.shuffle(o)
            want[(s, bi)] = o
    got_orders, got_results = {}, {}
    for mine, orders in allres:
        got_orders.update(orders)
        got_results.update(mine)
    assert got_orders == want, "sharded ranks drew different visiting orders than one process would"
    assert sorted(got_results) == sorted(want)
    assert set(cdist.shard_range(6, 0, 2)) == {0, 1, 2} and set(cdist.shard_range(6, 1, 2)) == {3, 4, 5}
    assert [len(cdist.shard_range(7, r, 3)) for r in range(3)] == [3, 2, 2]
