"""CPU, world_size 2, gloo: host logic of the multi-GPU path -- shard arithmetic, the single
all-gather of packed (score, index) pairs, and that merging per-shard top-k lists reproduces the
global selection (descending, ties -> lowest global index).  The device merge itself
(bbo_topk_merge) is covered by the -m gpu tests."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_shard_range_partitions_everything():
    from bbo_b200.distributed import shard_range
    for q in (0, 1, 7, 8, 1000, 2 ** 20 + 3):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(q, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == q
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_roundtrip_bit_exact():
    from bbo_b200.distributed import pack_topk, unpack_topk
    s = torch.tensor([1.5, -0.0, float("-inf"), 3e-310, 7.25], dtype=torch.float64)
    i = torch.tensor([5, 0, -1, 2 ** 40, 9], dtype=torch.int64)
    s2, i2 = unpack_topk(pack_topk(s, i))
    assert torch.equal(s.view(torch.int64), s2.view(torch.int64)) and torch.equal(i, i2)


def _worker(rank, world, port, q_total, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bbo_b200.distributed import gather_topk, shard_range
        from oracle import gp_oracle as O
        scores = np.round(np.random.default_rng(123).standard_normal(q_total), 2)   # ties on purpose
        lo, hi = shard_range(q_total, rank, world)
        idx, top = O.top_k_fast(scores[lo:hi], k)             # what each GPU's kernel produces
        s = torch.full((k,), float("-inf"), dtype=torch.float64)
        i = torch.full((k,), -1, dtype=torch.int64)
        s[:len(top)] = torch.as_tensor(top)
        i[:len(idx)] = torch.as_tensor(idx + lo)              # global index = local + shard offset
        gs, gi = gather_topk(s, i)                            # the ONE collective
        assert gs.shape == (world * k,) and gi.shape == (world * k,)
        valid = gi >= 0
        order = np.lexsort((gi[valid].numpy(), -gs[valid].numpy()))[:k]
        merged = gi[valid].numpy()[order]
        ret[rank] = merged.tolist()
    finally:
        dist.destroy_process_group()


def test_gather_and_merge_world2():
    from oracle import gp_oracle as O
    q_total, k, world = 1001, 16, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, q_total, k, ret), nprocs=world, join=True)
    scores = np.round(np.random.default_rng(123).standard_normal(q_total), 2)
    ref, _ = O.top_k_fast(scores, k)
    assert ret[0] == ret[1] == ref.tolist()      # identical on every rank and equal to the 1-GPU result
