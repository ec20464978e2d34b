"""Multi-GPU host logic on CPU: world_size-2 (and 4) gloo groups run the row-range shard plan and
the cap all-gather, with the oracle standing in for the per-rank device compute.  Checks that the
sharded result (leaves, digests, cap) is bit-identical to the single-process commit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, log_n, ncols, rate_bits, cap_height, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vfhe_b200 as V
    from oracle import binding as B
    B.set_threads(1)
    cols = V.synthetic_columns(ncols, 1 << log_n, seed=77)
    plan = V.shard_plan(log_n, rate_bits, cap_height, rank, world)
    full = B.commit(cols, rate_bits, cap_height)      # oracle = stand-in for the device compute
    my_leaves = full["leaves"][plan.first_leaf: plan.first_leaf + plan.nleaves]
    # a rank only needs its own rows: Merkle over them with ncap subtree roots
    local_h = (plan.ncap.bit_length() - 1)
    digests, roots = B.merkle_new(my_leaves, local_h)
    cap = V.gather_cap(torch.from_numpy(roots.view(np.int64)), world).numpy().view(np.uint64)
    ok = (np.array_equal(cap, full["cap"])
          and np.array_equal(digests, full["digests"][plan.digest_offset: plan.digest_offset + plan.ndigests])
          and digests.shape[0] == plan.ndigests)
    gathered = [None] * world
    dist.all_gather_object(gathered, bool(ok))
    if rank == 0:
        q.put(all(gathered))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,log_n,ncols,rate_bits,cap_height",
                         [(2, 6, 9, 3, 4), (2, 5, 20, 1, 1), (4, 6, 5, 3, 4), (2, 4, 3, 3, 7)])
def test_sharded_commit_equals_single(world, log_n, ncols, rate_bits, cap_height):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, log_n, ncols, rate_bits, cap_height, q))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_plan_arithmetic():
    import sys
    sys.path.insert(0, ROOT)
    import vfhe_b200 as V
    m = 1 << 19
    plans = [V.shard_plan(16, 3, 4, r, 8) for r in range(8)]
    assert [p.first_leaf for p in plans] == [r * (m // 8) for r in range(8)]
    assert all(p.nleaves == 1 << 16 and p.ncap == 2 for p in plans)
    assert sum(p.ndigests for p in plans) == 2 * (m - 16)
    assert plans[3].digest_offset == 3 * 2 * (2 * (m // 16) - 2)
    one = V.shard_plan(16, 3, 4, 0, 1)
    assert one.nleaves == m and one.ncap == 16 and one.ndigests == 2 * (m - 16)
    for bad in [dict(world=3, rank=0), dict(world=16, rank=0), dict(world=2, rank=2)]:
        with pytest.raises(ValueError):
            V.shard_plan(16, 3, 4, bad["rank"], bad["world"])
    with pytest.raises(ValueError):
        V.shard_plan(16, 3, 1, 0, 4)     # more shards than cap subtrees


def _proof_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vfhe_b200 as V
    sp = V.ShardedProof(rank, world)                      # CPU tensors: gloo
    rng = np.random.default_rng(5)
    ncap, m, width, layers = 16, 1 << 10, 7, 6
    full_cap = rng.integers(0, 2**64, size=(ncap, 4), dtype=np.uint64)
    own = ncap // world
    cap = np.zeros_like(full_cap)
    cap[rank * own:(rank + 1) * own] = full_cap[rank * own:(rank + 1) * own]
    ok = np.array_equal(sp.complete_cap(cap), full_cap)
    qidx = rng.integers(0, m, size=28, dtype=np.uint64)
    all_rows = rng.integers(0, 2**64, size=(28, width), dtype=np.uint64)
    all_sibs = rng.integers(0, 2**64, size=(28, layers, 4), dtype=np.uint64)
    mine = sp.owned(qidx, m)
    ok = ok and np.array_equal(mine, (qidx // np.uint64(m // world)) == rank)
    rows, sibs = np.zeros_like(all_rows), np.zeros_like(all_sibs)
    rows[mine], sibs[mine] = all_rows[mine], all_sibs[mine]
    sp.collect([rows, sibs])
    ok = ok and np.array_equal(rows, all_rows) and np.array_equal(sibs, all_sibs)
    gathered = [None] * world
    dist.all_gather_object(gathered, bool(ok))
    if rank == 0:
        q.put(all(gathered))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_proof_host_glue(world):
    """ShardedProof over gloo: caps completed from the per-rank entries, query openings collected
    from their owners — the host side of a step proof whose batches are sharded by row range."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_proof_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
