"""Host-side logic of the multi-GPU path on CPU: world_size-2/3 gloo process groups exercise the
row all-gather (uneven shares), the edge-count all-gather and the shard arithmetic. The kernels
themselves never run on the CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scema_b200.distributed import gather_counts, gather_rows, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full_ref = torch.arange(n * k, dtype=torch.float64).reshape(n, k) * 0.5
        b, e = shard_bounds(n, world)[rank]
        full = gather_rows(full_ref[b:e].clone(), n, world)
        assert torch.equal(full, full_ref)
        counts, offs, flag = gather_counts(10 * rank + 3, torch.device("cpu"), flag=int(rank == world - 1))
        assert counts == [10 * r + 3 for r in range(world)]
        assert offs == [sum(counts[:r]) for r in range(world)]
        assert flag == 1 and gather_counts(1, torch.device("cpu"))[2] == 0
        # one shard reports "survivors too dense" (SCEMA_ERR_DENSE) for the tcgen05 filter: every rank repeats with the
        # DMMA filter, and, when that is too dense on another rank, with the filter-free kernel
        from scema_b200.binding import ScemaError
        from scema_b200.distributed import ShardedCluster

        class FakeContext:
            calls = []

            def compare(self, thr, variant, shard=0, n_shards=1):
                self.calls.append(variant)
                if (variant == 3 and shard == 1) or (variant == 0 and shard == 0):
                    raise ScemaError(7, "dense")
                return 100 * variant + shard

        fake = FakeContext()
        sc = ShardedCluster(fake)
        ne, counts, offs = sc.compare_all_ranks(1e-6, 3, None, torch.device("cpu"))
        assert fake.calls == [3, 0, 2] and sc.variant_used == 2
        assert ne == 200 + rank and counts == [200 + r for r in range(world)]
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _run(world, n, k, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, k, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_shard_bounds():
    for n in (0, 1, 7, 1000, 1000003):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def test_gather_even_world2(tmp_path):
    _run(2, 64, 6, tmp_path)


def test_gather_uneven_world3(tmp_path):
    _run(3, 101, 12, tmp_path)


def test_tile_shards_partition_the_groups():
    """Every strip group of the K2 schedule belongs to exactly one shard (same arithmetic as
    compare_run in pairs.cu: group g -> shard g % n_shards)."""
    for groups in (1, 5, 97, 1000):
        for w in (1, 2, 3, 8):
            owned = []
            for s in range(w):
                local = (groups - s + w - 1) // w if groups > s else 0
                owned += [lg * w + s for lg in range(local)]
            assert sorted(owned) == list(range(groups))


def test_aligned_shard_bounds():
    """Equal shares of whole 2048-row panels (what the overlapped exchange all-gathers in place): contiguous, cover
    [0, n), every inner boundary a multiple of 2048, every share but the last ones full."""
    from scema_b200.distributed import aligned_shard_bounds
    for n in (1, 2047, 2048, 20011, 1000000, 4000000):
        for w in (1, 2, 3, 4, 8):
            per, b = aligned_shard_bounds(n, w)
            assert per % 2048 == 0 and per * w >= n
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert all(s % 2048 == 0 for s, _ in b if s < n)
            assert all(e - s == per for s, e in b if e < n)
    assert aligned_shard_bounds(1000000, 8)[1][0] == (0, 126976) and aligned_shard_bounds(1000000, 8)[1][-1] == (888832, 1000000)


def _edges_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        from scema_b200.distributed import gather_edges
        n = 1000
        rng = np.random.default_rng(7)
        a_all = rng.integers(0, n - 1, size=300).astype(np.uint32)
        b_all = (a_all + 1 + rng.integers(0, 5, size=300)).clip(max=n - 1).astype(np.uint32)
        keys = np.unique(a_all.astype(np.int64) * n + b_all)
        a_all, b_all = (keys // n).astype(np.uint32), (keys % n).astype(np.uint32)
        d_all = (a_all * 1e-9 + b_all * 1e-12).astype(np.float64)
        mine = np.arange(len(keys)) % world == rank             # interleaved shards, like the tile items
        counts = [int((np.arange(len(keys)) % world == r).sum()) for r in range(world)]
        A, B, D = gather_edges(a_all[mine], b_all[mine], d_all[mine], n, counts, torch.device("cpu"))
        assert np.array_equal(A, a_all) and np.array_equal(B, b_all) and np.array_equal(D.view(np.uint64), d_all.view(np.uint64))
        open(os.path.join(out_dir, f"edges_ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gather_edges_world3(tmp_path):
    """The union of the ranks' (interleaved) edge lists, gathered and put back in canonical order (what bench.py verifies)."""
    port = _free_port()
    mp.spawn(_edges_worker, args=(3, port, str(tmp_path)), nprocs=3, join=True)
    assert all(os.path.exists(tmp_path / f"edges_ok{r}") for r in range(3))
