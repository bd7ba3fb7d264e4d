"""Run under torchrun on >= 2 GPUs: tile-sharded resample + all-pairs, union of the ranks' edges
checked against the CPU oracle on rank 0.  torchrun --nproc-per-node N tools/multigpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scema_b200
from scema_b200 import synth
from scema_b200.distributed import ShardedCluster, aligned_shard_bounds, shard_bounds


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, P, thr = int(os.environ.get("CHECK_N", "20011")), 10, 1e-6
    b, e = shard_bounds(n, world)[rank]
    off = synth.offsets(6, e - b, 16, 5, 70, first=b)
    steps = synth.histories(6, e - b, 16, 5e-3, synth.default_pert(thr, P), off, first=b)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    hc = scema_b200.HistCluster(local, stream.cuda_stream)
    hc.set_histories(steps, off)
    sc = ShardedCluster(hc)
    # DMMA, FMA, the tcgen05 filter (the bench default) one-shot and streamed
    for variant, streamed in ((0, False), (1, False), (3, False), (3, True)):
        if streamed:
            got = []
            ne, counts, offs, full = sc.run(n, P, thr, variant, sink=lambda x, y, z: got.append((x, y, z)))
            a, bb, d = (np.concatenate([g[k] for g in got]) if got else np.zeros(0, dtype=(np.uint32, np.uint32, np.float64)[k])
                        for k in range(3))
        else:
            ne, counts, offs, full = sc.run(n, P, thr, variant)
            a, bb, d = hc.get_edges()
        assert ne == counts[rank] and len(a) == ne
        if variant == 3:
            assert hc.counters()["tc_slices"] == 1, hc.counters()   # the tcgen05 filter really ran on this shard
        # gather the edge lists on every rank (padded to the largest count)
        m = max(max(counts), 1)
        buf = torch.zeros((m, 3), dtype=torch.float64, device=dev)
        if ne:
            buf[:ne, 0] = torch.from_numpy(a.astype(np.float64)).to(dev)
            buf[:ne, 1] = torch.from_numpy(bb.astype(np.float64)).to(dev)
            buf[:ne, 2] = torch.from_numpy(d).to(dev)
        allb = torch.empty((world * m, 3), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allb, buf)
        if rank == 0:
            from oracle.pyoracle import Oracle
            o = Oracle()
            rows = full.cpu().numpy()
            off_all = synth.offsets(6, n, 16, 5, 70)
            st_all = synth.histories(6, n, 16, 5e-3, synth.default_pert(thr, P), off_all)
            want_rows = o.splinify_batch(st_all, off_all, P)
            assert np.array_equal(rows.view(np.uint64), want_rows.view(np.uint64)), "gathered spline matrix differs"
            wi, wj, wd, _ = o.all_pairs(want_rows, thr)
            globals().update(want_rows=want_rows, wi=wi, wj=wj, wd=wd)
            parts = [allb[r * m: r * m + counts[r]].cpu().numpy() for r in range(world)]
            got = np.concatenate(parts, axis=0)
            order = np.lexsort((got[:, 1], got[:, 0]))
            got = got[order]
            assert len(got) == len(wi), (len(got), len(wi))
            assert np.array_equal(got[:, 0].astype(np.uint32), wi) and np.array_equal(got[:, 1].astype(np.uint32), wj)
            assert np.array_equal(np.ascontiguousarray(got[:, 2]).view(np.uint64), wd.view(np.uint64))
            print(f"multigpu_check ok: world={world} variant={variant} streamed={streamed} edges={len(wi)} per-rank={counts}", flush=True)
    # the overlapped path (scema_tc_shard_*: own-row operand images all-gathered, FP64 rows on a side stream), twice on the
    # same context, then the plain path again on the same aligned shares
    per, ab = aligned_shard_bounds(n, world)
    b, e = ab[rank]
    off = synth.offsets(6, e - b, 16, 5, 70, first=b)
    steps = synth.histories(6, e - b, 16, 5e-3, synth.default_pert(thr, P), off, first=b)
    hc.set_histories(steps, off)
    sc2 = ShardedCluster(hc, side_group=dist.new_group(backend="nccl"), bounds=ab)
    for mode in ("overlapped", "overlapped", "plain"):
        if mode == "overlapped":
            ne, counts, offs, full = sc2.run_overlapped(n, P, thr)
            assert sc2.path == "overlapped", sc2.path
        else:
            ne, counts, offs, full = sc2.run(n, P, thr, 3)
        a, bb, d = hc.get_edges()
        assert ne == counts[rank] and len(a) == ne and hc.counters()["tc_slices"] == 1
        from scema_b200.distributed import gather_edges
        A, B, D = gather_edges(a, bb, d, n, counts, dev)
        if rank == 0:
            rows = full.cpu().numpy()
            assert np.array_equal(rows.view(np.uint64), want_rows.view(np.uint64)), "gathered spline matrix differs"
            assert len(A) == len(wi) and np.array_equal(A, wi) and np.array_equal(B, wj) and np.array_equal(D.view(np.uint64), wd.view(np.uint64))
            print(f"multigpu_check ok: world={world} tcgen05 {mode} path edges={len(wi)} per-rank={counts}", flush=True)
    hc.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
