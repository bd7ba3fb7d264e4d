"""Static tile schedule of the tcgen05 filter (scema_b200/csrc/tc_sched.h), enumerated on the host: every tile of the
upper triangle inside the launch's row / column range is visited exactly once over all shards and units, for both
kernel flavours (one CTA per 128-row tile, CTA pair per 256-row tile), incl. the column panels of the host-buffer
pipeline and the row panels of the streamed compare."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_schedule_covers_every_tile_once(tmp_path):
    exe = str(tmp_path / "tc_sched_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "helpers", "tc_sched_check.cc")])
    r = subprocess.run([exe, "4000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout[-2000:]


def test_shared_memory_plan_fits_for_every_supported_row_width():
    """Host logic of the tcgen05 launch (scema_tc_plan): for every K the variant accepts, both kernel flavours and both
    slice counts, the A buffers plus a power-of-two B ring of at least two stages fit the 224 KB the kernel may use;
    single-chunk rows keep the compile-time plan (two A buffers at 32 KB, 192 KB in all)."""
    import ctypes as C
    import numpy as np
    from scema_b200 import binding
    L = binding.lib()
    plan = np.zeros(6, dtype=np.uint32)
    seen_wide = 0
    for K in range(1, 700):
        for slices in (1, 2):
            for cg in (1, 2):
                rc = L.scema_tc_plan(K, slices, cg, plan.ctypes.data)
                nc = (K + 4 + 63) // 64
                if nc > 10 or (nc > 1 and slices == 2):
                    assert rc != 0, (K, slices, cg)
                    continue
                assert rc == 0, (K, slices, cg)
                chunks, a_bytes, n_abuf, lg_nst, stage, used = (int(x) for x in plan)
                assert chunks == nc and n_abuf in (1, 2) and 1 <= lg_nst <= 3
                assert stage == (1 if slices == 1 else 2) * (256 // cg) * 128
                assert used == n_abuf * a_bytes + (stage << lg_nst) <= 224 * 1024
                assert a_bytes % 1024 == 0 and stage % 1024 == 0          # swizzle atoms stay 1024-byte aligned
                if nc == 1:
                    assert a_bytes == 32768 and n_abuf == 2 and used == 192 * 1024
                else:
                    assert a_bytes == nc * 16384
                    seen_wide += 1
    assert seen_wide > 1000
    assert L.scema_tc_plan(0, 1, 1, plan.ctypes.data) != 0 and L.scema_tc_plan(60, 3, 1, plan.ctypes.data) != 0


def test_pipeline_ranges_are_whole_panels():
    """Host logic of the host-buffer pipeline (scema_pipeline_plan): at most 16 ranges that tile [0, n), every inner
    boundary a whole panel (2048 rows, so column panels start on 256-row tiles), no empty range."""
    import ctypes as C
    import numpy as np
    from scema_b200 import binding
    L = binding.lib()
    b = np.zeros(32, dtype=np.uint64)
    k = C.c_uint32(0)
    for n in [4096, 4097, 6000, 21000, 65535, 65536, 100000, 131072, 999999, 1000000, 1048576, 4000000, 4000001, (1 << 32) - 2]:
        assert L.scema_pipeline_plan(n, b.ctypes.data, 32, C.byref(k)) == 0
        r = int(k.value)
        bounds = [int(x) for x in b[: r + 1]]
        assert 1 <= r <= 16 and bounds[0] == 0 and bounds[-1] == n
        assert all(x < y for x, y in zip(bounds, bounds[1:]))
        assert all(x % 2048 == 0 for x in bounds[:-1])
        if n >= 2 * 65536:
            assert r >= 2


def test_filter_choice_cost_model():
    """scema_tc_choose (host logic behind the up-front choice of the filter): counts of a FP64 sample of 8192 pairs
    (one / two slices with centred copies, one / two slices with raw copies, DMMA) -> 1 / 2 slices, DMMA (0) or the
    filter-free kernel (-1), centred or raw copies, and the queue entries expected for the choice."""
    from scema_b200 import binding
    P = 1000000 * 999999 // 2
    big = 1 << 40
    ch = binding.tc_choose
    assert ch(P, 60, (0, 0, 0, 0, 0), 8192, big) == (1, 1, 0)                          # nothing survives: hi slices, centred
    assert ch(P, 60, (2, 0, 8192, 8192, 0), 8192, big) == (1, 1, 0)                    # two hits are noise for sizing the queue
    c, cen, est = ch(P, 60, (8, 0, 8192, 8192, 0), 8192, big)                          # 0.1 % survive one slice: still cheapest
    assert (c, cen) == (1, 1) and abs(est - P * 8 / 8192) <= P // 10 ** 6
    assert ch(P, 60, (4000, 3, 8192, 8192, 0), 8192, big)[:2] == (2, 1)                # half survive one slice, few two
    assert ch(P, 60, (8192, 8192, 8192, 8192, 2), 8192, big)[0] == 0                   # only the FP64 band separates
    assert ch(P, 60, (8192, 8192, 8192, 8192, 8192), 8192, big)[0] == -1               # everything is a neighbour
    assert ch(P, 60, (4000, 300, 4000, 300, 0), 8192, 1 << 20)[0] == 0                 # no room for a queue: DMMA (no survivors)
    assert ch(P, 300, (4000, 4000, 4000, 4000, 0), 8192, big)[0] == 0                  # wide rows have no two-slice kernel
    assert ch(P, 3000, (0, 0, 0, 0, 0), 8192, big)[0] == 0                             # beyond 10 chunks: no tcgen05 filter at all
    assert ch(45, 60, (0, 0, 0, 0, 0), 8192, big)[:2] == (1, 1)
    assert ch(P, 60, (8192, 8192, 20, 10, 0), 8192, big)[:2] == (1, 0)                 # a centre far from most rows: raw copies
    assert ch(P, 60, (8192, 8192, 180, 10, 0), 8192, big)[:2] == (2, 0)                # ... and two slices when they pay
    assert ch(P, 60, (20, 10, 18, 10, 0), 8192, big)[:2] == (1, 1)                     # ... but only when clearly better
