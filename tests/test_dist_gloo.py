"""Frame sharding and the rect-list gather over a world_size-2 gloo group (CPU only)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from rectdetect_b200 import dist as rdist
from rectdetect_b200.api import RECT_DTYPE


def test_shard_ranges_cover_the_batch():
    for n in (0, 1, 7, 8, 512, 513):
        for w in (1, 2, 4, 8):
            spans = [rdist.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) <= -(-n // w) if n else True


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    lists = []
    for c in (0, 3, 1, 0, 5):
        r = np.zeros(c, RECT_DTYPE)
        r["c2"] = rng.random((c, 4, 2))
        r["status"] = rng.integers(0, 4, c)
        lists.append(r)
    first, back = rdist.unpack_rect_lists(rdist.pack_rect_lists(17, lists))
    assert first == 17 and len(back) == len(lists)
    assert all(a.tobytes() == b.tobytes() for a, b in zip(lists, back))
    first, back = rdist.unpack_rect_lists(rdist.pack_rect_lists(0, []))
    assert first == 0 and len(back) == 0 and list(back) == []
    # a RectLists (what the batch engine returns: one flat array + offsets) packs without touching the per-frame views
    from rectdetect_b200.api import RectLists
    rl = RectLists(np.concatenate(lists), [len(r) for r in lists])
    first, again = rdist.unpack_rect_lists(rdist.pack_rect_lists(3, rl))
    assert first == 3 and len(again) == len(lists) and again[-1].tobytes() == lists[-1].tobytes() and again[0].size == 0
    assert [a.tobytes() for a in again[1:3]] == [b.tobytes() for b in lists[1:3]]


def _fake_rects(frame):
    r = np.zeros(frame % 4, RECT_DTYPE)
    r["value"] = frame
    r["status"] = np.arange(frame % 4)
    return r


def _worker(rank, world, port, nframes, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = rdist.shard_range(nframes, world, rank)
    full = rdist.gather_rect_lists(lo, [_fake_rects(f) for f in range(lo, hi)], nframes)
    if rank == 0:
        ok = len(full) == nframes and all(full[f].tobytes() == _fake_rects(f).tobytes() for f in range(nframes))
        q.put(ok)
    else:
        q.put(full is None)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 11, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert res == [True, True]
