"""Multi-GPU plumbing on the CPU (gloo, world_size 2): image sharding and the detection gather."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import REPO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeBox:
    def __init__(self, row):
        self.xmin, self.ymin, self.xmax, self.ymax, self.confidence = row


class _FakeApp:
    """Stands in for InferenceCascadeApp: image j yields j % 3 deterministic boxes."""

    def run_inference_on_images(self, images, merge=True):
        out = []
        for j in images:
            out.append([_FakeBox((j, 10 * j + k, j + 5, 10 * j + k + 7, float(k + 1))) for k in range(j % 3)])
        return out


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    from rapidobjectdetectionusingcascadedcnns_b200 import dist as rdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    images = list(range(11))
    assert rdist.shard_indices(11, rank, world) == list(range(rank, 11, world))
    res = rdist.run_sharded(_FakeApp(), images)
    empty = rdist.gather_detections(np.zeros((0, 6)))           # ragged: nothing anywhere
    q.put((rank, res, empty.shape))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_run_gathers_all_detections_in_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [[(j, 10 * j + k, j + 5, 10 * j + k + 7, float(k + 1)) for k in range(j % 3)] for j in range(11)]
    for rank, res, eshape in got:
        assert eshape == (0, 6)
        assert [[tuple(float(v) for v in b) for b in per] for per in res] == \
               [[tuple(float(v) for v in b) for b in per] for per in expect]


class _FailingApp(_FakeApp):
    """Rank 1's shard has no usable image: the app raises like `app/abstract_inference_app.py:106-107`."""

    def __init__(self, rank):
        self.rank = rank

    def run_inference_on_images(self, images, merge=True):
        if self.rank == 1:
            raise ValueError("Could not extract any windows from the given image")
        return super().run_inference_on_images(images, merge)


class _ManyBoxesApp(_FakeApp):
    """More detections than the initial exchange capacity: the block must grow on every rank alike."""

    def run_inference_on_images(self, images, merge=True):
        return [[_FakeBox((j, k, j + 1, k + 1, 1.0)) for k in range(3000)] for j in images]


def _worker_edge(rank, world, port, q):
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    from rapidobjectdetectionusingcascadedcnns_b200 import dist as rdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rdist.run_sharded(_FailingApp(rank), list(range(6)))
        failed = None
    except rdist.ShardError as e:
        failed = str(e)
    rows = rdist.run_sharded(_ManyBoxesApp(), list(range(4)), as_rows=True)     # 2 x 3000 rows per rank > 4096
    q.put((rank, failed, rows.shape, float(rows[:, 0].min()), float(rows[:, 0].max()),
           bool((np.diff(rows[:, 0]) >= 0).all())))
    dist.barrier()
    dist.destroy_process_group()


def test_failing_rank_raises_everywhere_and_capacity_grows():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_edge, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, failed, shape, lo, hi, sorted_ok in got:
        assert failed is not None and "rank(s) [1]" in failed          # nobody hangs, everybody learns who failed
        assert shape == (12000, 6) and (lo, hi) == (0.0, 3.0) and sorted_ok


def test_bind_rank_to_cores_slices_the_affinity_set():
    from rapidobjectdetectionusingcascadedcnns_b200 import dist as rdist
    if not hasattr(os, "sched_getaffinity"):
        return
    before = os.sched_getaffinity(0)
    try:
        got = rdist.bind_rank_to_cores(1, 2)
        if len(before) >= 4:
            avail = sorted(before)
            per = len(avail) // 2
            assert got == avail[per:2 * per] and os.sched_getaffinity(0) == set(got)
        else:
            assert got is None
    finally:
        os.sched_setaffinity(0, before)


def test_single_process_is_passthrough():
    from rapidobjectdetectionusingcascadedcnns_b200 import dist as rdist
    rows = np.arange(12, dtype=np.float64).reshape(2, 6)
    assert rdist.gather_detections(rows) is rows
    assert rdist.run_sharded(_FakeApp(), [0, 1, 2])[2][1][4] == 2.0
