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


def test_single_process_is_passthrough():
    from rapidobjectdetectionusingcascadedcnns_b200 import dist as rdist
    rows = np.arange(12, dtype=np.float64).reshape(2, 6)
    assert rdist.gather_detections(rows) is rows
    assert rdist.run_sharded(_FakeApp(), [0, 1, 2])[2][1][4] == 2.0
