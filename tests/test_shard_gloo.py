"""CPU: the N>1 host logic (block partition of the ptx batch + gather of proof records) under world_size 2 / 3 gloo."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from taiga_b200 import shard


def test_shard_range_partitions():
    for n in (1, 2, 7, 64, 1024):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard.shard_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    lc, lv = 4480, 4448
    cp = [bytes([i]) * lc for i in range(4)]
    vp = [bytes([100 + i]) * lv for i in range(8)]
    rec = shard.pack_records(cp, vp, lc, lv)
    assert rec.shape == (2, 2 * lc + 4 * lv)
    c1, v1 = shard.unpack_record(rec[1], lc, lv)
    assert c1 == cp[2:4] and v1 == vp[4:8]


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(n_total, world, rank)
    local = np.zeros((hi - lo, 96), np.uint8)
    for i in range(lo, hi):
        local[i - lo] = (i * 7 + np.arange(96)) % 251   # a deterministic "proof record" per ptx
    full = shard.gather_records(local, n_total)
    q.put((rank, full.tobytes()))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_gather_records_gloo():
    for world, n_total in ((2, 5), (3, 7)):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
        for p in procs:
            p.start()
        results = [q.get(timeout=120) for _ in range(world)]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        expect = np.stack([(i * 7 + np.arange(96)) % 251 for i in range(n_total)]).astype(np.uint8).tobytes()
        for _, blob in results:
            assert blob == expect
