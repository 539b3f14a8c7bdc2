"""N>1 host logic on CPU: two gloo ranks shard a batch of .basis files by image, each transcodes only its share, and
the union equals the single-process result (checked through per-image CRCs exchanged with all_gather -- the exchange
exists only in this test; the product path has no collective).  The per-image work is done by the ORACLE here
(test infrastructure): the CUDA path needs a GPU and is covered by tests/test_gpu_*.py."""
import ctypes
import os
import sys
import zlib

import numpy as np
import pytest

from basisu_rs_b200.shard import image_cost, plan_shards


def test_plan_is_a_partition_and_balanced():
    for world in (1, 2, 3, 4, 8):
        costs = [1.0] * 37
        shards = plan_shards(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(37))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
        assert all(i % world == r for r, s in enumerate(shards) for i in s)
    costs = [image_cost(262144, i % 2 == 1) for i in range(64)] + [image_cost(16, False)] * 5
    for world in (2, 4, 8):
        shards = plan_shards(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(len(costs)))
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) <= 1.02 * (sum(costs) / world) + max(costs)
        assert plan_shards(costs, world) == shards          # deterministic: every rank derives the same plan
    assert plan_shards([], 4) == [[], [], [], []]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    import torch
    import torch.distributed as dist
    import conftest
    from basis_writer import uastc_file
    from uastc_synth import random_blocks
    from basisu_rs_b200.shard import transcode_batch
    import etc1s_common as ec

    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    orc = ec.bind(ctypes.CDLL(str(conftest.build_oracle())))
    files = []
    for i in range(11):                                    # ragged sizes, deterministic contents
        bx, by = 3 + i, 2 + (i % 4)
        files.append(uastc_file(random_blocks(bx * by, seed=100 + i).tobytes(), bx, by))

    def oracle_read_to(target, buf):
        e, imgs = ec.oracle_read_to(orc, target, buf)
        assert e == 0
        return None, imgs

    mine = transcode_batch(files, 2, rank, world, read_to=oracle_read_to)          # target 2 = BC7
    crcs = torch.zeros(len(files), dtype=torch.int64)
    for i, imgs in mine.items():
        crcs[i] = zlib.crc32(b"".join(im[3] for im in imgs)) + 1
    gathered = [torch.zeros_like(crcs) for _ in range(world)]
    dist.all_gather(gathered, crcs)
    owners = torch.stack(gathered)
    assert int((owners != 0).sum(dim=0).min()) == 1 and int((owners != 0).sum(dim=0).max()) == 1     # every image exactly once
    merged = owners.sum(dim=0)
    want = torch.tensor([zlib.crc32(b"".join(im[3] for im in oracle_read_to(2, f)[1])) + 1 for f in files], dtype=torch.int64)
    ok = bool((merged == want).all())
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok, sorted(mine)))


@pytest.mark.timeout(180)
def test_two_gloo_ranks_cover_the_batch_exactly_once():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert sorted(res[0][2] + res[1][2]) == list(range(11))
