"""CPU, world_size 2, gloo: the host logic of the batch-sharded multi-GPU path (quantize_b200/dist.py).

The compute on each rank is the fake-quant restatement of the reference layer (host.QuantConv2d._forward, CPU torch) —
the engine itself has no CPU path; what is under test is sharding, gathering and the max-over-ranks timing reduction."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from quantize_b200 import dist as qdist


def test_shard_range_partitions_every_batch():
    for n in (0, 1, 7, 8, 256, 257):
        for w in (1, 2, 3, 8):
            spans = [qdist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, batch, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(1)
    from quantize_b200 import host, models
    qdist.init_from_env("gloo")
    model = models.build_quantized("resnet20", 8, 8, seed=0)           # replicated weights: same seed on every rank
    x = models.synthetic_batch("resnet20", batch, seed=1)              # the global batch, identical on every rank
    host.calibrate(model, x[:4])                                        # same calibration data on every rank
    with torch.no_grad():
        local = model(qdist.shard_batch(x))                             # each rank: its own images only
        gathered = qdist.gather_batch(local, batch)
        full = model(x)
    slowest = qdist.max_over_ranks(10.0 + rank)
    qdist.barrier()
    q.put((rank, tuple(local.shape), bool(torch.allclose(gathered, full, rtol=1e-5, atol=1e-6)),
           bool(torch.equal(gathered.argmax(1), full.argmax(1))), slowest))
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])
def test_two_rank_batch_sharding_gloo(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sizes = [r[1][0] for r in results]
    assert sum(sizes) == batch and abs(sizes[0] - sizes[1]) <= 1
    assert all(r[2] and r[3] for r in results)          # gathered shards == full-batch result, same top-1
    assert all(r[4] == 11.0 for r in results)           # max over ranks
