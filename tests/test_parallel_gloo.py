"""CPU, world_size 2 over gloo: the batch-sharding arithmetic of the N>1 path.

The GPU forward cannot run here, so the per-image function is the CPU oracle at 128x128: sharding the
batch over two ranks and gathering must reproduce the single-process result bit for bit (images are
independent once ActNorm is primed), for even and ragged batches."""
import contextlib
import io
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cfen_vit_dehazing_b200 import parallel


def test_shard_arithmetic():
    assert parallel.shard_sizes(256, 8) == [32] * 8
    assert parallel.shard_sizes(5, 2) == [3, 2]
    assert parallel.shard_sizes(1, 4) == [1, 0, 0, 0]
    for batch in (0, 1, 5, 16, 257):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        import cfen_vit_dehazing_b200 as prod
        from oracle import cfen_oracle as orc
        opt = orc.default_opt(2, 128)
        torch.manual_seed(1)                       # replicated weights: same seed on every rank
        with contextlib.redirect_stdout(io.StringIO()):
            net = prod.define_G(opt, prod.common.default_conv)
        x0, x = orc.make_inputs(128, [1, batch])
        sd = orc.prime(net.state_dict(), x0)
        mine = parallel.shard_batch(x, world, rank)
        outs = orc.forward(sd, mine) if mine.shape[0] else [x.new_zeros(0, c, 128, 128) for c in (3, 1, 3)]
        full = parallel.gather_outputs(outs, batch)
        if rank == 0:
            ref = orc.forward(sd, x)
            q.put([float((a - b).abs().max()) for a, b in zip(full, ref)])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [2, 3])
def test_two_rank_sharding_equals_single_process(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    diffs = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert max(diffs) <= 1e-5, diffs
