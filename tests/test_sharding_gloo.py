"""World-size-2 `gloo` tests (CPU) of the multi-GPU host logic: batch sharding like DataParallel.scatter, output
gathering like DataParallel.gather (SBL/train.py:114-115), max-over-ranks timing.  The per-shard "encoder" here is
the CPU oracle's encoder stack on small inputs: shards run independently and the gathered result must equal the
single-process result exactly (clips are independent in eval mode)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sbl_for_multilingual_lip_reading_b200 import sharding, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _encode(x, sd):
    from oracle import visual_encoder_oracle as O
    with torch.no_grad():
        return O.encoder_forward(x, [x.shape[1]] * x.shape[0], sd, n_layers=1)[0]


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        sd = synth.encoder_state_dict(2, 1)
        x = torch.randn(n, 6, 512, generator=torch.Generator().manual_seed(5))
        b, e = sharding.shard_bounds(n, world, rank)
        local = _encode(x[b:e], sd) if e > b else x.new_zeros((0, 6, 512))
        full = sharding.gather_outputs(local, n)
        slowest = sharding.max_over_ranks(10.0 + rank, "cpu")
        q.put((rank, (b, e), full, slowest))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [4, 5, 1])
def test_two_rank_shard_and_gather_equals_single_process(n):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = torch.randn(n, 6, 512, generator=torch.Generator().manual_seed(5))
    ref = _encode(x, synth.encoder_state_dict(2, 1))
    bounds = [r[1] for r in res]
    assert bounds[0][0] == 0 and bounds[-1][1] == n and bounds[0][1] == bounds[1][0]
    for _, _, full, slowest in res:
        assert full.shape == ref.shape
        assert torch.allclose(full, ref, atol=1e-5, rtol=1e-5)
        assert slowest == 11.0


def test_shard_bounds_match_tensor_chunk():
    for n in range(0, 20):
        for world in (1, 2, 3, 4, 8):
            chunks = list(torch.arange(n).chunk(world)) if n else []
            for rank in range(world):
                b, e = sharding.shard_bounds(n, world, rank)
                want = chunks[rank].tolist() if rank < len(chunks) else []
                assert list(range(b, e)) == want, (n, world, rank)
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)
