"""CPU, world_size 2, gloo: the weight broadcast and utterance sharding used by bench.py --gpus N."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from facodec_b200 import distributed as D
    sds = None
    if rank == 0:
        g = torch.Generator().manual_seed(5)
        sds = {"encoder": {"a.weight_v": torch.randn(4, 3, 7, generator=g), "a.bias": torch.randn(4, generator=g)},
               "decoder": {"m.alpha": torch.randn(1, 6, 1, generator=g)}}
    out = D.broadcast_state_dicts(sds, src=0)
    csum = float(sum(v.double().sum() for m in out.values() for v in m.values()))
    lo, hi = D.shard_range(33, rank, world)
    mx = D.max_over_ranks(float(rank + 1))
    q.put((rank, csum, sorted(out.keys()), tuple(out["encoder"]["a.weight_v"].shape), lo, hi, mx))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]                       # identical weights on both ranks
    assert res[0][2] == ["decoder", "encoder"] and res[0][3] == (4, 3, 7)
    assert (res[0][4], res[0][5]) == (0, 17) and (res[1][4], res[1][5]) == (17, 33)
    assert res[0][6] == res[1][6] == 2.0
