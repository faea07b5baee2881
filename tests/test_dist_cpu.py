"""World-size-2 gloo tests (CPU) for the multi-GPU host logic: clip sharding and the projector-grad all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import hvlm_b200
    from hvlm_b200 import dist as hd
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    hd.init_process_group("gloo")
    clips = hd.shard_clips(7, rank, world)
    proj = torch.nn.Linear(16, 8)
    proj.weight.grad = torch.full_like(proj.weight, float(rank + 1))
    proj.bias.grad = torch.full_like(proj.bias, float(10 * (rank + 1)))
    fin = hd.allreduce_projector_grads(proj, async_op=(rank == 0) or True)
    fin()
    mx = hd.max_over_ranks(float(rank + 5))
    hd.barrier()
    q.put((rank, clips, float(proj.weight.grad[0, 0]), float(proj.bias.grad[0]), mx))
    dist.destroy_process_group()


def test_shard_and_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    for r in res:
        assert r[2] == 1.5 and r[3] == 15.0 and r[4] == 6.0


def test_single_process_is_noop():
    import hvlm_b200
    from hvlm_b200 import dist as hd
    proj = torch.nn.Linear(4, 4)
    proj.weight.grad = torch.ones_like(proj.weight)
    assert hd.allreduce_projector_grads(proj) is None
    assert hd.shard_clips(5, 0, 1) == [0, 1, 2, 3, 4]
    assert hd.max_over_ranks(3.0) == 3.0
