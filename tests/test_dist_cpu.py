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
    # a rank without any projector gradient (text-only micro-batch) still issues the same fixed-size collective and
    # receives the mean (ranks must never disagree on the size of, or skip, the call)
    p2 = torch.nn.Linear(16, 8)
    if rank == 0:
        p2.weight.grad = torch.full_like(p2.weight, 4.0)
    hd.allreduce_projector_grads(p2)
    # the reducer: preallocated flat bucket, reduce_async right after "backward", wait deferred to the projector's next use
    p3 = torch.nn.Linear(16, 8)
    red = hd.ProjectorGradReducer(p3, use_sink=False)
    applied = []
    red.attach(on_reduced=lambda: applied.append(float(p3.weight.grad[0, 0])))
    p3.weight.grad = torch.full_like(p3.weight, float(2 * rank + 1))
    if rank == 1:
        p3.bias.grad = torch.full_like(p3.bias, 6.0)
    red.reduce_async()
    assert applied == [] and red.pending is not None
    p3._hvlm_pre_forward()                 # what arch._project calls in front of the projector GEMM
    assert red.pending is None
    red.close()
    q.put((rank, clips, float(proj.weight.grad[0, 0]), float(proj.bias.grad[0]), mx, float(p2.weight.grad[0, 0]),
           float(p2.bias.grad[0]), applied, float(p3.bias.grad[0]), red.bucket.numel()))
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
        assert r[5] == 2.0 and r[6] == 0.0                      # mean of (4, missing -> 0); bias: nobody had one
        assert r[7] == [2.0] and r[8] == 3.0 and r[9] == 16 * 8 + 8


def test_single_process_is_noop():
    import hvlm_b200
    from hvlm_b200 import dist as hd
    proj = torch.nn.Linear(4, 4)
    proj.weight.grad = torch.ones_like(proj.weight)
    assert hd.allreduce_projector_grads(proj) is None
    assert hd.shard_clips(5, 0, 1) == [0, 1, 2, 3, 4]
    assert hd.max_over_ranks(3.0) == 3.0
