"""The N>1 host logic (frame sharding, max-over-ranks timing reduction, frame counting) on CPU with the gloo
backend, world size 2 -- the same code path bench.py uses under torchrun with NCCL."""
import os
import socket

import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from occdepth_b200 import parallel
    w, r, _ = parallel.init(backend="gloo")
    lo, hi = parallel.frame_shard(5, r, w)
    parallel.barrier()
    mx = parallel.max_over_ranks(10.0 + r)         # rank 1 is "slower"
    total = parallel.gather_counts(hi - lo)
    q.put((r, lo, hi, mx, total))
    import torch.distributed as dist
    dist.destroy_process_group()


def test_frame_shard_is_a_partition():
    from occdepth_b200.parallel import frame_shard
    for n in (0, 1, 7, 8, 9):
        for world in (1, 2, 3, 8):
            cuts = [frame_shard(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1:3] == (0, 3) and res[1][1:3] == (3, 5)
    assert all(abs(r[3] - 11.0) < 1e-9 for r in res)      # max over ranks
    assert all(r[4] == 5 for r in res)


def _exchange_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    from occdepth_b200 import parallel
    from occdepth_b200.engine import CL
    parallel.init(backend="gloo")
    ctx = parallel.SlabContext(halo=3)
    n = 4
    cl = CL(torch.zeros(1, n + 6, 2, 3, 8, dtype=torch.bfloat16), 8, 0, 3, n)
    for p in range(n):      # plane p of rank r holds the value 10*r + p
        cl.buf[0, 3 + p] = 10 * rank + p
    ctx.exchange_op(cl).run()
    full = torch.zeros(world * n * 2 * 3 * 8, dtype=torch.bfloat16)
    ctx.all_gather_op(cl.interior().reshape(-1).contiguous(), full).run()
    q.put((rank, cl.buf[0, :, 0, 0, 0].float().tolist(), full.view(world, n, -1)[:, :, 0].float().tolist()))
    dist.destroy_process_group()


def test_halo_exchange_and_all_gather_gloo():
    """the slab partition's two communication ops (neighbour halo exchange, all-gather) on 3 CPU ranks"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 3
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # rank 1 (middle): left margin = rank 0's last 3 planes, right margin = rank 2's first 3 planes
    assert res[1][1] == [1, 2, 3, 10, 11, 12, 13, 20, 21, 22]
    assert res[0][1] == [0, 0, 0, 0, 1, 2, 3, 10, 11, 12]        # global boundary stays zero (conv padding)
    assert res[2][1] == [11, 12, 13, 20, 21, 22, 23, 0, 0, 0]
    for r in res:
        assert r[2] == [[0, 1, 2, 3], [10, 11, 12, 13], [20, 21, 22, 23]]
