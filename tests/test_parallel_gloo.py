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
