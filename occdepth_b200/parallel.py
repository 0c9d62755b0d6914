"""Multi-GPU plumbing for the forward path: one process per GPU, frames sharded across ranks.

Frames are independent: the data path needs NO collective -- every rank runs the whole
forward on its own frames.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only for the
barrier around a timed region, the max-over-ranks reduction of device times and for gathering predictions.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None, device=None):
    """Initialise the default process group from the torchrun environment (no-op for world size 1)."""
    world, rank, local = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, **kw)
    return world, rank, local


def frame_shard(n_frames, rank, world):
    """Contiguous, balanced shard [lo, hi) of n_frames for `rank` (first n_frames % world ranks get one extra)."""
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (device times must be reduced as the MAX over ranks)."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device if device is not None else torch.device("cpu"))
    if t.is_cuda:
        t = t.float()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(count, device=None):
    """sum of an integer over ranks (e.g. frames processed)"""
    if not dist.is_initialized():
        return int(count)
    t = torch.tensor([int(count)], dtype=torch.int64, device=device if device is not None else torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


# --------------------------------------------------------------------------------------------------
# X-slab partition of ONE frame across ranks (BASELINE.json configs[2]): the voxel grid is split along X (the
# camera-depth axis; SURVEY.md section 8e explains why not the literal Z), every 3-D activation carries `halo`
# margin planes, and the only data-path communication is the neighbour exchange of those planes before a
# convolution whose taps reach across the slab boundary, plus one all-gather of the CRP mega-context.
class HaloExchangeOp:
    """Plan op: fill the margins of a halo-carrying activation from the two X-neighbours (send/recv pairs)."""

    def __init__(self, ctx, cl):
        self.ctx, self.name, self.flops = ctx, "halo_exchange", 0
        buf, h, n = cl.buf, cl.halo, cl.dlen
        w = min(h, n)
        self.bytes = 2 * w * buf[0, 0].numel() * buf.element_size()
        self.send_left, self.recv_left = buf[0, h:h + w], buf[0, h - w:h]
        self.send_right, self.recv_right = buf[0, h + n - w:h + n], buf[0, h + n:h + n + w]
        self._keep = buf

    def run(self, stream=None):
        c = self.ctx
        ops = []
        if c.rank > 0:
            ops += [dist.P2POp(dist.isend, self.send_left, c.rank - 1), dist.P2POp(dist.irecv, self.recv_left, c.rank - 1)]
        if c.rank < c.world - 1:
            ops += [dist.P2POp(dist.isend, self.send_right, c.rank + 1),
                    dist.P2POp(dist.irecv, self.recv_right, c.rank + 1)]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()


class AllGatherOp:
    """Plan op: dst (full, contiguous) = concat over ranks of src (this rank's contiguous shard)."""

    def __init__(self, ctx, src, dst):
        self.ctx, self.src, self.dst, self.name, self.flops = ctx, src, dst, "all_gather", 0

    def run(self, stream=None):
        dist.all_gather_into_tensor(self.dst, self.src)


class SlabContext:
    def __init__(self, rank=None, world=None, halo=3):
        w, r, _ = env_world()
        self.rank = r if rank is None else rank
        self.world = w if world is None else world
        self.halo = halo
        self.n_exchanges = 0

    def exchange_op(self, cl):
        self.n_exchanges += 1
        return HaloExchangeOp(self, cl)

    def all_gather_op(self, src, dst):
        return AllGatherOp(self, src, dst)

    def slab(self, n):
        """planes [lo, hi) of an extent-n axis owned by this rank (n must divide evenly)"""
        if n % self.world:
            raise RuntimeError("slab partition: extent %d is not divisible by %d ranks" % (n, self.world))
        per = n // self.world
        return self.rank * per, (self.rank + 1) * per


# --------------------------------------------------------------------------------------------------
# Single-process stand-in for R ranks on ONE device (tests / debugging of the slab partition without R GPUs):
# the R plans are executed in lock-step and the exchange ops copy directly between the plans' buffers.
class SimSlabGroup:
    def __init__(self, world, halo=3):
        self.world, self.ex, self.ag = world, {}, {}
        self.ctxs = [SimSlabContext(self, r, world, halo) for r in range(world)]

    @staticmethod
    def run_lockstep(plans, stream=None):
        from . import _lib
        st = _lib.stream_ptr() if stream is None else stream
        n = len(plans[0].ops)
        assert all(len(p.ops) == n for p in plans), "ranks must build identical op sequences"
        for i in range(n):
            for p in plans:
                p.ops[i].run(st)


class _SimExchangeOp(HaloExchangeOp):
    def __init__(self, ctx, cl, idx):
        super().__init__(ctx, cl)
        self.idx = idx

    def run(self, stream=None):
        peers = self.ctx.group.ex[self.idx]
        r = self.ctx.rank
        if r > 0:
            self.recv_left.copy_(peers[r - 1].send_right)
        if r < self.ctx.world - 1:
            self.recv_right.copy_(peers[r + 1].send_left)


class _SimAllGatherOp(AllGatherOp):
    def __init__(self, ctx, src, dst, idx):
        super().__init__(ctx, src, dst)
        self.idx = idx

    def run(self, stream=None):
        peers = self.ctx.group.ag[self.idx]
        flat = self.dst.reshape(self.ctx.world, -1)
        for r in range(self.ctx.world):
            flat[r].copy_(peers[r].src.reshape(-1))


class SimSlabContext(SlabContext):
    def __init__(self, group, rank, world, halo):
        super().__init__(rank, world, halo)
        self.group, self.n_gathers = group, 0

    def exchange_op(self, cl):
        op = _SimExchangeOp(self, cl, self.n_exchanges)
        self.group.ex.setdefault(self.n_exchanges, {})[self.rank] = op
        self.n_exchanges += 1
        return op

    def all_gather_op(self, src, dst):
        op = _SimAllGatherOp(self, src, dst, self.n_gathers)
        self.group.ag.setdefault(self.n_gathers, {})[self.rank] = op
        self.n_gathers += 1
        return op
