"""Slab-partitioned OccDepth.forward over real ranks (torchrun, NCCL): parity vs the un-partitioned forward on the
same GPU + timing.   torchrun --nproc-per-node N tools/slab_check.py [--full]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from occdepth_b200 import parallel  # noqa: E402


def main():
    world, rank, local = parallel.env_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    parallel.init("nccl", dev)
    m = bench.build_model().to(dev)
    img, pix, fov = bench.make_inputs(seed=0)
    batch = {"img": img.to(dev), "projected_pix_2": [pix.to(dev)], "fov_mask_2": [fov.to(dev)]}
    with torch.no_grad():
        ref = m(batch)["ssc_logit"]
        ctx = parallel.SlabContext(halo=3)
        m.enable_slab_parallel(ctx)
        out = m(batch)
        torch.cuda.synchronize()
        lo, hi = ctx.slab(ref.shape[2])
        got = out["ssc_logit"]
        err = float((got - ref[:, :, lo:hi]).abs().max() / ref.abs().max())
        for _ in range(3):
            m(batch)
        parallel.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 5
        for _ in range(K):
            m(batch)
        e1.record()
        parallel.barrier()
        torch.cuda.synchronize()
        ms = parallel.max_over_ranks(e0.elapsed_time(e1) / K, dev)
    errs = [None] * world
    dist.all_gather_object(errs, err)
    if rank == 0:
        print("slab partition over %d ranks: local logits %s, max rel diff vs un-partitioned %.2e, %d halo exchanges, "
              "%.3f ms/frame (max over ranks) -> %.1f M voxels/s" % (world, tuple(got.shape), max(errs),
                                                                       ctx.n_exchanges, ms, bench.N_OUT / ms / 1e3))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
