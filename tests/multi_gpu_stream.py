"""Run under torchrun (one process per GPU) by tests/test_gpu_multi.py, or alone (world 1):
a short frame stream sharded over the ranks, rect lists gathered once; rank 0 saves them.

    python -m torch.distributed.run --nproc-per-node N tests/multi_gpu_stream.py OUT.npz N_FRAMES W H CASCADE
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import clfacedetection_b200 as clfd
    from clfacedetection_b200 import stream

    out, n_frames, W, H, name = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = clfd.Context(local_rank)
    cas = clfd.Cascade(os.path.join(ROOT, "data", "haarcascades", f"haarcascade_{name}.xml"))
    B = 16
    det = clfd.Detector(ctx, cas, W, H, max_batch=B, scale_factor=1.2)
    src = stream.StreamSource(W, H, n_canvases=2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    secs, rects = stream.timed_stream(det, src, n_frames, rank, world, B, barrier)
    if rank == 0:
        np.savez(out, rects=stream.sorted_rects(rects), world=world, seconds=secs)
    det.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
