"""Multi-GPU plumbing for the row-sharded render: one process per GPU, scene
replicated, rows interleaved over ranks, one all-gather of the float64 framebuffer per frame.
Pure torch.distributed (NCCL on GPUs, gloo in the CPU tests) - nothing here touches the kernels."""
import torch
import torch.distributed as dist


def interleaved_rows(rank, world, height):
    """Rows rank, rank+world, ... -> (y_first, y_step, n_rows) for mcrt_render_rows_strided_dev."""
    return rank, world, len(range(rank, height, world))


def max_rows(world, height):
    return len(range(0, height, world))


def gather_frame(local, height, world, out=None):
    """local: [max_rows(world,height), W, 3] rows of this rank (padding rows ignored).
    Returns the assembled frame [height, W, 3] on every rank. Row k*world + r comes from rank r."""
    if world == 1:
        return local[:height]
    rows, W, C = local.shape
    gathered = out if out is not None else torch.empty((world * rows, W, C), dtype=local.dtype, device=local.device)
    gathered = gathered.view(world * rows, W, C)   # concatenation layout (what gloo and NCCL both accept)
    dist.all_gather_into_tensor(gathered, local.contiguous())
    return gathered.view(world, rows, W, C).permute(1, 0, 2, 3).reshape(rows * world, W, C)[:height].contiguous()
