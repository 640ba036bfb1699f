"""Multi-GPU plumbing for the row-sharded render: one process per GPU, scene
replicated, rows interleaved over ranks. The frame exchange is fused into the film resolve: every
rank owns a full-frame float3 buffer that its peers map through CUDA IPC (PeerFrames), and the resolve
kernel of each rank stores its rows into all of them over NVLink (mcrt_render_rows_strided_peers).
torch.distributed carries the 64-byte handles and the barriers (NCCL on GPUs, gloo in the CPU tests);
gather_frame is the plain all-gather form kept for hosts without peer access."""
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


def exchange_handles(handle, world, device=None):
    """All ranks' 64-byte IPC handles, in rank order (all_gather of a uint8 tensor)."""
    assert len(handle) == 64
    if world == 1:
        return [bytes(handle)]
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
    out = torch.empty((world, 64), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out.view(-1), mine)
    return [bytes(row.tolist()) for row in out.cpu()]


class PeerFrames:
    """One full-frame buffer [height, width, 3] per rank (float32 by default: the north star's float3
    framebuffer), each mapped into every other rank. `ptrs` = device pointers usable on this rank, own
    buffer first is NOT assumed: ptrs[r] is rank r's frame."""

    def __init__(self, integrator, rank, world, height, width, float32=True, device=None):
        self.integrator, self.rank, self.world = integrator, rank, world
        self.height, self.width, self.float32 = height, width, float32
        self.nbytes = height * width * 3 * (4 if float32 else 8)
        self.own, handle = integrator.frame_alloc(self.nbytes)
        handles = exchange_handles(handle, world, device)
        self.ptrs = [self.own if r == rank else integrator.frame_open(handles[r]) for r in range(world)]

    def tensor(self):
        """The rank's own frame as a torch tensor (no copy)."""
        import ctypes
        dtype = torch.float32 if self.float32 else torch.float64
        n = self.height * self.width * 3
        iface = {"shape": (n,), "typestr": "<f4" if self.float32 else "<f8", "data": (self.own, False), "version": 2}
        holder = type("FrameMem", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device=torch.device("cuda", self.integrator.device)).view(self.height, self.width, 3)

    def render(self, camera, sqrtspp=None):
        """Renders this rank's interleaved rows into every rank's frame. Call barrier() before reading."""
        y_first, y_step, n_rows = interleaved_rows(self.rank, self.world, self.height)
        return self.integrator.render_rows_strided_peers(camera, self.ptrs, y_first, y_step, n_rows, self.float32, sqrtspp)

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()

    def close(self):
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                self.integrator.frame_close(p)
        if self.world > 1:
            dist.barrier()   # nobody frees a buffer a peer still maps
        self.integrator.frame_free(self.own)
        self.ptrs = []


# ---- photon pass over several GPUs (photon-mapper.cpp:61-78)
def emission_range(total, rank, world):
    """Contiguous range of the emission index space [0, total) for `rank`: (first, count); remainder spread over the
    first ranks. The index space is light-major (light 0's emissions, then light 1's, ...), so a range may span lights."""
    base, rem = divmod(int(total), world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def device_view(ptr, n_elements, dtype, device):
    """torch view (no copy) of `n_elements` of `dtype` at raw device address `ptr`."""
    if n_elements == 0 or not ptr:
        return torch.empty(0, dtype=dtype, device=device)
    typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.uint8: "|u1"}[dtype]
    holder = type("DeviceMem", (), {"__cuda_array_interface__": {"shape": (int(n_elements),), "typestr": typestr, "data": (int(ptr), False), "version": 2}})()
    return torch.as_tensor(holder, device=device)


def all_gather_photons(mine, world, device=None):
    """Concatenation, in rank order, of every rank's flat photon array (lengths differ per rank): one all-gather of the
    lengths, one of the arrays padded to the longest."""
    if world == 1:
        return mine.contiguous()
    n = torch.tensor([mine.numel()], dtype=torch.int64, device=device)
    counts = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, n)
    counts = counts.cpu().tolist()
    longest = max(max(counts), 1)
    padded = torch.zeros(longest, dtype=mine.dtype, device=device)
    padded[:mine.numel()] = mine
    out = torch.empty(world * longest, dtype=mine.dtype, device=device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * longest: r * longest + counts[r]] for r in range(world)]).contiguous()


def render_filtered(integrator, camera, rank, world, device=None):
    """A frame through the camera's reconstruction filter (mcrt_set_film) over `world` GPUs: every rank splats its
    interleaved rows into whole-frame sums, the sums are all-reduced, every rank resolves. -> [H, W, 3] float64."""
    H, W = camera.height, camera.width
    rgb = torch.zeros((H, W, 3), dtype=torch.float64, device=device)
    wsum = torch.zeros((H, W), dtype=torch.float64, device=device)
    y_first, y_step, n_rows = interleaved_rows(rank, world, H)
    integrator.render_film_sums_strided_dev(camera, rgb.data_ptr(), wsum.data_ptr(), y_first, y_step, n_rows)
    if world > 1:
        dist.all_reduce(rgb)
        dist.all_reduce(wsum)
    out = torch.empty_like(rgb)
    torch.cuda.synchronize(device)
    integrator.film_resolve_dev(rgb.data_ptr(), wsum.data_ptr(), H * W, out.data_ptr())
    return out
