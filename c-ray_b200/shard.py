"""Tile sharding and the framebuffer gather for multi-GPU rendering (one process per GPU).

The frame is cut into the reference's tile grid (quantizeImage, src/datatypes/tile.c:66-117); tile k of the
row-major grid goes to rank k % world — pixel-passes are independent and seeds depend only on
(pixel, pass, spp), so no data-path collective is needed while rendering.  At the end every rank packs its
fp32 tiles and ONE gather (torch.distributed: NCCL on GPUs, gloo in the CPU tests) brings them to rank 0,
which scatters them into its framebuffer.  Works on any (H, W, 3) float32 tensor view of the framebuffer
whose row 0 is the image top (row H-1-y, texture.c:24-28).
"""
import torch


def tiles_of(W, H, t):
    """Row-major tile rectangles (x0, y0, x1, y1), y up, end exclusive, ragged at the right/top edges."""
    return [(x, y, min(x + t, W), min(y + t, H)) for y in range(0, H, t) for x in range(0, W, t)]


def rank_rects(W, H, tile, rank, world):
    return tiles_of(W, H, tile)[rank::world]


def rect_view(fb, rect):
    x0, y0, x1, y1 = rect
    H = fb.shape[0]
    return fb[H - y1:H - y0, x0:x1]


def pack(fb, rects):
    if not rects:
        return fb.new_zeros(0)
    return torch.cat([rect_view(fb, r).reshape(-1) for r in rects])


def unpack_into(fb, rects, flat):
    off = 0
    for r in rects:
        x0, y0, x1, y1 = r
        n = (x1 - x0) * (y1 - y0) * 3
        rect_view(fb, r).copy_(flat[off:off + n].view(y1 - y0, x1 - x0, 3))
        off += n


def gather_to_rank0(fb, W, H, tile, rank, world, dist):
    """One gather of every rank's packed tiles to rank 0 (equal-size padded messages)."""
    if world == 1:
        return
    all_rects = [rank_rects(W, H, tile, r, world) for r in range(world)]
    sizes = [sum((x1 - x0) * (y1 - y0) * 3 for (x0, y0, x1, y1) in rs) for rs in all_rects]
    pad = max(sizes)
    buf = torch.zeros(pad, device=fb.device, dtype=torch.float32)
    mine = pack(fb, all_rects[rank])
    buf[:mine.numel()] = mine
    outs = [torch.empty(pad, device=fb.device, dtype=torch.float32) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, outs, dst=0)
    if rank == 0:
        for r in range(1, world):
            unpack_into(fb, all_rects[r], outs[r])
