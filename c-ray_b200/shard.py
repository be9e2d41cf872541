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


_index_cache = {}


def _pixel_indices(W, H, tile, world, device):
    """Per rank: the flat pixel indices (into fb.view(-1, 3)) of its tiles in pack() order, built once per geometry."""
    key = (W, H, tile, world, str(device))
    hit = _index_cache.get(key)
    if hit is None:
        cols = torch.arange(W, dtype=torch.int64)
        per_rank = []
        for r in range(world):
            parts = []
            for (x0, y0, x1, y1) in rank_rects(W, H, tile, r, world):
                rows = torch.arange(H - y1, H - y0, dtype=torch.int64)          # storage rows of the tile, top first
                parts.append((rows[:, None] * W + cols[None, x0:x1]).reshape(-1))
            per_rank.append((torch.cat(parts) if parts else torch.zeros(0, dtype=torch.int64)).to(device))
        hit = _index_cache[key] = per_rank
    return hit


def gather_to_rank0(fb, W, H, tile, rank, world, dist):
    """One gather of every rank's packed tiles to rank 0 (equal-size padded messages).  Packing and unpacking are one
    index_select / index_copy_ each (the tile grid of a 1080p frame is 510 tiles: copying them one by one costs hundreds
    of small launches on rank 0)."""
    if world == 1:
        return
    idx = _pixel_indices(W, H, tile, world, fb.device)
    pad = max(int(i.numel()) for i in idx) * 3
    px = fb.view(-1, 3)
    buf = torch.zeros(pad, device=fb.device, dtype=torch.float32)
    mine = idx[rank]
    if mine.numel():
        buf[:mine.numel() * 3] = px.index_select(0, mine).reshape(-1)
    outs = [torch.empty(pad, device=fb.device, dtype=torch.float32) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, outs, dst=0)
    if rank == 0:
        for r in range(1, world):
            n = int(idx[r].numel())
            if n:
                px.index_copy_(0, idx[r], outs[r][:n * 3].view(n, 3))
