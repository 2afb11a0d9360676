"""Multi-GPU partitioning of the path. One process per GPU, torch.distributed for the plumbing
(gloo on CPU boxes, nccl on GPUs).

* Frame-sharded sweeps (BASELINE.json configs[2]): independent frames, contiguous blocks per rank, NO collective on the
  data path (frame_block, render_block, gather_frame_values).
* One large frame split by screen tiles (north_star's optional mode): every rank owns a band of 16-pixel tile rows of
  both passes; the shadow-map bands are all-gathered between the passes — the one real exchange step of the path, since
  pass 2 looks up arbitrary light-space texels (IShader.h:107-129) — and the colour + depth bands after pass 2
  (tile_row_bands, exchange_bands, render_split_frame). Bands are contiguous byte ranges of row-major planes, so they
  are gathered in place: ncclAllGather when the bands are equal, one ncclBroadcast per band otherwise."""
import numpy as np


def frame_block(rank, world, total):
    """Contiguous block [first, first+count) of `total` frames owned by `rank`; the first `total % world` ranks get one more."""
    base, extra = divmod(total, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def gather_frame_values(local_values, total, rank, world, device="cpu"):
    """All ranks obtain the per-frame values (e.g. checksums) of the whole sweep in frame order.
    local_values: uint64/int64 array for this rank's block."""
    import torch
    import torch.distributed as dist

    first, count = frame_block(rank, world, total)
    assert len(local_values) == count
    if world == 1:
        return np.asarray(local_values, np.uint64).copy()
    width = max(frame_block(r, world, total)[1] for r in range(world))
    buf = torch.zeros(width, dtype=torch.int64, device=device)
    buf[:count] = torch.from_numpy(np.asarray(local_values, np.uint64).view(np.int64).copy()).to(device)
    out = [torch.zeros(width, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(out, buf)
    parts = []
    for r in range(world):
        c = frame_block(r, world, total)[1]
        parts.append(out[r][:c].cpu().numpy().view(np.uint64))
    return np.concatenate(parts)


def render_block(ctx, hana, scene_objs, shader, W, H, first, count, frames_per_turn, batch=32, enable_shadow=True):
    """Render frames [first, first+count) of the orbit sweep on this rank's GPU; returns their checksums (uint64)."""
    model, dtex, ntex = scene_objs
    sweep = ctx.sweep(W, H, min(batch, max(count, 1)))
    sums = []
    done = 0
    while done < count:
        n = min(batch, count - done)
        arr = hana.orbit_sweep_uniforms(W, H, first + done, n, frames_per_turn=frames_per_turn, enable_shadow=enable_shadow)
        sweep.render(model, shader, arr, dtex, ntex, n_frames=n)
        sums.append(sweep.checksums(n))
        done += n
    sweep.close()
    return np.concatenate(sums) if sums else np.zeros(0, np.uint64)


# ----------------------------------------------------------------------------- one frame split by screen tiles
TILE = 16


def tile_row_bands(height, world):
    """[(first tile row, tile rows)] per rank: contiguous, sizes differ by at most one, empty bands when world > rows."""
    rows = (height + TILE - 1) // TILE
    return [frame_block(r, world, rows) for r in range(world)]


def band_pixel_rows(band, height):
    """Pixel rows [y0, y1) of a band of tile rows (the last tile row may be partial)."""
    return min(band[0] * TILE, height), min((band[0] + band[1]) * TILE, height)


def exchange_bands(plane, row_bytes, bands_rows, rank, world):
    """All-gather IN PLACE: `plane` is a flat uint8 torch tensor (one row-major image plane, `row_bytes` per pixel row)
    in which this rank's rows bands_rows[rank] = (y0, y1) are valid; afterwards every rank holds every band.
    Equal non-empty bands: one all_gather_into_tensor (ncclAllGather) whose input is the rank's own slice of the output;
    ragged bands: one broadcast per band."""
    import torch.distributed as dist

    if world == 1:
        return
    sizes = [(y1 - y0) * row_bytes for y0, y1 in bands_rows]
    offs = [y0 * row_bytes for y0, _ in bands_rows]
    contiguous = all(offs[r] + sizes[r] == offs[r + 1] for r in range(world - 1))
    if contiguous and len(set(sizes)) == 1 and sizes[0] > 0:
        out = plane[offs[0]:offs[0] + sizes[0] * world]
        dist.all_gather_into_tensor(out, plane[offs[rank]:offs[rank] + sizes[rank]])
        return
    for r in range(world):
        if sizes[r]:
            dist.broadcast(plane[offs[r]:offs[r] + sizes[r]], src=r)


class _DevicePlane:
    """A raw device allocation as a __cuda_array_interface__ object, so torch can view it without a copy."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def device_plane_tensor(ptr, nbytes, device):
    import torch

    return torch.as_tensor(_DevicePlane(ptr, nbytes), device=device)


def render_split_frame(ctx, hana, sweep, scene_objs, shader, uniforms, rank, world, device, exchange_shadow=True):
    """One frame (the sweep's frame 0), split by tile rows over `world` GPUs. Every rank ends with the complete frame
    in its sweep. exchange_shadow=False renders the whole (cheap, depth-only) shadow pass on every GPU instead of
    exchanging its bands (SURVEY.md §8e "alternative without the first exchange").
    Returns {"shadow_bytes": bytes this rank received between the passes, "frame_bytes": ... after pass 2}."""
    model, dtex, ntex = scene_objs
    W, H = sweep.width, sweep.height
    bands = tile_row_bands(H, world)
    rows = [band_pixel_rows(b, H) for b in bands]
    mine = bands[rank]
    shadowed = bool(uniforms.enable_shadow)
    empty = mine[1] == 0
    # a rank without rows renders nothing: a band of count 0 means "all rows" to the C ABI, so give it a row past the end
    past = ((H + TILE - 1) // TILE, 1)
    sweep.set_bands(shadow=(past if empty else mine) if exchange_shadow else (0, 0), main=past if empty else mine)
    moved = {"shadow_bytes": 0, "frame_bytes": 0}
    if shadowed:
        sweep.render_pass(hana.PASS_SHADOW, model, shader, [uniforms], dtex, ntex)
        if exchange_shadow:
            ptr, pitch, stride = sweep.shadow_plane()
            ctx.sync()  # this rank's pass 1 is in HBM before NCCL reads it (NCCL runs on torch's streams, not the context's)
            plane = device_plane_tensor(ptr, stride, device)
            exchange_bands(plane, pitch, rows, rank, world)
            moved["shadow_bytes"] = sum((y1 - y0) * pitch for r, (y0, y1) in enumerate(rows) if r != rank)
            _sync(device)
    sweep.render_pass(hana.PASS_MAIN, model, shader, [uniforms], dtex, ntex)
    cptr, dptr, stride_px = sweep.device_planes()
    ctx.sync()
    for ptr in (cptr, dptr):
        exchange_bands(device_plane_tensor(ptr, W * H * 4, device), W * 4, rows, rank, world)
    moved["frame_bytes"] = 2 * sum((y1 - y0) * W * 4 for r, (y0, y1) in enumerate(rows) if r != rank)
    _sync(device)
    sweep.set_bands()
    return moved


def render_split_frame_async(ctx, hana, sweep, scene_objs, shader, uniforms, rank, world, device, exchange_shadow=True,
                             max_attempts=4):
    """render_split_frame without a host synchronisation inside the frame: the context launches on torch's current
    stream, so pass 1 -> all-gather of the shadow bands -> pass 2 -> all-gather of the colour / depth bands are ordered
    by the stream (NCCL collectives wait for, and are waited for by, the current stream). Scratch needs cannot be read
    back in between, so the ranks agree afterwards (one all-reduce of a flag) whether any of them ran out; if so all of
    them queue the frame again with the grown scratch. Returns the number of attempts."""
    import torch
    import torch.distributed as dist

    model, dtex, ntex = scene_objs
    W, H = sweep.width, sweep.height
    bands = tile_row_bands(H, world)
    rows = [band_pixel_rows(b, H) for b in bands]
    mine = bands[rank]
    past = ((H + TILE - 1) // TILE, 1)
    mine = past if mine[1] == 0 else mine
    shadowed = bool(uniforms.enable_shadow)
    sptr, pitch, sstride = sweep.shadow_plane()   # pointers are stable: fetch them before anything is queued
    cptr, dptr, _ = sweep.device_planes()
    splane = device_plane_tensor(sptr, sstride, device)
    planes = [device_plane_tensor(p, W * H * 4, device) for p in (cptr, dptr)]
    on_gpu = str(device).startswith("cuda")
    import contextlib
    scope = contextlib.nullcontext()
    if on_gpu:
        # a stream of torch's own (its default stream is handle 0, which the C ABI reads as "the context's stream")
        global _SPLIT_STREAM
        if _SPLIT_STREAM is None:
            _SPLIT_STREAM = torch.cuda.Stream()
        _SPLIT_STREAM.wait_stream(torch.cuda.current_stream())
        ctx.sync()
        ctx.set_stream(_SPLIT_STREAM.cuda_stream)
        scope = torch.cuda.stream(_SPLIT_STREAM)
    try:
      with scope:
        sweep.set_bands(shadow=mine if exchange_shadow else (0, 0), main=mine)
        for attempt in range(1, max_attempts + 1):
            if shadowed:
                sweep.render_pass_async(hana.PASS_SHADOW, model, shader, [uniforms], dtex, ntex)
                if exchange_shadow:
                    exchange_bands(splane, pitch, rows, rank, world)
            sweep.render_pass_async(hana.PASS_MAIN, model, shader, [uniforms], dtex, ntex)
            for pl in planes:
                exchange_bands(pl, W * 4, rows, rank, world)
            ok = torch.tensor([int(sweep.passes_ok())], dtype=torch.int32, device=device)
            if world > 1:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()):
                return attempt
        raise hana.HanaError(-5, "split frame still short of scratch after %d attempts" % max_attempts)
    finally:
        sweep.set_bands()
        if on_gpu:
            _SPLIT_STREAM.synchronize()
            ctx.set_stream(None)


_SPLIT_STREAM = None


def _sync(device):
    import torch

    if str(device).startswith("cuda"):
        torch.cuda.synchronize()
