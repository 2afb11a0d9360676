"""Frame-sharded sweeps (BASELINE.json configs[2]): independent frames, contiguous blocks per rank,
no collective on the data path. torch.distributed is only the launcher plumbing (barrier, gathering per-frame
checksums / timings); gloo on CPU boxes, nccl on GPUs."""
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
