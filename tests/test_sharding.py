"""N>1 path: frames sharded over ranks in contiguous blocks, no data-path collective.
CPU: world_size-2 gloo run of the host logic with the CPU oracle standing in for the renderer.
GPU: 2 ranks on 2 GPUs render their blocks; per-frame checksums must equal the single-GPU sweep's."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from conftest import load_package
hana = load_package()
from hana_softwarerenderer_b200.api import frame_checksum
mode = sys.argv[1]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
TOTAL, TURN, W, H = 10, 16, 96, 64
first, count = hana.sharding.frame_block(rank, world, TOTAL)
scene = hana.synthetic_scene("blob", tex=32)
if mode == "cpu":
    dist.init_process_group("gloo")
    from oracle import horacle as Hh
    port = Hh.Port()
    arr = hana.orbit_sweep_uniforms(W, H, first, count, frames_per_turn=TURN)
    sums = []
    for k in range(count):
        r = port.draw_model(Hh.BLINN, Hh.HanaUniforms.from_bytes(arr[k].to_bytes()), scene.a2v, W, H, diffuse=scene.diffuse, normal=scene.normal)
        sums.append(frame_checksum(r["color"], r["depth"]))
    allv = hana.sharding.gather_frame_values(np.array(sums, np.uint64), TOTAL, rank, world, "cpu")
else:
    import torch
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = hana.Context(local)
    objs = scene.upload(ctx)
    sums = hana.sharding.render_block(ctx, hana, objs, hana.BLINN, W, H, first, count, TURN, batch=4)
    allv = hana.sharding.gather_frame_values(sums, TOTAL, rank, world, "cuda")
if rank == 0:
    np.save(sys.argv[2], allv)
dist.barrier()
dist.destroy_process_group()
'''


def run_workers(mode, world, out, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + os.getpid() % 2000), str(script), mode, str(out)]
    subprocess.run(cmd, check=True, timeout=600, cwd=ROOT)
    return np.load(out)


def test_frame_block_partition(hana):
    for total in (1, 7, 1024, 1000):
        for world in (1, 2, 3, 8):
            blocks = [hana.sharding.frame_block(r, world, total) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f1 == f0 + c0
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def test_two_rank_gloo_matches_single_process(hana, tmp_path):
    two = run_workers("cpu", 2, tmp_path / "two.npy", tmp_path)
    one = run_workers("cpu", 1, tmp_path / "one.npy", tmp_path)
    assert two.shape == (10,) and np.array_equal(two, one)
    assert len(set(one.tolist())) == 10  # distinct cameras -> distinct frames


@pytest.mark.gpu
def test_two_gpus_match_one_gpu(hana, tmp_path):
    if hana.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    two = run_workers("gpu", 2, tmp_path / "two.npy", tmp_path)
    one = run_workers("gpu", 1, tmp_path / "one.npy", tmp_path)
    assert np.array_equal(two, one)


# ----------------------------------------------------------------------------- one frame split by screen tiles
SPLIT_WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from conftest import load_package
hana = load_package()
mode = sys.argv[1]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
W, H = 200, 136            # 9 tile rows (the last one partial): ragged bands on 2 ranks (one broadcast per band)
if mode == "cpu":
    # host logic only: every rank owns the rows of its band of a plane; after the exchange all ranks hold all rows
    dist.init_process_group("gloo")
    ok = True
    for height, row_bytes in ((H, W * 4), (128, 64), (16, 8)):
        bands = hana.sharding.tile_row_bands(height, world)
        rows = [hana.sharding.band_pixel_rows(b, height) for b in bands]
        truth = (np.arange(height * row_bytes, dtype=np.int64) * 2654435761 % 251).astype(np.uint8)
        plane = torch.zeros(height * row_bytes, dtype=torch.uint8)
        y0, y1 = rows[rank]
        plane[y0 * row_bytes:y1 * row_bytes] = torch.from_numpy(truth[y0 * row_bytes:y1 * row_bytes])
        hana.sharding.exchange_bands(plane, row_bytes, rows, rank, world)
        ok = ok and bool(np.array_equal(plane.numpy(), truth))
    res = np.array([int(ok)], np.uint64)
else:
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = hana.Context(local)
    scene = hana.synthetic_scene("blob", tex=64)
    objs = scene.upload(ctx)
    sums = []
    for (w, h) in ((W, H), (640, 512)):       # ragged bands (broadcasts), then equal bands (one in-place all-gather)
        u = hana.default_uniforms(w, h, True)
        for exchange in (True, False):
            sweep = ctx.sweep(w, h, 1)
            hana.sharding.render_split_frame(ctx, hana, sweep, objs, hana.NORMALMAP, u, rank, world, "cuda", exchange_shadow=exchange)
            sums.append(int(sweep.checksums(1)[0]))
            sweep.close()
        sweep = ctx.sweep(w, h, 1)   # the stream-ordered form: no host synchronisation inside the frame
        hana.sharding.render_split_frame_async(ctx, hana, sweep, objs, hana.NORMALMAP, u, rank, world, "cuda")
        sums.append(int(sweep.checksums(1)[0]))
        sweep.close()
    res = np.array(sums, np.uint64)
    t = torch.from_numpy(res.view(np.int64).copy()).cuda()
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    res = np.concatenate([o.cpu().numpy().view(np.uint64) for o in out])  # every rank must hold the same complete frame
if rank == 0:
    np.save(sys.argv[2], res)
dist.barrier()
dist.destroy_process_group()
'''


def run_split_workers(mode, world, out, tmp_path):
    script = tmp_path / "split_worker.py"
    script.write_text(SPLIT_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(31500 + os.getpid() % 2000), str(script), mode, str(out)]
    subprocess.run(cmd, check=True, timeout=600, cwd=ROOT)
    return np.load(out)


def test_tile_row_bands(hana):
    for height in (1, 16, 17, 136, 1080, 4320):
        rows_total = (height + 15) // 16
        for world in (1, 2, 3, 8):
            bands = hana.sharding.tile_row_bands(height, world)
            assert bands[0][0] == 0 and sum(c for _, c in bands) == rows_total
            for (f0, c0), (f1, _) in zip(bands, bands[1:]):
                assert f1 == f0 + c0
            px = [hana.sharding.band_pixel_rows(b, height) for b in bands]
            assert px[0][0] == 0 and max(y1 for _, y1 in px) == height
            assert all(a[1] == b[0] for a, b in zip(px, px[1:]))


def test_band_exchange_two_rank_gloo(hana, tmp_path):
    assert run_split_workers("cpu", 2, tmp_path / "x.npy", tmp_path).tolist() == [1]


@pytest.mark.gpu
def test_split_frame_bands_on_one_gpu(hana, ctx):
    """The per-GPU half of the tile split on ONE device: bands of tile rows rendered one after the other (each pass
    restricted to its band) compose the very frame an unrestricted render gives."""
    W, H = 200, 136
    scene = hana.synthetic_scene("blob", tex=64)
    model, dtex, ntex = scene.upload(ctx)
    u = hana.default_uniforms(W, H, True)
    ref = ctx.sweep(W, H, 1)
    ref.render(model, hana.NORMALMAP, [u], dtex, ntex)
    rc, rd = ref.download(0)
    for world in (2, 3, 4):
        bands = hana.sharding.tile_row_bands(H, world)
        sw = ctx.sweep(W, H, 1)
        for b in bands:  # pass 1, band by band, into the same maps
            sw.set_bands(shadow=b, main=b)
            sw.render_pass(hana.PASS_SHADOW, model, hana.NORMALMAP, [u], dtex, ntex)
        for b in bands:  # pass 2 from the complete maps
            sw.set_bands(shadow=b, main=b)
            sw.render_pass(hana.PASS_MAIN, model, hana.NORMALMAP, [u], dtex, ntex)
        c, d = sw.download(0)
        assert np.array_equal(d.view(np.uint32), rd.view(np.uint32)) and np.array_equal(c[..., :3], rc[..., :3]), world
        sw.close()
    for o in (ref, model, dtex, ntex):
        o.close()


@pytest.mark.gpu
def test_split_frame_api_errors_and_band_isolation(hana, ctx):
    """Argument checks of the band API, and that a band render leaves every tile row outside the band untouched."""
    W, H = 128, 96
    scene = hana.synthetic_scene("blob", tex=32)
    model, dtex, ntex = scene.upload(ctx)
    u = hana.default_uniforms(W, H, False)
    sw = ctx.sweep(W, H, 1)
    with pytest.raises(hana.HanaError):
        sw.set_bands(shadow=(-1, 2))
    with pytest.raises(hana.HanaError):
        sw.render_pass(7, model, hana.BLINN, [u], dtex, ntex)
    sw.render(model, hana.GROUND, [u], dtex, ntex, clear_rgba=(9, 9, 9, 1))      # whole frame, grey background
    c0, d0 = sw.download(0)
    sw.set_bands(main=(2, 2))                                                      # tile rows 2..3 = pixel rows 32..63
    sw.render_pass(hana.PASS_MAIN, model, hana.BLINN, [u], dtex, ntex, clear_rgba=(0, 0, 0, 1))
    c1, d1 = sw.download(0)
    assert np.array_equal(c1[:32], c0[:32]) and np.array_equal(c1[64:], c0[64:])   # outside the band: the first render
    assert np.array_equal(d1[:32], d0[:32]) and np.array_equal(d1[64:], d0[64:])
    assert not np.array_equal(c1[32:64], c0[32:64])                                # inside: re-rendered with Blinn on black
    sw.set_bands()
    sw.render(model, hana.BLINN, [u], dtex, ntex)
    c2, d2 = sw.download(0)
    assert np.array_equal(c1[32:64, :, :3], c2[32:64, :, :3]) and np.array_equal(d1[32:64], d2[32:64])
    for o in (sw, model, dtex, ntex):
        o.close()


@pytest.mark.gpu
def test_split_frame_two_gpus_match_one_gpu(hana, tmp_path):
    if hana.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    two = run_split_workers("gpu", 2, tmp_path / "two.npy", tmp_path)
    one = run_split_workers("gpu", 1, tmp_path / "one.npy", tmp_path)
    assert two.shape == (12,) and one.shape == (6,)
    assert one[0] == one[1] == one[2] and one[3] == one[4] == one[5] and one[0] != one[3]
    assert np.array_equal(two[:6], one) and np.array_equal(two[6:], one)  # both ranks hold both complete frames
