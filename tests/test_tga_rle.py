"""RLE TGA output (TGAImage::write_tga_file(rle=true) + unload_rle_data, tgaimage.cpp:145-246).

CPU: the parallel packetiser's model (tests/rle_model.py, the statement the CUDA kernel in csrc/hana_tga.cuh transcribes)
against the sequential algorithm on adversarial pixel streams, and the sequential model against the library's host writer
(hana_tga_write, itself pinned byte-identical to the reference's writer by tests/test_assets_present.py).
GPU: files encoded on the device (hana_sweep_encode_tga / hana_sweep_fetch_tga) against hana_tga_write, byte for byte, on
rendered frames and on crafted frames whose runs sit on every packet-length boundary."""
import os

import numpy as np
import pytest

from rle_model import parallel_packets, sequential_packets


def crafted_stream(rng, n, kind):
    if kind == 0:
        return rng.randint(0, 3, n).astype(np.uint32)
    if kind == 1:
        out, have = [], 0
        while have < n:
            L = int(rng.choice([1, 2, 3, 4, 125, 126, 127, 128, 129, 130, 253, 254, 255, 256, 257, 383, 384, 385, rng.randint(1, 700)]))
            if rng.rand() < 0.5:
                out.append(np.full(L, rng.randint(0, 1 << 24), np.uint32))
            else:
                v = rng.randint(0, 1 << 24, L).astype(np.uint32)
                v[1:][v[1:] == v[:-1]] ^= 1
                out.append(v)
            have += L
        return np.concatenate(out)[:n]
    if kind == 2:
        return np.zeros(n, np.uint32)
    if kind == 3:
        return np.arange(n, dtype=np.uint32)
    return np.repeat(rng.randint(0, 1 << 24, (n + 1) // 2).astype(np.uint32), 2)[:n]


def test_parallel_model_equals_sequential_algorithm():
    rng = np.random.RandomState(1)
    for trial in range(300):
        n = int(rng.choice([1, 2, 3, 5, 127, 128, 129, 255, 256, 257, 1000, rng.randint(1, 3000)]))
        px = crafted_stream(rng, n, trial % 5)
        ref = sequential_packets(px)
        for chunk in (1, 7, 64, 4096):
            assert parallel_packets(px, chunk) == ref, (trial, n, chunk)


def test_emulated_kernels_equal_sequential_algorithm():
    """csrc/hana_tga_core.cuh — the arithmetic the encoder's kernels are made of — compiled for the host and walked with
    the kernels' control flow (tests/emu/emu_tga.cpp): payload byte-identical to the sequential algorithm's, the closed-form
    byte count of every word equal to the sum over its pixels, whatever the number of threads of the structure kernel."""
    from emu_tga_check import run
    rng = np.random.RandomState(4)
    for trial in range(400):
        n = int(rng.choice([1, 2, 31, 32, 33, 127, 128, 129, 255, 256, 257, 1023, 1024, 1025, 4096, rng.randint(1, 6000)]))
        px = crafted_stream(rng, n, trial % 5)
        ref = sequential_packets(px)
        for threads in (1, 5, 1024):
            assert run(px, threads) == ref, (trial, n, threads)


def test_sequential_model_equals_host_writer(hana, tmp_path):
    rng = np.random.RandomState(2)
    for kind in range(5):
        w, h = 97, 41
        px = crafted_stream(rng, w * h, kind)
        img = np.stack([px & 255, (px >> 8) & 255, (px >> 16) & 255], -1).astype(np.uint8).reshape(h, w, 3)
        p = str(tmp_path / ("k%d.tga" % kind))
        hana.tga_write(p, img, rle=True)
        data = open(p, "rb").read()
        assert data[18:-26] == sequential_packets(px)


def host_file(hana, tmp_path, rgba, name):
    """What TGAImage::write_tga_file(rle=true) writes for a y-up RGBA8 frame: rows top-down, B,G,R."""
    p = str(tmp_path / name)
    hana.tga_write(p, np.ascontiguousarray(rgba[::-1, :, 2::-1]), rle=True)
    return open(p, "rb").read()


@pytest.mark.gpu
def test_device_files_of_rendered_frames(hana, ctx, african_head, blob, tmp_path):
    for scene, (W, Hh), F in ((blob, (320, 240), 5), (african_head, (1920, 1080), 3), (blob, (333, 217), 2)):
        arr = hana.orbit_sweep_uniforms(W, Hh, 7, F, frames_per_turn=64)
        objs = scene.upload(ctx)
        sw = ctx.sweep(W, Hh, F)
        sw.render(objs[0], hana.BLINN, arr, objs[1], objs[2])
        files = sw.tga_files(0, F)
        for f in range(F):
            col, _ = sw.download(f)
            want = host_file(hana, tmp_path, col, "f%d.tga" % f)
            assert len(files[f]) == len(want) and files[f] == want, "frame %d of %dx%d: device file differs" % (f, W, Hh)
        sub = sw.tga_files(1, 1)  # a sub-range, encoded again
        assert sub[0] == files[1]
        for o in (sw,) + tuple(objs):
            o.close()


@pytest.mark.gpu
def test_device_files_of_crafted_frames(hana, ctx, tmp_path):
    """Pixel streams written straight into the frame ring: runs and raw stretches of 126..130, 254..258, ... pixels,
    crossing rows and the encoder's 4096-pixel chunks, all-equal and all-different frames."""
    import torch
    from hana_softwarerenderer_b200.sharding import device_plane_tensor
    rng = np.random.RandomState(3)
    for (W, Hh) in ((640, 96), (1000, 37), (128, 128), (4099, 5), (1, 1), (3, 7), (31, 1), (33, 2), (127, 3), (132, 9), (2048, 2)):
        F = 10
        sw = ctx.sweep(W, Hh, F)
        cptr, _, stride = sw.device_planes()
        ring = device_plane_tensor(cptr, stride * 4 * F, "cuda:0")
        frames = []
        for f in range(F):
            px = crafted_stream(rng, W * Hh, f % 5)
            rgba = np.zeros((Hh, W, 4), np.uint8)
            file_order = np.stack([(px >> 16) & 255, (px >> 8) & 255, px & 255], -1).astype(np.uint8).reshape(Hh, W, 3)  # R,G,B with px = B|G<<8|R<<16
            rgba[::-1, :, :3] = file_order
            rgba[..., 3] = rng.randint(0, 256)  # alpha takes no part
            frames.append(rgba)
        ring.copy_(torch.from_numpy(np.stack(frames).reshape(-1)))
        torch.cuda.synchronize()
        files = sw.tga_files(0, F)
        for f in range(F):
            want = host_file(hana, tmp_path, frames[f], "c%d.tga" % f)
            assert files[f] == want, "crafted frame %d (%dx%d, kind %d): device file differs" % (f, W, Hh, f % 5)
        sw.close()


@pytest.mark.gpu
def test_device_files_of_large_frames(hana, ctx, tmp_path):
    """BASELINE.json configs[3] / configs[4] frame sizes (3840x2160, 7680x4320): 260 k / 1 M words per frame, spans of 254 /
    1013 words per thread of the structure kernel, batch offsets beyond 2^24 — crafted streams, byte for byte."""
    import torch
    from hana_softwarerenderer_b200.sharding import device_plane_tensor
    rng = np.random.RandomState(8)
    for (W, Hh, kinds) in ((3840, 2160, (1, 0)), (7680, 4320, (1,))):
        F = len(kinds)
        sw = ctx.sweep(W, Hh, F)
        cptr, _, stride = sw.device_planes()
        ring = device_plane_tensor(cptr, stride * 4 * F, "cuda:0")
        frames = []
        for kind in kinds:
            px = crafted_stream(rng, W * Hh, kind)
            rgba = np.zeros((Hh, W, 4), np.uint8)
            rgba[::-1, :, :3] = np.stack([(px >> 16) & 255, (px >> 8) & 255, px & 255], -1).astype(np.uint8).reshape(Hh, W, 3)
            rgba[..., 3] = 7
            frames.append(rgba)
        ring.copy_(torch.from_numpy(np.stack(frames).reshape(-1)))
        torch.cuda.synchronize()
        files = sw.tga_files(0, F)
        for f in range(F):
            want = host_file(hana, tmp_path, frames[f], "L%d.tga" % f)
            assert len(files[f]) == len(want) and files[f] == want, "large crafted frame %dx%d kind %d: device file differs" % (W, Hh, kinds[f])
        sw.close()
