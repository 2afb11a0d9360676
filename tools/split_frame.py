#!/usr/bin/env python
"""One very large frame split by screen tiles over N GPUs (north_star's optional mode):
BASELINE.json configs[4] (9 216 large triangles, depth complexity ~8, 7680x4320, NormalMap + shadow).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/split_frame.py [--reps 5]

Every rank rasterises its band of tile rows of both passes; the shadow-map bands are all-gathered between the passes
and the colour + depth bands after pass 2 (NCCL over NVLink, hana sharding.render_split_frame). Rank 0 prints one JSON
line: ms per frame (max over ranks, device-synchronised wall clock around the whole split render, exchanges included),
the bytes each rank received, and whether the frame's checksum equals the unsplit render of the same GPU."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 5
    small = "--small" in sys.argv
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hana = ge.load_package()
    ctx = hana.Context(local)
    W, H = (1920, 1080) if small else (7680, 4320)
    a2v = hana.scene.synthetic_layers(8, 32, 18, seed=99)
    dif, nm = hana.scene.noise_textures(99, 1024)
    sc = hana.Scene("c5", a2v, dif, nm)
    objs = sc.upload(ctx)
    u = hana.default_uniforms(W, H, True)

    ref = ctx.sweep(W, H, 1)
    ref.render(objs[0], hana.NORMALMAP, [u], objs[1], objs[2])
    want = int(ref.checksums(1)[0])
    ctx.timer_start()
    for _ in range(reps):
        ref.render(objs[0], hana.NORMALMAP, [u], objs[1], objs[2])
    ms_unsplit = ctx.timer_stop() / reps
    ref.close()

    out = {}
    for exchange in (True, False):
        sw = ctx.sweep(W, H, 1)
        moved = hana.sharding.render_split_frame(ctx, hana, sw, objs, hana.NORMALMAP, u, rank, world, "cuda", exchange_shadow=exchange)
        got = int(sw.checksums(1)[0])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            hana.sharding.render_split_frame(ctx, hana, sw, objs, hana.NORMALMAP, u, rank, world, "cuda", exchange_shadow=exchange)
        torch.cuda.synchronize()
        t = torch.tensor([(time.perf_counter() - t0) * 1e3 / reps], dtype=torch.float64, device="cuda")
        ok = torch.tensor([int(got == want)], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        out["exchange_shadow_bands" if exchange else "redundant_shadow_pass"] = {
            "ms_per_frame": float(t.item()), "checksum_equals_unsplit_on_every_rank": bool(ok.item()), **moved}
        sw.close()
    # stream-ordered form: passes and NCCL exchanges on one stream, one host synchronisation per frame
    sw = ctx.sweep(W, H, 1)
    attempts = hana.sharding.render_split_frame_async(ctx, hana, sw, objs, hana.NORMALMAP, u, rank, world, "cuda")
    got = int(sw.checksums(1)[0])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        hana.sharding.render_split_frame_async(ctx, hana, sw, objs, hana.NORMALMAP, u, rank, world, "cuda")
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) * 1e3 / reps], dtype=torch.float64, device="cuda")
    ok = torch.tensor([int(got == want)], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    out["stream_ordered"] = {"ms_per_frame": float(t.item()), "checksum_equals_unsplit_on_every_rank": bool(ok.item()),
                             "first_call_attempts": attempts}
    sw.close()
    if rank == 0:
        print(json.dumps({"config": "configs[4] split by tile rows", "width": W, "height": H, "n_gpus": world, "faces": sc.nfaces,
                          "ms_per_frame_unsplit_1gpu": ms_unsplit, **out}))
    for o in objs:
        o.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
