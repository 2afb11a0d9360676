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
