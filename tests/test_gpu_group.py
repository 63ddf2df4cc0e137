"""-m gpu: the native multi-GPU group (vistrace_b200/csrc/vt_group.cu, include/vistrace_b200.h "multi-GPU") through the C ABI.

With one GPU on the box the single-member forms of both kinds of group run (same code path: shard geometry, compact shard
buffers, tile-strided copies); with two or more, a single-process group over two GPUs and a two-rank multi-process group
(torchrun worker, NCCL over NVLink) run as well.  The bar is the strongest one available: the group's hit buffers and images
equal the single-GPU call's BYTE FOR BYTE (which the parity tests pin to the reference).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vt(built):
    import vistrace_b200

    assert vistrace_b200.lib().vt_device_count() >= 1, "no CUDA device"
    return vistrace_b200


def _case():
    from vistrace_b200 import scenes

    scene = scenes.scene_heightfield(64)
    rays = scenes.pinhole_rays(333, 187, (0, -80, 60), (0, 0, 5))  # 62 271 pixels: ragged last tile for every tile size used below
    return scene, rays


def _check_group(vt, group, scene, rays, monkeypatch):
    single = vt.Accel(0).populate(scene)
    want_hits, want_attrs = single.traverse(rays, want_attrs=True)
    hits, attrs = group.traverse(rays, want_attrs=True)
    assert hits.tobytes() == want_hits.tobytes() and attrs.tobytes() == want_attrs.tobytes()
    assert len(group.traverse(rays[:0])) == 0
    assert group.traverse(rays[:5]).tobytes() == want_hits[:5].tobytes()  # fewer rays than some tile / slice sizes
    spp = 3
    want_img, want_live = single.render_diffuse_wave(rays, spp, seed=9, weight=0.5)
    for tile in ("8192", "1000", "64"):
        monkeypatch.setenv("VT_GROUP_TILE", tile)
        for chunk in ("524288", "3000"):  # one chunk per shard / many chunks over the wave lanes
            monkeypatch.setenv("VT_WAVE_TILE", chunk)
            img, live = group.render_diffuse_wave(rays, spp, seed=9, weight=0.5)
            assert live == want_live, (tile, chunk)
            np.testing.assert_array_equal(img, want_img, err_msg=f"tile {tile} chunk {chunk}")
    assert np.abs(want_img).sum() > 0


def test_single_process_group_one_gpu(vt, monkeypatch):
    scene, rays = _case()
    group = vt.Group(devices=[0]).populate(scene)
    assert (group.world, group.rank, group.local_members) == (1, 0, 1)
    _check_group(vt, group, scene, rays, monkeypatch)
    launches = group.launch_count
    assert launches > 0
    group.close()


def test_multi_process_group_of_one_rank(vt, monkeypatch):
    scene, rays = _case()
    group = vt.Group(device=0, rank=0, world=1).populate(scene)
    _check_group(vt, group, scene, rays, monkeypatch)
    # device-resident shard in, frame-sized device image out, everything on the caller's stream
    import torch

    n, spp = len(rays), 2
    idx = group.shard_indices(n)
    d_rays = torch.from_numpy(np.ascontiguousarray(rays[idx]).view(np.uint8).reshape(-1).copy()).cuda()
    d_fb = torch.full((n * 3,), -1.0, dtype=torch.float32, device="cuda")
    group.render_diffuse_wave_device(d_rays.data_ptr(), n, spp, 4, 1.0, d_fb.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want, _ = vt.Accel(0).populate(scene).render_diffuse_wave(rays, spp, seed=4, weight=1.0)
    np.testing.assert_array_equal(d_fb.cpu().numpy().reshape(-1, 3), want)
    group.close()


def test_replica_handles_refuse_host_side_queries(vt):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    scene, rays = _case()
    group = vt.Group(devices=[0, 1]).populate(scene)
    replica = group.accel(1)
    assert replica.traverse(rays).tobytes() == group.accel(0).traverse(rays).tobytes()  # the replica is a complete scene
    with pytest.raises(RuntimeError):
        replica.get_bvh()
    with pytest.raises(RuntimeError):
        replica.refit(scene)
    group.close()


def test_single_process_group_two_gpus(vt, monkeypatch):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    scene, rays = _case()
    group = vt.Group(devices=[0, 1]).populate(scene)
    assert (group.world, group.local_members) == (2, 2)
    _check_group(vt, group, scene, rays, monkeypatch)
    group.close()


def test_multi_process_group_two_ranks_nccl(vt):
    """One process per GPU (the bench.py --gpus N arrangement): rank 0 builds, the image is ncclBroadcast, shards are gathered on
    rank 0 with ncclSend / ncclRecv — and rank 0's results equal the single-GPU ones byte for byte."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29573", f"{ROOT}/tests/group_worker.py"], capture_output=True, text=True, timeout=150)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GROUP_OK" in r.stdout
