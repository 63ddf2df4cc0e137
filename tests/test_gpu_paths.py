"""-m gpu: multi-bounce path waves with compaction (vt_accel_trace_paths, the requeued generators, VT_TRAVERSE_QUEUE_ATTRS).

The bar: the compacted waves give, bit for bit, the image and the per-wave ray counts of (a) the same call with compaction
switched off and (b) the wave-by-wave sequence of the individual C-ABI calls (traverse + TraceResult, shadow rays, bounce rays),
which tests/test_gpu_parity.py pins to the reference."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = 0x9E3779B97F4A7C15


@pytest.fixture(scope="module")
def vt(built):
    import vistrace_b200

    assert vistrace_b200.lib().vt_device_count() >= 1, "no CUDA device"
    return vistrace_b200


def _piecewise(accel, scene, rays, bounces, sun, sun_rgb, seed, weight):
    """The same path waves through the individual host-buffer calls + numpy shading."""
    from vistrace_b200 import abi

    f4 = np.float32
    n = len(rays)
    fb = np.zeros((n, 3), f4)
    thr = np.ones((n, 3), f4)
    counts = [n]
    hits, attrs = accel.traverse(rays, want_attrs=True)
    alive = np.ones(n, bool)
    for k in range(bounces + 1):
        hit = alive & (attrs["prim"] != abi.VT_MISS)
        sky = hit & ((attrs["flags"] & abi.VT_ATTR_HIT_SKY) != 0)
        surf = hit & ~sky
        t = (thr * attrs["albedo"]).astype(f4)
        srays, _ = accel.shadow_rays(attrs, sun)
        # only live vertices spawn shadow rays: mask the dead ones the way the engine's queue does
        srays["tmax"][~surf] = -1.0
        occluded = accel.traverse(srays, any_hit=True)["prim"] != abi.VT_MISS
        counts.append(int(surf.sum()))
        fb[sky] += (f4(weight) * t[sky]).astype(f4)
        lit = surf & ~occluded
        fb[lit] += (f4(weight) * (t[lit] * np.asarray(sun_rgb, f4)[None, :]).astype(f4)).astype(f4)
        thr[surf] = t[surf]
        if k == bounces:
            break
        brays, _ = accel.bounce_rays(attrs, 1, seed=(seed + GOLD * (k + 1)) & 0xFFFFFFFFFFFFFFFF)
        brays["tmax"][~surf] = -1.0
        counts.append(int(surf.sum()))
        hits, attrs = accel.traverse(brays, want_attrs=True)
        alive = surf
    return fb, counts


@pytest.mark.parametrize("scene_name", ["heightfield", "props"])
def test_path_waves_compacted_equal_uncompacted_and_piecewise(vt, scene_name):
    import torch

    from vistrace_b200 import abi, scenes

    if scene_name == "heightfield":
        scene = scenes.scene_heightfield(64)  # open height field: many paths leave through the sky early
        rays = scenes.pinhole_rays(320, 180, (0, -80, 60), (0, 0, 5))
    else:
        scene = scenes.scene_props(8, 21, 11, 12)
        rays = scenes.pinhole_rays(320, 180, (0, -95, 40), (0, 0, 10))
    accel = vt.Accel(0).populate(scene)
    n, bounces, seed, weight = len(rays), 3, 77, 0.25
    sun = np.array((0.3, 0.2, 0.93), np.float32)
    sun = (sun / np.sqrt((sun * sun).sum(dtype=np.float32))).astype(np.float32)
    sun_rgb = (1.0, 0.9, 0.8)
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).cuda()
    sh = torch.cuda.current_stream().cuda_stream
    out = {}
    for compact in (True, False):
        d_fb = torch.zeros(n * 3, dtype=torch.float32, device="cuda")
        counts = accel.trace_paths_device(d_rays.data_ptr(), n, bounces, sun, sun_rgb, seed, weight, d_fb.data_ptr(), want_counts=True, compact=compact, stream=sh)
        torch.cuda.synchronize()
        out[compact] = (d_fb.cpu().numpy().reshape(-1, 3), counts)
    np.testing.assert_array_equal(out[True][1], out[False][1])
    np.testing.assert_array_equal(out[True][0], out[False][0])
    want_fb, want_counts = _piecewise(accel, scene, rays, bounces, sun, sun_rgb, seed, weight)
    assert list(out[True][1]) == want_counts, (list(out[True][1]), want_counts)
    np.testing.assert_array_equal(out[True][0], want_fb)
    c = out[True][1]
    assert c[0] == n and 0 < c[-1] < c[2] <= c[1] <= n  # paths die along the way: the later waves are sparse
    assert accel.invalid_rays == 0
    # a second call accumulates into the same framebuffer
    d_fb = torch.from_numpy(want_fb.reshape(-1).copy()).cuda()
    accel.trace_paths_device(d_rays.data_ptr(), n, bounces, sun, sun_rgb, seed, weight, d_fb.data_ptr(), stream=sh)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_fb.cpu().numpy().reshape(-1, 3), (want_fb + want_fb).astype(np.float32))
    # two samples in flight: the handle's two scratch slots on two streams, one framebuffer each — the images of the one-at-a-time calls
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    fbs = [torch.zeros(n * 3, dtype=torch.float32, device="cuda") for _ in range(2)]
    torch.cuda.synchronize()
    for it in range(3):  # three rounds: the slots are reused while the other one is still busy
        for k in range(2):
            accel.trace_paths_device(d_rays.data_ptr(), n, bounces, sun, sun_rgb, seed + 5 * k, weight, fbs[k].data_ptr(), stream=streams[k].cuda_stream, slot=k)
    torch.cuda.synchronize()
    for k in range(2):
        one = torch.zeros(n * 3, dtype=torch.float32, device="cuda")
        accel.trace_paths_device(d_rays.data_ptr(), n, bounces, sun, sun_rgb, seed + 5 * k, weight, one.data_ptr(), stream=sh)
        torch.cuda.synchronize()
        w = one.cpu().numpy()
        np.testing.assert_array_equal(fbs[k].cpu().numpy(), ((w + w).astype(np.float32) + w).astype(np.float32), err_msg=f"slot {k}")


def test_batched_sample_bsdf_diffuse_lobe_against_the_reference(vt, oracle_mod):
    """K3c (vt_accel_sample_bsdf_rays) against the reference's own SampleBSDF (oracle/_ref; the C port where it is absent) fed with
    the numbers the kernel drew (vt_sample_uniform01): lobe and the set of spawned rays exact, scattered / weight / pdf within
    1e-5 relative (device cosf / sinf / powf vs glibc), ray origin = CalcRayOrigin bit for bit.  Records come from a real
    wave (primary hits of a scene with metallic and rough MRAO materials) plus synthetic frames with metallic = 1 and back faces."""
    from test_oracle import _bsdf_case, _material_case
    from vistrace_b200 import abi, scenes

    kind = "reference" if oracle_mod.available("reference") else "port"
    scene, rays, _ = _material_case()
    accel = vt.Accel(0).populate(scene)
    hits, attrs = accel.traverse(rays, want_attrs=True)
    syn, wo_syn, _ = _bsdf_case(4000, seed=8)
    syn["prim"], syn["pos"] = 0, np.random.default_rng(1).uniform(-50, 50, (len(syn), 3)).astype(np.float32)
    syn["geometric_normal"] = syn["normal"]
    syn_rays = np.zeros(len(syn), abi.RAY)
    syn_rays["d"] = (-wo_syn * np.float32(2.5)).astype(np.float32)  # un-normalised: the kernel normalises like AccelStruct.cpp:826
    rays_all = np.concatenate([rays, syn_rays])
    attrs_all = np.concatenate([attrs, syn])
    n, spp, seed = len(attrs_all), 2, 4242
    out, samples, live = accel.sample_bsdf_rays(rays_all, attrs_all, spp, seed=seed)
    slots = np.arange(n * spp)
    rnd = np.stack([vt.sample_uniform01(slots, d, seed) for d in range(3)], 1)
    a2 = np.repeat(attrs_all, spp)
    d = np.repeat(rays_all["d"], spp, axis=0)
    f4 = np.float32
    dot = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]).astype(f4)
    with np.errstate(divide="ignore", invalid="ignore"):
        wo = (-(d * (f4(1) / np.sqrt(dot))[:, None])).astype(f4)  # -glm::normalize(dir)
    want, _ = oracle_mod.sample_bsdf_diffuse(a2, wo, rnd, kind)
    eligible = (a2["prim"] != abi.VT_MISS) & ((a2["flags"] & abi.VT_ATTR_HIT_SKY) == 0)
    assert eligible.sum() > 10000
    np.testing.assert_array_equal(samples["lobe"][eligible], want["lobe"][eligible])
    assert (samples["lobe"][~eligible] == 0).all() and (samples["pdf"][~eligible] == 0).all()
    lobe = eligible & (want["lobe"] == 1)
    assert 0 < (eligible & (want["lobe"] == 0)).sum()  # fully metallic records: nothing sampled
    for f in ("scattered", "weight"):
        g, w = samples[f][lobe].astype(np.float64), want[f][lobe].astype(np.float64)
        err = np.abs(g - w).max(-1) / np.maximum(np.sqrt((w * w).sum(-1)), 1e-6)
        assert err.max() <= 1e-5, (f, err.max())  # tolerance from BASELINE.json north_star
    g, w = samples["pdf"][lobe].astype(np.float64), want["pdf"][lobe].astype(np.float64)
    assert (np.abs(g - w) / np.maximum(np.abs(w), 1e-6)).max() <= 1e-5
    spawned = out["tmax"] >= 0
    np.testing.assert_array_equal(spawned, lobe & np.isfinite(samples["scattered"]).all(-1) & (samples["scattered"] != 0).any(-1))
    assert live == int(spawned.sum())
    np.testing.assert_array_equal(out["d"][spawned], samples["scattered"][spawned])
    side = np.where((samples["scattered"][spawned] * a2["geometric_normal"][spawned]).sum(-1, dtype=f4) >= 0, f4(1), f4(-1))[:, None]
    np.testing.assert_array_equal(out["o"][spawned], scenes.calc_ray_origin(a2["pos"][spawned], (a2["geometric_normal"][spawned] * side).astype(f4)))
    # cosine-weighted: E[cos(theta)] = 2/3 about the shading normal on the incident side
    ns = np.where((wo[spawned] * a2["normal"][spawned]).sum(-1) >= 0, 1.0, -1.0)[:, None] * a2["normal"][spawned]
    cos = (out["d"][spawned] * ns).sum(-1) / np.linalg.norm(out["d"][spawned], axis=1)
    assert abs(cos.mean() - 2.0 / 3.0) < 0.03


def _two_engines(vt, scene, monkeypatch):
    """The same scene and hierarchy with tail work-sharing off and on (VT_TAIL_SHARE is read when the engine is populated)."""
    monkeypatch.setenv("VT_TAIL_SHARE", "0")
    off = vt.Accel(0).populate(scene)
    monkeypatch.setenv("VT_TAIL_SHARE", "1")
    on = vt.Accel(0).populate(scene, bvh=off.get_bvh())
    return off, on


@pytest.mark.parametrize("scene_name", ["terrain", "foliage"])
def test_tail_work_sharing_gives_the_unshared_answer(vt, scene_name, monkeypatch):
    """K1's tail phase (idle lanes walk pending sub-trees of the warp's last rays, vt_traverse.cu) must not change a single byte:
    with the canonical tie rule the closest hit does not depend on who walks which sub-tree.  Launch sizes that run dry at once,
    sizes that are no multiple of the warp, waves with masked slots, grazing rays (the long chains the phase exists for)."""
    from vistrace_b200 import abi, scenes

    if scene_name == "terrain":
        scene = scenes.scene_terrain_closed(300)
        rays = scenes.pinhole_rays(640, 360, (0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
    else:
        scene = scenes.scene_foliage(n_cards=30000, tex_size=64, ground_quads=16)
        rays = scenes.pinhole_rays(480, 270, (0, -48, 20), (0, 0, 8))
    off, on = _two_engines(vt, scene, monkeypatch)
    assert on.layout == off.layout
    hits, attrs = off.traverse(rays, want_attrs=True)
    assert on.traverse(rays).tobytes() == hits.tobytes()
    horizon = len(rays) * 3 // 4
    for cnt in (1, 2, 31, 33, 100, 1000, 4097, 20001):
        sub = np.ascontiguousarray(rays[horizon:horizon + cnt])
        assert on.traverse(sub).tobytes() == off.traverse(sub).tobytes(), cnt
    for seed in (1, 2):
        brays, live = off.bounce_rays(attrs, 2, seed=seed)  # masked slots included
        assert live > 0
        a, b = off.traverse(brays), on.traverse(brays)
        assert a.tobytes() == b.tobytes()
        shadow = on.traverse(brays, any_hit=True)  # any-hit launches never share: same kernel either way
        assert shadow.tobytes() == off.traverse(brays, any_hit=True).tobytes()
    assert off.invalid_rays == on.invalid_rays == 0


def test_per_ray_statistics_add_up(vt):
    """vt_accel_traverse_ray_stats: the per-ray steps / tests sum to the totals of vt_accel_traverse_stats."""
    from vistrace_b200 import scenes

    scene = scenes.scene_terrain_closed(200)
    rays = scenes.pinhole_rays(320, 180, (0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
    accel = vt.Accel(0).populate(scene)
    steps, tests = accel.traverse_ray_stats(rays)
    tot_steps, tot_tests = accel.traverse_stats(rays)
    assert int(steps.sum()) == tot_steps and int(tests.sum()) == tot_tests
    assert steps.max() > 2 * steps.mean()  # a few grazing rays are far longer than the rest: what bounds small launches


def test_two_host_frames_in_flight_equal_the_synchronous_frames(vt):
    """vt_accel_render_diffuse_wave_begin / _wait: frames begun back to back (two in flight, alternating staging buffers, tiles of
    consecutive frames sharing the wave lanes) land the images of the synchronous call, in order; a third begin is refused."""
    import torch

    from vistrace_b200 import abi, scenes

    scene = scenes.scene_terrain_closed(200)
    rays = scenes.pinhole_rays(400, 225, (0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
    n, spp = len(rays), 3
    accel = vt.Accel(0).populate(scene)
    want = [accel.render_diffuse_wave(rays, spp, seed=40 + k, weight=0.25)[0].copy() for k in range(5)]
    h_rays_t = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
    h_rays = h_rays_t.numpy().view(abi.RAY)
    h_rays[:] = rays
    fb_t = [torch.empty(n * 12, dtype=torch.uint8).pin_memory() for _ in range(2)]
    fbs = [t.numpy().view(np.float32).reshape(n, 3) for t in fb_t]
    import os
    os.environ["VT_WAVE_TILE"] = "20000"  # several tiles per frame over the four lanes
    try:
        got = []
        for k in range(5):
            if k >= 2:
                accel.render_diffuse_wave_wait()
                got.append(fbs[k % 2].copy())  # frame k - 2 is complete; its buffer is about to be reused
            fbs[k % 2][:] = -1.0
            accel.render_diffuse_wave_begin(h_rays, spp, 40 + k, 0.25, fbs[k % 2])
        with pytest.raises(RuntimeError, match="two frames"):
            accel.render_diffuse_wave_begin(h_rays, spp, 99, 0.25, fbs[0])
        for k in (3, 4):
            accel.render_diffuse_wave_wait()
            got.append(fbs[k % 2].copy())
        with pytest.raises(RuntimeError, match="no frame in flight"):
            accel.render_diffuse_wave_wait()
    finally:
        del os.environ["VT_WAVE_TILE"]
    for k in range(5):
        np.testing.assert_array_equal(got[k], want[k], err_msg=f"frame {k}")
    # the synchronous call still works afterwards
    np.testing.assert_array_equal(accel.render_diffuse_wave(rays, spp, seed=40, weight=0.25)[0], want[0])
