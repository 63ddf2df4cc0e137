"""Seeded synthetic scenes and ray batches for the BASELINE.json configs.

The reference ingests Source-engine maps and models through a running game
(source/objects/AccelStruct.cpp:183-499, 566-758); none of that exists headless,
so the benchmark configs use procedural meshes that reproduce the *contract* of
ingestion: world brushes / displacements are one-sided triangles of entity 0
(AccelStruct.cpp:408-411), prop triangles are two-sided (source/objects/Model.cpp:82-89),
foliage uses `alphatest | nocull` materials over an RGBA8888 VTF mip chain.

Everything is float32 numpy, vectorised, and deterministic for a given seed.
"""
import numpy as np

from . import abi

f4 = np.float32


# --------------------------------------------------------------------- helpers
def _hash01(ix, iy, seed):
    """Integer lattice hash -> [0,1) float64 (deterministic, vectorised)."""
    h = (ix.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ (iy.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F))
    h ^= np.uint64((int(seed) * 0x165667B19E3779F9) & 0xFFFFFFFFFFFFFFFF)
    h ^= h >> np.uint64(29)
    h *= np.uint64(0xBF58476D1CE4E5B9)
    h ^= h >> np.uint64(32)
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def value_noise(x, y, seed, octaves=5):
    """Multi-octave value noise in [-1, 1]."""
    out = np.zeros_like(x, dtype=np.float64)
    amp, freq, norm = 1.0, 1.0, 0.0
    for o in range(octaves):
        xs, ys = x * freq, y * freq
        x0, y0 = np.floor(xs), np.floor(ys)
        fx, fy = xs - x0, ys - y0
        sx, sy = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy)
        ix, iy = x0.astype(np.int64) & 0xFFFFF, y0.astype(np.int64) & 0xFFFFF
        a = _hash01(ix, iy, seed + o)
        b = _hash01(ix + 1, iy, seed + o)
        c = _hash01(ix, iy + 1, seed + o)
        d = _hash01(ix + 1, iy + 1, seed + o)
        out += amp * ((a * (1 - sx) + b * sx) * (1 - sy) + (c * (1 - sx) + d * sx) * sy)
        norm += amp
        amp *= 0.5
        freq *= 2.0
    return out / norm * 2.0 - 1.0


def _normalize(v):
    n = np.sqrt((v * v).sum(-1, keepdims=True))
    return v / np.maximum(n, 1e-20)


def _tris_from_indexed(verts, normals, tangents, uvs, faces, material=0, ent_idx=0, one_sided=False, alphas=None):
    """Indexed mesh -> TRI_IN records."""
    t = np.zeros(len(faces), abi.TRI_IN)
    t["p"] = verts[faces].astype(f4)
    t["normals"] = normals[faces].astype(f4)
    t["tangents"] = tangents[faces].astype(f4)
    t["uvs"] = uvs[faces].astype(f4)
    t["alphas"] = 1.0 if alphas is None else alphas[faces].astype(f4)
    t["material"] = material
    t["ent_idx"] = ent_idx
    t["one_sided"] = 1 if one_sided else 0
    return t


def reference_normal(tris):
    """n = cross(e1, e2), e1 = p0 - p1, e2 = p2 - p0 (source/objects/Primitives.h:82,93)."""
    p = tris["p"].astype(np.float64)
    return np.cross(p[:, 0] - p[:, 1], p[:, 2] - p[:, 0])


def orient(tris, want):
    """Flip winding where the reference normal disagrees with `want` (n_tris x 3 direction)."""
    n = reference_normal(tris)
    flip = (n * want).sum(-1) < 0
    for k in ("p", "normals", "tangents", "uvs", "alphas"):
        a = tris[k]
        tmp = a[flip, 1].copy()
        a[flip, 1] = a[flip, 2]
        a[flip, 2] = tmp
    return tris


# ---------------------------------------------------------------------- meshes
def heightfield(n_quads, extent=100.0, amp=8.0, seed=1234, noise_scale=0.06, material=0, uv_tiles=16.0):
    """Value-noise displaced grid: n_quads^2 quads = 2 n_quads^2 one-sided world triangles
    over [-extent/2, extent/2]^2, front faces up."""
    n = n_quads + 1
    lin = np.linspace(-extent / 2, extent / 2, n)
    gx, gy = np.meshgrid(lin, lin, indexing="xy")
    h = value_noise(gx * noise_scale, gy * noise_scale, seed) * amp
    verts = np.stack([gx, gy, h], -1).reshape(-1, 3)
    # per-vertex normals from central differences
    step = extent / n_quads
    dzdx = np.gradient(h, step, axis=1)
    dzdy = np.gradient(h, step, axis=0)
    nrm = _normalize(np.stack([-dzdx, -dzdy, np.ones_like(h)], -1)).reshape(-1, 3)
    tan = _normalize(np.stack([np.ones_like(h), np.zeros_like(h), dzdx], -1)).reshape(-1, 3)
    uvs = np.stack([(gx / extent + 0.5) * uv_tiles, (gy / extent + 0.5) * uv_tiles], -1).reshape(-1, 2)
    i, j = np.meshgrid(np.arange(n_quads), np.arange(n_quads), indexing="xy")
    v00 = (j * n + i).ravel()
    v10, v01, v11 = v00 + 1, v00 + n, v00 + n + 1
    faces = np.concatenate([np.stack([v00, v10, v11], -1), np.stack([v00, v11, v01], -1)], 0)
    # interleave the two triangles of a quad so neighbours stay close in memory
    faces = faces.reshape(2, -1, 3).transpose(1, 0, 2).reshape(-1, 3)
    alphas = (0.5 + 0.5 * value_noise(gx * 0.02, gy * 0.02, seed + 77, 3)).reshape(-1)
    t = _tris_from_indexed(verts, nrm, tan, uvs, faces, material, 0, True, alphas)
    return orient(t, np.array([0.0, 0.0, 1.0]))


def box(lo, hi, inward=True, material=0, ent_idx=0, one_sided=True):
    """12 triangles; inward=True gives a room whose front faces look at the inside."""
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    c = np.array([[lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [hi[0], hi[1], lo[2]], [lo[0], hi[1], lo[2]],
                  [lo[0], lo[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], hi[2]]])
    quads = [(0, 1, 2, 3, (0, 0, -1)), (4, 5, 6, 7, (0, 0, 1)), (0, 1, 5, 4, (0, -1, 0)),
             (2, 3, 7, 6, (0, 1, 0)), (1, 2, 6, 5, (1, 0, 0)), (3, 0, 4, 7, (-1, 0, 0))]
    t = np.zeros(12, abi.TRI_IN)
    want = np.zeros((12, 3))
    for q, (a, b, cc, d, out) in enumerate(quads):
        out = np.array(out, np.float64)
        face_n = -out if inward else out
        tang = _normalize(c[b] - c[a])
        for k, idx in enumerate(((a, b, cc), (a, cc, d))):
            r = t[2 * q + k]
            r["p"] = c[list(idx)]
            r["normals"] = face_n
            r["tangents"] = tang
            uvq = {a: (0, 0), b: (1, 0), cc: (1, 1), d: (0, 1)}
            r["uvs"] = [uvq[i] for i in idx]
            want[2 * q + k] = face_n
    t["alphas"] = 1.0
    t["material"] = material
    t["ent_idx"] = ent_idx
    t["one_sided"] = 1 if one_sided else 0
    return orient(t, want)


def torus(nu, nv, R=1.0, r=0.4, bump=0.08, seed=0, material=0, ent_idx=0):
    """Displaced torus, nu*nv quads = 2*nu*nv two-sided prop triangles in object space."""
    u = np.arange(nu) / nu * 2 * np.pi
    v = np.arange(nv) / nv * 2 * np.pi
    uu, vv = np.meshgrid(u, v, indexing="xy")
    rr = r * (1.0 + bump * value_noise(uu * 3 / np.pi + 11.0, vv * 3 / np.pi + 5.0, seed, 3))
    cx, cy = np.cos(uu), np.sin(uu)
    verts = np.stack([(R + rr * np.cos(vv)) * cx, (R + rr * np.cos(vv)) * cy, rr * np.sin(vv)], -1).reshape(-1, 3)
    nrm = np.stack([np.cos(vv) * cx, np.cos(vv) * cy, np.sin(vv)], -1).reshape(-1, 3)
    tan = np.stack([-cy, cx, np.zeros_like(cx)], -1).reshape(-1, 3)
    uvs = np.stack([uu / (2 * np.pi) * 4.0, vv / (2 * np.pi)], -1).reshape(-1, 2)
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="xy")
    v00 = (j * nu + i).ravel()
    v10 = (j * nu + (i + 1) % nu).ravel()
    v01 = (((j + 1) % nv) * nu + i).ravel()
    v11 = (((j + 1) % nv) * nu + (i + 1) % nu).ravel()
    faces = np.concatenate([np.stack([v00, v10, v11], -1), np.stack([v00, v11, v01], -1)], 0)
    faces = faces.reshape(2, -1, 3).transpose(1, 0, 2).reshape(-1, 3)
    t = _tris_from_indexed(verts, nrm, tan, uvs, faces, material, ent_idx, False)
    # wind so the reference normal points out of the surface
    ctr = t["p"].astype(np.float64).mean(1)
    ang = np.arctan2(ctr[:, 1], ctr[:, 0])
    ring = np.stack([R * np.cos(ang), R * np.sin(ang), np.zeros_like(ang)], -1)
    return orient(t, ctr - ring)


def transform_tris(tris, rot, trans, scale=1.0):
    """Rigid transform baked into world space, float32 arithmetic (the scene generator's own
    bake; the reference's SkinTriangle path is exercised separately)."""
    out = tris.copy()
    rot = np.asarray(rot, np.float64)
    out["p"] = (tris["p"].astype(np.float64) * scale @ rot.T + np.asarray(trans, np.float64)).astype(f4)
    out["normals"] = (tris["normals"].astype(np.float64) @ rot.T).astype(f4)
    out["tangents"] = (tris["tangents"].astype(np.float64) @ rot.T).astype(f4)
    return out


def _random_rotations(rng, n):
    q = _normalize(rng.normal(size=(n, 4)))
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], 1)


def rigid_mat4(rot, trans, scale=1.0):
    """glm::mat4 (column-major, 16 floats) of x -> scale * rot @ x + trans."""
    m = np.zeros((4, 4), np.float64)
    m[:3, :3] = np.asarray(rot, np.float64) * scale
    m[:3, 3] = trans
    m[3, 3] = 1.0
    return m.T.astype(f4).reshape(16)  # element [col*4 + row]


def skin_case(n_tris=600, n_bones=5, seed=77):
    """Deterministic input for SkinTriangle (source/objects/AccelStruct.cpp:66-101): object-space triangles with
    normals/tangents, 1-3 weighted bones per vertex, random bone and bind-pose matrices."""
    rng = np.random.default_rng(seed)
    tris = torus(25, 12, seed=seed)[:n_tris].copy()
    n = len(tris)
    skin = np.zeros(n, abi.TRI_SKIN)
    skin["num_bones"] = rng.integers(1, 4, (n, 3))
    skin["bone_ids"] = rng.integers(0, n_bones, (n, 3, 3))
    w = rng.uniform(0.05, 1.0, (n, 3, 3))
    w *= np.arange(3)[None, None, :] < skin["num_bones"][:, :, None]
    skin["weights"] = (w / w.sum(-1, keepdims=True)).astype(f4)
    rots = _random_rotations(rng, 2 * n_bones)
    bones = np.stack([rigid_mat4(rots[b], rng.uniform(-50, 50, 3), rng.uniform(0.5, 3.0)) for b in range(n_bones)])
    binds = np.stack([rigid_mat4(rots[n_bones + b], rng.uniform(-2, 2, 3)) for b in range(n_bones)])
    return tris, skin, bones, binds


def scene_props_skinned(n_props=256, nu=63, nv=31, ground_quads=64, seed=4321, extent=200.0):
    """Config 2 the way PopulateAccel builds it: every prop is an object-space mesh baked to world space by
    SkinTriangle with its entity's bone matrix (the one-bone overload, source/objects/AccelStruct.cpp:103-108, as for
    a rigid prop) through vt_skin_triangles; entity 0 = one-sided world (ground + room)."""
    from . import binding

    rng = np.random.default_rng(seed)
    parts = [heightfield(ground_quads, extent, 2.0, seed, 0.03), box([-extent / 2, -extent / 2, -8.0], [extent / 2, extent / 2, 80.0], True, material=1)]
    rots = _random_rotations(rng, n_props)
    pos = np.stack([rng.uniform(-extent * 0.45, extent * 0.45, n_props), rng.uniform(-extent * 0.45, extent * 0.45, n_props),
                    rng.uniform(4.0, 40.0, n_props)], -1)
    scl = rng.uniform(2.0, 6.0, n_props)
    ident = rigid_mat4(np.eye(3), (0, 0, 0))
    for e in range(n_props):
        obj = torus(nu, nv, seed=seed + e, material=2 + (e % 6), ent_idx=e + 1)
        parts.append(binding.skin_triangles(obj, None, rigid_mat4(rots[e], pos[e], scl[e])[None], ident[None]))
    mats = abi.default_materials(8)
    mats["surf_flags"][1] = abi.VT_SURF_SKY
    mats["colour"][2:, :3] = rng.uniform(0.2, 1.0, (6, 3))
    ents = np.zeros(n_props + 1, abi.ENTITY)
    ents["id"] = np.concatenate([[0], 100 + np.arange(n_props)])
    ents["colour"] = 1.0
    ents["colour"][1:, :3] = rng.uniform(0.5, 1.0, (n_props, 3))
    return abi.SceneData(np.concatenate(parts), mats, ents)


def leaf_texture(size=256, seed=5, clamp=False):
    """Synthetic RGBA8888 'leaf' texture with a full mip chain, returned in VTF memory
    order (smallest mip first, libs/VTFParser/VTFParser.cpp:44-78)."""
    lin = (np.arange(size) + 0.5) / size
    x, y = np.meshgrid(lin, lin, indexing="xy")
    cx, cy = x - 0.5, y - 0.5
    ang = np.arctan2(cy, cx)
    rad = np.sqrt(cx * cx + cy * cy)
    edge = 0.36 + 0.09 * np.cos(5 * ang + seed) + 0.05 * value_noise(x * 9, y * 9, seed, 3)
    alpha = np.clip((edge - rad) * 14.0 + 0.5, 0.0, 1.0)
    vein = 0.5 + 0.5 * np.cos(40 * (cx * np.cos(seed) + cy * np.sin(seed)))
    rgb = np.stack([0.15 + 0.2 * vein, 0.45 + 0.4 * value_noise(x * 5, y * 5, seed + 1, 4) * 0.5 + 0.2 * vein,
                    0.1 + 0.1 * vein], -1)
    img = np.concatenate([np.clip(rgb, 0, 1), alpha[..., None]], -1)
    mips = [img]
    while mips[-1].shape[0] > 1:
        m = mips[-1]
        mips.append(0.25 * (m[0::2, 0::2] + m[1::2, 0::2] + m[0::2, 1::2] + m[1::2, 1::2]))
    chain = [np.round(m * 255.0).astype(np.uint8).reshape(-1) for m in reversed(mips)]  # smallest first
    flags = (abi.VT_TEXFLAG_CLAMPS | abi.VT_TEXFLAG_CLAMPT) if clamp else 0
    return (size, size, len(mips), flags, np.concatenate(chain))


def noise_texture(size=64, seed=1, clamp=False, kind="colour"):
    """Synthetic RGBA8888 texture with a full mip chain in VTF memory order (smallest mip first).
    kind: "colour" (smooth RGBA noise), "normal" (tangent-space normals around +Z), "blend" (r,g in the ranges
    WorldVertexTransition blend modulate textures use)."""
    lin = (np.arange(size) + 0.5) / size
    x, y = np.meshgrid(lin, lin, indexing="xy")
    ch = [value_noise(x * (5 + k), y * (5 + k), seed * 7 + k, 4) * 0.5 + 0.5 for k in range(4)]
    if kind == "normal":
        nx, ny = (ch[0] - 0.5) * 1.2, (ch[1] - 0.5) * 1.2
        nz = np.sqrt(np.clip(1.0 - nx * nx - ny * ny, 0.05, 1.0))
        img = np.stack([nx * 0.5 + 0.5, ny * 0.5 + 0.5, nz * 0.5 + 0.5, ch[3]], -1)
    elif kind == "blend":
        img = np.stack([ch[0] * 0.4, 0.3 + 0.5 * ch[1], ch[2], ch[3]], -1)
    else:
        img = np.stack(ch, -1)
    mips = [np.clip(img, 0, 1)]
    while mips[-1].shape[0] > 1:
        m = mips[-1]
        mips.append(0.25 * (m[0::2, 0::2] + m[1::2, 0::2] + m[0::2, 1::2] + m[1::2, 1::2]))
    chain = [np.round(m * 255.0).astype(np.uint8).reshape(-1) for m in reversed(mips)]
    flags = (abi.VT_TEXFLAG_CLAMPS | abi.VT_TEXFLAG_CLAMPT) if clamp else 0
    return (size, size, len(mips), flags, np.concatenate(chain))


def scene_materials(ground_quads=24, n_props=12, seed=31, extent=100.0):
    """Every TraceResult shading-input branch (source/objects/TraceResult.cpp:89-253): normal maps (one and two,
    blended), WorldVertexTransition blending (smoothstep and masked), second base texture, all detail blend modes of
    TextureCombine (:11-43), MRAO (one and two), UV transforms and texScale, a water material.  Ground triangles
    cycle through the materials and carry random vertex alphas (the blend factor, :72); props reuse them."""
    rng = np.random.default_rng(seed)
    texs = [noise_texture(64, 1), noise_texture(32, 2, clamp=True), noise_texture(32, 3, kind="blend"), noise_texture(64, 4, kind="normal"),
            noise_texture(32, 5, kind="normal", clamp=True), noise_texture(32, 6), noise_texture(16, 7), noise_texture(64, 8)]
    BASE, BASE2, BLEND, NORM, NORM2, MRAO, MRAO2, DETAIL = range(8)
    n_detail_modes = 12  # DetailBlendMode 0..11 (Material.h:12-26); 10 and 11 fall through to the base colour
    mats = abi.default_materials(6 + n_detail_modes)
    mats["surf_flags"][1] = abi.VT_SURF_SKY
    skew = (0.9, 0.2, 0.0, 0.13, -0.15, 1.1, 0.0, 0.37)  # mat2x4: col0 = (a, b, ., tx) drives u, col1 drives v
    m = mats[0]  # blended ground: everything at once
    m["base_texture"], m["base_texture2"], m["blend_texture"] = BASE, BASE2, BLEND
    m["normal_map"], m["normal_map2"], m["mrao"], m["mrao2"] = NORM, NORM2, MRAO, MRAO2
    m["base_tex_mat"], m["base_tex_mat2"], m["normal_map_mat"], m["blend_tex_mat"] = skew, abi._IDENT, skew, abi._IDENT
    m["tex_scale"], m["colour"] = 2.0, (0.9, 0.8, 0.7, 0.95)
    m = mats[2]  # masked blending (blendFactor = blend texture's green, :121-122)
    m["base_texture"], m["base_texture2"], m["blend_texture"], m["masked_blending"] = BASE, BASE2, BLEND, 1
    m = mats[3]  # single normal map + MRAO, no blending
    m["base_texture"], m["normal_map"], m["mrao"], m["normal_map_mat"] = BASE, NORM, MRAO, skew
    m = mats[4]  # masked blending WITHOUT a blend texture (blendFactor = 0.5, :110), two normal maps
    m["base_texture"], m["base_texture2"], m["normal_map"], m["normal_map2"], m["masked_blending"] = BASE, BASE2, NORM, NORM2, 1
    m = mats[5]  # water, no textures at all (fallback base texture)
    m["water"], m["colour"] = 1, (0.2, 0.4, 0.8, 0.5)
    for k in range(n_detail_modes):
        m = mats[6 + k]
        m["base_texture"], m["detail"], m["detail_blend_mode"] = BASE, DETAIL, k
        m["detail_scale"], m["detail_blend_factor"], m["detail_mat"] = 3.0, 0.35 + 0.05 * k, skew
    ground = heightfield(ground_quads, extent, 4.0, seed, 0.05)
    usable = np.array([0, 2, 3, 4, 5] + list(range(6, 6 + n_detail_modes)))
    ground["material"] = usable[np.arange(len(ground)) % len(usable)]
    ground["alphas"] = rng.uniform(0, 1, (len(ground), 3))
    parts = [ground, box([-extent / 2, -extent / 2, -9.0], [extent / 2, extent / 2, 50.0], True, material=1)]
    rots = _random_rotations(rng, n_props)
    for e in range(n_props):
        obj = torus(21, 11, seed=seed + e, material=int(usable[e % len(usable)]), ent_idx=e + 1)
        obj["alphas"] = rng.uniform(0, 1, (len(obj), 3))
        parts.append(transform_tris(obj, rots[e], (rng.uniform(-35, 35), rng.uniform(-35, 35), rng.uniform(6, 25)), rng.uniform(2.0, 5.0)))
    ents = np.zeros(n_props + 1, abi.ENTITY)
    ents["id"] = np.concatenate([[0], 200 + np.arange(n_props)])
    ents["colour"] = 1.0
    ents["colour"][1:] = rng.uniform(0.4, 1.0, (n_props, 4))
    return abi.SceneData(np.concatenate(parts), mats, ents, texs)


# ---------------------------------------------------------------------- scenes
def scene_heightfield(n_quads=224, seed=1234, closed=True, extent=100.0, amp=8.0):
    """Config 1: 2*n_quads^2 terrain triangles (+12 for the enclosing room when closed)."""
    parts = [heightfield(n_quads, extent, amp, seed)]
    if closed:
        parts.append(box([-extent / 2, -extent / 2, -amp - 5.0], [extent / 2, extent / 2, 60.0], True, material=1))
    mats = abi.default_materials(2)
    mats["colour"][1] = (0.6, 0.7, 0.9, 1.0)
    mats["surf_flags"][1] = abi.VT_SURF_SKY
    return abi.SceneData(np.concatenate(parts), mats)


def scene_props(n_props=256, nu=63, nv=31, ground_quads=64, seed=4321, extent=200.0):
    """Config 2: n_props displaced tori (2*nu*nv two-sided triangles each, per-entity rigid
    transform baked to world space), entity 0 = one-sided world (ground + room)."""
    rng = np.random.default_rng(seed)
    parts = [heightfield(ground_quads, extent, 2.0, seed, 0.03), box([-extent / 2, -extent / 2, -8.0], [extent / 2, extent / 2, 80.0], True, material=1)]
    rots = _random_rotations(rng, n_props)
    pos = np.stack([rng.uniform(-extent * 0.45, extent * 0.45, n_props), rng.uniform(-extent * 0.45, extent * 0.45, n_props),
                    rng.uniform(4.0, 40.0, n_props)], -1)
    scl = rng.uniform(2.0, 6.0, n_props)
    for e in range(n_props):
        obj = torus(nu, nv, seed=seed + e, material=2 + (e % 6), ent_idx=e + 1)
        parts.append(transform_tris(obj, rots[e], pos[e], scl[e]))
    mats = abi.default_materials(8)
    mats["surf_flags"][1] = abi.VT_SURF_SKY
    cols = rng.uniform(0.2, 1.0, (6, 3))
    mats["colour"][2:, :3] = cols
    ents = np.zeros(n_props + 1, abi.ENTITY)
    ents["id"] = np.concatenate([[0], 100 + np.arange(n_props)])
    ents["colour"] = 1.0
    ents["colour"][1:, :3] = rng.uniform(0.5, 1.0, (n_props, 3))
    return abi.SceneData(np.concatenate(parts), mats, ents)


def scene_terrain_closed(n_quads=1582, seed=1234, n_props=0, extent=400.0, amp=30.0):
    """Config 3 / 5: large closed scene — displaced terrain inside a room, optional props."""
    parts = [heightfield(n_quads, extent, amp, seed, 0.02), box([-extent / 2, -extent / 2, -amp - 10.0], [extent / 2, extent / 2, 150.0], True, material=1)]
    n_mats = 2
    ents = np.zeros(n_props + 1, abi.ENTITY)
    ents["colour"] = 1.0
    if n_props:
        rng = np.random.default_rng(seed + 1)
        rots = _random_rotations(rng, n_props)
        pos = np.stack([rng.uniform(-extent * 0.45, extent * 0.45, n_props), rng.uniform(-extent * 0.45, extent * 0.45, n_props),
                        rng.uniform(amp, amp + 60.0, n_props)], -1)
        scl = rng.uniform(3.0, 10.0, n_props)
        base = torus(126, 62, seed=seed)
        for e in range(n_props):
            obj = base.copy()
            obj["material"] = 2 + (e % 6)
            obj["ent_idx"] = e + 1
            parts.append(transform_tris(obj, rots[e], pos[e], scl[e]))
        n_mats = 8
        ents["id"][1:] = 100 + np.arange(n_props)
    mats = abi.default_materials(n_mats)
    mats["surf_flags"][1] = abi.VT_SURF_SKY
    return abi.SceneData(np.concatenate(parts), mats, ents)


def scene_foliage(n_cards=20000, seed=99, extent=100.0, tex_size=256, ground_quads=32):
    """Config 4: two-triangle leaf cards with `alphatest | nocull` materials over synthetic VTFs
    (one wrapping, one clamped), identity baseTexMat; plus a one-sided ground and room."""
    rng = np.random.default_rng(seed)
    ctr = np.stack([rng.uniform(-extent * 0.45, extent * 0.45, n_cards), rng.uniform(-extent * 0.45, extent * 0.45, n_cards),
                    rng.uniform(1.0, 30.0, n_cards)], -1)
    rots = _random_rotations(rng, n_cards)
    half = rng.uniform(0.6, 1.8, n_cards)
    ax, ay, an = rots[:, :, 0] * half[:, None], rots[:, :, 1] * half[:, None], rots[:, :, 2]
    c00, c10, c11, c01 = ctr - ax - ay, ctr + ax - ay, ctr + ax + ay, ctr - ax + ay
    uvscale = rng.choice([1.0, 2.0], n_cards)  # some cards tile the texture (exercises the wrap path)
    uvoff = rng.uniform(-1.0, 1.0, (n_cards, 2)) * (uvscale[:, None] > 1)
    uv = lambda a, b: np.stack([a * uvscale + uvoff[:, 0], b * uvscale + uvoff[:, 1]], -1)
    t = np.zeros(2 * n_cards, abi.TRI_IN)
    t["p"][0::2] = np.stack([c00, c10, c11], 1)
    t["p"][1::2] = np.stack([c00, c11, c01], 1)
    t["uvs"][0::2] = np.stack([uv(0, 0), uv(1, 0), uv(1, 1)], 1)
    t["uvs"][1::2] = np.stack([uv(0, 0), uv(1, 1), uv(0, 1)], 1)
    # slightly bent vertex normals so normal interpolation is not trivial
    for k, corner in enumerate((c00, c10, c11)):
        t["normals"][0::2, k] = _normalize(an + 0.3 * _normalize(corner - ctr))
    for k, corner in enumerate((c00, c11, c01)):
        t["normals"][1::2, k] = _normalize(an + 0.3 * _normalize(corner - ctr))
    t["tangents"] = np.repeat(_normalize(ax), 2, 0)[:, None, :]
    t["alphas"] = rng.uniform(0, 1, (2 * n_cards, 3))
    t["material"] = np.repeat(2 + (np.arange(n_cards) % 2), 2)
    t["ent_idx"] = np.repeat(1 + (np.arange(n_cards) % 7), 2)
    t["one_sided"] = 0
    ground = heightfield(ground_quads, extent, 1.5, seed, 0.05)
    room = box([-extent / 2, -extent / 2, -6.0], [extent / 2, extent / 2, 60.0], True, material=1)
    mats = abi.default_materials(4)
    mats["surf_flags"][1] = abi.VT_SURF_SKY
    mats["base_texture"][0] = 0  # ground samples the wrapping texture for albedo only (no alphatest flag)
    for m, tex in ((2, 0), (3, 1)):
        mats["flags"][m] = abi.VT_MATFLAG_ALPHATEST | abi.VT_MATFLAG_NOCULL
        mats["base_texture"][m] = tex
        mats["colour"][m] = (0.9, 1.0, 0.8, 1.0)
    ents = np.zeros(8, abi.ENTITY)
    ents["id"] = [0, 11, 12, 13, 14, 15, 16, 17]
    ents["colour"] = 1.0
    ents["colour"][1:, :3] = rng.uniform(0.6, 1.0, (7, 3))
    texs = [leaf_texture(tex_size, 5, clamp=False), leaf_texture(tex_size, 9, clamp=True)]
    return abi.SceneData(np.concatenate([t, ground, room]), mats, ents, texs)


# ------------------------------------------------------------------------ rays
def pinhole_rays(width, height, eye, look, up=(0, 0, 1), vfov=60.0, tmin=0.0, tmax=np.finfo(f4).max):
    """Pinhole primary rays, pixel-centre sampling as libs/bvh/test/benchmark.cpp:129-150
    (row-major: index = width*j + i), float32 arithmetic."""
    eye = np.asarray(eye, f4)
    d = np.asarray(look, f4) - eye
    d = (d / np.sqrt((d * d).sum(dtype=f4))).astype(f4)
    iu = np.cross(d, np.asarray(up, f4)).astype(f4)
    iu = (iu / np.sqrt((iu * iu).sum(dtype=f4))).astype(f4)
    iv = np.cross(iu, d).astype(f4)
    iv = (iv / np.sqrt((iv * iv).sum(dtype=f4))).astype(f4)
    w = f4(np.tan(f4(vfov) * f4(np.pi / 180.0 * 0.5)))
    ratio = f4(height) / f4(width)
    iu = iu * w
    iv = iv * w * ratio
    i = np.arange(width, dtype=f4)
    j = np.arange(height, dtype=f4)
    u = (f4(2) * (i + f4(0.5)) / f4(width) - f4(1)).astype(f4)
    v = (f4(2) * (j + f4(0.5)) / f4(height) - f4(1)).astype(f4)
    dirs = (iu[None, None, :] * u[None, :, None] + iv[None, None, :] * v[:, None, None] + d[None, None, :]).astype(f4)
    dirs = dirs.reshape(-1, 3)
    dirs = (dirs / np.sqrt((dirs * dirs).sum(-1, keepdims=True, dtype=f4))).astype(f4)
    rays = np.zeros(width * height, abi.RAY)
    rays["o"] = eye
    rays["d"] = dirs
    rays["tmin"] = tmin
    rays["tmax"] = tmax
    return rays


def calc_ray_origin(pos, normal):
    """vistrace.CalcRayOrigin (source/VisTrace.cpp:1495-1517): offset a hit position along the
    normal by an integer ulp step (or a fixed step near the origin)."""
    pos = np.asarray(pos, f4)
    normal = np.asarray(normal, f4)
    origin, f_scale, i_scale = f4(1.0 / 32.0), f4(1.0 / 65536.0), f4(256.0)
    i_off = (normal * i_scale).astype(np.int32)  # truncation toward zero like glm::ivec3(vec3)
    i_pos = (pos.view(np.int32) + np.where(pos < 0, -i_off, i_off)).view(f4)
    return np.where(np.abs(pos) < origin, pos + normal * f_scale, i_pos).astype(f4)


def _uniform01(idx, dim, key):
    """Counter-based RNG: hash(ray index, dimension, key) -> [0,1) float32."""
    return _hash01(idx.astype(np.int64), np.full(idx.shape, dim, np.int64), key).astype(f4)


def bounce_rays(attrs, spp=1, key=3, tmax=np.finfo(f4).max):
    """Cosine-weighted diffuse bounce rays about the shading normal of each hit using
    hemisphere_cos (source/libraries/BSDF.cpp:69-77: z = sqrt(r1), sinTheta = sqrt(1-r1),
    phi = 2*pi*r2) in the hit's TBN frame, origin = CalcRayOrigin(pos, geometric normal
    facing the viewer).  Returns (rays, parent index)."""
    hit = np.nonzero(attrs["prim"] != abi.VT_MISS)[0]
    hit = hit[(attrs["flags"][hit] & abi.VT_ATTR_HIT_SKY) == 0]
    parent = np.repeat(hit, spp)
    sample = np.tile(np.arange(spp), len(hit))
    ctr = parent.astype(np.int64) * spp + sample
    r1, r2 = _uniform01(ctr, 0, key), _uniform01(ctr, 1, key)
    z = np.sqrt(r1)
    st = np.sqrt(f4(1.0) - r1)
    phi = f4(2.0 * np.pi) * r2
    lx, ly = st * np.cos(phi), st * np.sin(phi)
    a = attrs[parent]
    front = ((a["flags"] & abi.VT_ATTR_FRONT_FACING) != 0)[:, None]
    sgn = np.where(front, f4(1), f4(-1)).astype(f4)
    n, t, b = a["normal"] * sgn, a["tangent"], a["binormal"] * sgn
    d = (t * lx[:, None] + b * ly[:, None] + n * z[:, None]).astype(f4)
    gn = a["geometric_normal"] * sgn
    rays = np.zeros(len(parent), abi.RAY)
    rays["o"] = calc_ray_origin(a["pos"], gn)
    rays["d"] = d
    rays["tmin"] = 0.0
    rays["tmax"] = tmax
    ok = np.isfinite(d).all(-1) & ((d * d).sum(-1) > 0)
    return rays[ok], parent[ok]


def shadow_rays(attrs, sun_dir=(0.3, 0.2, 0.93), tmax=np.finfo(f4).max):
    """One shadow ray per (non-sky) hit toward a fixed sun direction (config 2)."""
    hit = np.nonzero((attrs["prim"] != abi.VT_MISS) & ((attrs["flags"] & abi.VT_ATTR_HIT_SKY) == 0))[0]
    a = attrs[hit]
    s = np.asarray(sun_dir, f4)
    s = (s / np.sqrt((s * s).sum(dtype=f4))).astype(f4)
    front = ((a["flags"] & abi.VT_ATTR_FRONT_FACING) != 0)[:, None]
    gn = a["geometric_normal"] * np.where(front, f4(1), f4(-1)).astype(f4)
    rays = np.zeros(len(hit), abi.RAY)
    rays["o"] = calc_ray_origin(a["pos"], gn)
    rays["d"] = s
    rays["tmin"] = 0.0
    rays["tmax"] = tmax
    return rays, hit


def random_rays(n, lo, hi, seed=7, tmax=np.finfo(f4).max):
    """Fully incoherent rays: origins uniform in a box, directions uniform on the sphere."""
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, abi.RAY)
    rays["o"] = rng.uniform(lo, hi, (n, 3)).astype(f4)
    rays["d"] = _normalize(rng.normal(size=(n, 3))).astype(f4)
    rays["tmin"] = 0.0
    rays["tmax"] = tmax
    return rays
