"""CPU tests of the VTF ingestion (vt_vtf_decode) against the reference's own parser.

Synthetic VTF files are assembled here byte by byte (header per libs/VTFParser/FileFormat/Structs.h:21-75; any 8/16 bytes
are a valid DXT block, so random payloads cover every decoder branch); the checker is the reference's VTFTexture
(oracle/_ref: constructor incl. DXT decompression + GetPixel).  The bar is bit-exact: byte / 255.f == the reference's float.
A committed golden fixture (tests/golden/vtf_small.npz, written by make_golden.py from the reference) pins the same
property where oracle/_ref is not available."""
import os
import struct

import numpy as np
import pytest

FMT = {"RGBA8888": 0, "ABGR8888": 1, "RGB888": 2, "BGR888": 3, "RGB565": 4, "I8": 5, "IA88": 6, "P8": 7, "A8": 8,
       "RGB888_BLUESCREEN": 9, "BGR888_BLUESCREEN": 10, "ARGB8888": 11, "BGRA8888": 12, "DXT1": 13, "DXT3": 14, "DXT5": 15,
       "BGRX8888": 16, "BGR565": 17, "BGRX5551": 18, "BGRA4444": 19, "DXT1_ONEBITALPHA": 20, "BGRA5551": 21, "UV88": 22,
       "UVWQ8888": 23, "RGBA16161616F": 24, "RGBA16161616": 25, "UVLX8888": 26}
BPP = {0: 4, 1: 4, 2: 3, 3: 3, 4: 2, 5: 1, 6: 2, 7: 1, 8: 1, 9: 3, 10: 3, 11: 4, 12: 4, 16: 4, 17: 2, 18: 2, 19: 2, 21: 2, 22: 2, 23: 4, 24: 8, 25: 8, 26: 4}
SUPPORTED = ["RGBA8888", "ABGR8888", "RGB888", "BGR888", "I8", "IA88", "A8", "RGB888_BLUESCREEN", "BGR888_BLUESCREEN", "ARGB8888",
             "BGRA8888", "DXT1", "DXT3", "DXT5", "BGRX8888", "DXT1_ONEBITALPHA", "UV88", "UVWQ8888", "UVLX8888"]
SUPPORTED.append("P8")  # no case in ParsePixel: reads as VTFPixel{} = opaque black
# the 16-bit formats decode to WIDE texels: four uint16 numerators + a divisor code per channel (vt_texture.texel_layout)
WIDE = {"RGB565": (255, 255, 255, 255), "BGR565": (255, 255, 255, 255), "BGRX5551": (255, 255, 255, 1), "BGRA5551": (255, 255, 255, 1),
        "BGRA4444": (255, 255, 255, 255), "RGBA16161616F": (65535,) * 4, "RGBA16161616": (65535,) * 4}


def texels_as_floats(rgba, layout):
    """The channel values the device (and the C port) form from a decoded chain: byte / 255.f, or numerator / divisor."""
    if layout == 0:
        return rgba.reshape(-1, 4).astype(np.float32) / np.float32(255.0)
    assert layout & 0x100
    div = np.array([{0: 255.0, 1: 65535.0, 2: 1.0}[(layout >> (2 * c)) & 3] for c in range(4)], np.float32)
    return rgba.view(np.uint16).reshape(-1, 4).astype(np.float32) / div[None, :]


def image_size(w, h, d, f):
    if f in (13, 20):
        return ((max(w, 4) + 3) // 4) * ((max(h, 4) + 3) // 4) * 8 * d
    if f in (14, 15):
        return ((max(w, 4) + 3) // 4) * ((max(h, 4) + 3) // 4) * 16 * d
    return w * h * d * BPP[f]


def make_vtf(fmt, w, h, mips, frames=1, flags=0, minor=2, low=None, resources=False, seed=0, first_frame=0, depth=1):
    """One synthetic VTF file: random payload, optional DXT1 thumbnail (`low` = (w, h)), optional 7.3+ resource dictionary."""
    rng = np.random.default_rng(seed)
    f = FMT[fmt]
    faces = 1 if not flags & 0x4000 else (7 if first_frame != 0xFFFF and minor < 5 else 6)
    n = sum(image_size(max(1, w >> m), max(1, h >> m), max(1, depth >> m), f) for m in range(mips)) * frames * faces
    payload = rng.integers(0, 256, n, dtype=np.uint8)
    if fmt.startswith("DXT"):  # make both endpoint orders and both alpha modes common
        blk = 8 if f in (13, 20) else 16
        p = payload.reshape(-1, blk)
        swap = rng.random(len(p)) < 0.5
        c = p[:, blk - 8: blk - 4].copy().view(np.uint16)
        c[swap] = np.sort(c[swap], axis=1)
        p[:, blk - 8: blk - 4] = c.view(np.uint8)
        payload = p.reshape(-1)
    low_bytes = b""
    low_fmt, low_w, low_h = -1, 0, 0
    if low:
        low_fmt, (low_w, low_h) = 13, low
        low_bytes = rng.integers(0, 256, image_size(low_w, low_h, 1, 13), dtype=np.uint8).tobytes()
    n_res = 2 if resources else 0
    header_size = 80 + 8 * n_res if minor >= 3 else (80 if minor == 2 else 64)
    hdr = bytearray(80 + 8 * n_res)
    struct.pack_into("<4sIII", hdr, 0, b"VTF\0", 7, minor, header_size)
    struct.pack_into("<HHIHH", hdr, 16, w, h, flags, frames, first_frame)
    struct.pack_into("<3f", hdr, 32, 0.5, 0.5, 0.5)
    struct.pack_into("<f", hdr, 48, 1.0)
    struct.pack_into("<iBiBB", hdr, 52, f, mips, low_fmt, low_w, low_h)
    struct.pack_into("<H", hdr, 63, depth)
    struct.pack_into("<I", hdr, 68, n_res)
    body = bytes(hdr[:header_size])
    if resources:  # low-res thumbnail resource (tag 0x01) first, then the image (tag 0x30): offsets are absolute
        off_low = header_size
        off_img = header_size + len(low_bytes) + 24  # a gap the parser must skip by offset, not by position
        hb = bytearray(body)
        struct.pack_into("<3sBI", hb, 80, b"\x01\0\0", 0, off_low)
        struct.pack_into("<3sBI", hb, 88, b"\x30\0\0", 0, off_img)
        return bytes(hb) + low_bytes + bytes(24) + payload.tobytes()
    return body + low_bytes + payload.tobytes()


CASES = [(fmt, 16, 8, 5, {}) for fmt in SUPPORTED] + [
    ("DXT1", 64, 64, 7, {"low": (16, 16)}),
    ("DXT5", 32, 64, 7, {"low": (8, 16), "seed": 3}),
    ("DXT3", 20, 12, 3, {"seed": 4}),                                     # not a multiple of 4: partial blocks
    ("DXT5", 6, 10, 4, {"seed": 5}),
    ("DXT1_ONEBITALPHA", 4, 4, 3, {"seed": 6}),                           # mips below one block
    ("BGRA8888", 8, 8, 4, {"frames": 3, "frame": 2}),
    ("DXT5", 16, 16, 5, {"frames": 2, "frame": 1, "seed": 7}),
    ("BGR888", 8, 8, 4, {"flags": 0x4000, "face": 4}),                    # environment map, 7.2: 7 faces
    ("DXT1", 8, 8, 4, {"flags": 0x4000, "minor": 5, "resources": True, "face": 5, "low": (8, 8)}),  # 7.5: 6 faces, resource dictionary
    ("RGBA8888", 16, 16, 1, {"minor": 4, "resources": True, "flags": 0x4 | 0x8}),
    ("I8", 8, 4, 4, {"minor": 1}),                                        # 7.1: 64-byte header, no depth field
    ("DXT5", 16, 16, 5, {"minor": 3, "resources": True, "low": (16, 16), "seed": 9}),
]


def _ids():
    return [f"{c[0]}-{c[1]}x{c[2]}-{'-'.join(f'{k}{v}' for k, v in c[4].items())}" for c in CASES]


@pytest.mark.parametrize("case", CASES, ids=_ids())
def test_vtf_decode_matches_the_reference_parser(built, oracle_mod, case):
    import vistrace_b200 as vt

    if not oracle_mod.available("reference"):
        pytest.skip("needs oracle/_ref (the reference's VTFTexture); the golden fixture covers this elsewhere")
    fmt, w, h, mips, kw = case
    kw = dict(kw)
    frame, face = kw.pop("frame", 0), kw.pop("face", 0)
    data = make_vtf(fmt, w, h, mips, **kw)
    info = vt.vtf_info(data)
    assert (info["width"], info["height"], info["mip_count"], info["format"], info["supported"]) == (w, h, mips, FMT[fmt], 1)
    gw, gh, gm, gflags, rgba, layout = vt.vtf_decode(data, frame, face)
    n_tex = sum(max(1, w >> m) * max(1, h >> m) for m in range(mips))
    assert layout == 0 and len(rgba) == 4 * n_tex == info["rgba_bytes"] and gflags == kw.get("flags", 0)
    want = oracle_mod.vtf_pixels(data, n_tex, frame, face)
    assert want is not None, "the reference parser rejected the synthetic file"
    got = texels_as_floats(rgba, layout)
    assert got.tobytes() == want.tobytes()


def test_vtf_decode_random_files(built, oracle_mod):
    """Seeded fuzz: every supported format over odd sizes (1 .. 69 texels a side), partial mip chains, frames, environment-map
    faces, thumbnails and resource dictionaries — each file decodes to exactly the reference's texels."""
    import vistrace_b200 as vt

    if not oracle_mod.available("reference"):
        pytest.skip("needs oracle/_ref")
    rng = np.random.default_rng(7)
    for it in range(150):
        fmt = SUPPORTED[it % len(SUPPORTED)]
        w, h = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        if it % 3 == 0:
            w, h = 1 << int(rng.integers(0, 7)), 1 << int(rng.integers(0, 7))
        mips = int(rng.integers(1, int(np.floor(np.log2(max(w, h)))) + 2))
        kw = {"seed": it}
        if it % 5 == 0:
            kw["frames"] = int(rng.integers(1, 4))
        if it % 7 == 0:
            kw["low"] = (16, 16)
        if it % 11 == 0:
            kw.update(minor=int(rng.integers(3, 6)), resources=True)
        if it % 13 == 0:
            kw["flags"] = 0x4000
        data = make_vtf(fmt, w, h, mips, **kw)
        frame = int(rng.integers(0, kw.get("frames", 1)))
        faces = 1 if not kw.get("flags", 0) & 0x4000 else (7 if kw.get("minor", 2) < 5 else 6)
        face = int(rng.integers(0, faces))
        n_tex = sum(max(1, w >> m) * max(1, h >> m) for m in range(mips))
        want = oracle_mod.vtf_pixels(data, n_tex, frame, face)
        assert want is not None, (it, fmt, w, h, mips, kw)
        dec = vt.vtf_decode(data, frame, face)
        got = texels_as_floats(dec[4], dec[5])
        assert got.tobytes() == want.tobytes(), (it, fmt, w, h, mips, kw)


@pytest.mark.parametrize("fmt", sorted(WIDE))
def test_vtf_16_bit_formats_decode_to_wide_texels(built, oracle_mod, fmt):
    """RGB565, BGR565, BGRX5551, BGRA5551, BGRA4444, RGBA16161616(F): the reference's ParsePixel leaves [0, 1] for them (unmasked
    shifts: green of an RGB565 texel reaches 32.1) or divides by 65535, which no RGBA8888 texel can hold; the decoder emits the
    integer numerators and a divisor per channel instead, and numerator / divisor equals VTFTexture::GetPixel bit for bit."""
    import vistrace_b200 as vt

    for seed, (w, h, mips) in enumerate([(8, 8, 2), (16, 4, 5), (5, 9, 3), (64, 64, 7)]):
        data = make_vtf(fmt, w, h, mips, seed=seed, flags=0x4 if seed == 1 else 0)
        info = vt.vtf_info(data)
        want_layout = 0x100 | sum({255: 0, 65535: 1, 1: 2}[d] << (2 * c) for c, d in enumerate(WIDE[fmt]))
        assert (int(info["supported"]), int(info["format"]), int(info["texel_layout"])) == (1, FMT[fmt], want_layout)
        gw, gh, gm, gflags, rgba, layout = vt.vtf_decode(data)
        n_tex = sum(max(1, w >> m) * max(1, h >> m) for m in range(mips))
        assert layout == want_layout and len(rgba) == 8 * n_tex == info["rgba_bytes"]
        got = texels_as_floats(rgba, layout)
        if fmt in ("RGB565", "BGR565"):
            assert got[:, 1].max() > 1.0  # the reference's unmasked green: out of range on purpose
        if oracle_mod.available("reference"):
            want = oracle_mod.vtf_pixels(data, n_tex)
            assert want is not None and got.tobytes() == want.tobytes(), (fmt, w, h, mips)


def test_vtf_malformed_files_fail_like_the_reference(built, oracle_mod):
    import vistrace_b200 as vt

    good = make_vtf("DXT5", 16, 16, 5, low=(16, 16))
    bad = [good[:40], b"VTX\0" + good[4:], good[:4] + struct.pack("<II", 7, 6) + good[12:], good[:-1],
           good[:52] + struct.pack("<i", -1) + good[56:], good[:12] + struct.pack("<I", 4000) + good[16:]]
    for b in bad:
        with pytest.raises(RuntimeError):
            vt.vtf_decode(b)
        if oracle_mod.available("reference"):
            assert oracle_mod.vtf_pixels(b, 341) is None
    with pytest.raises(RuntimeError):
        vt.vtf_decode(good, frame=1)


def test_vtf_golden_fixture(built):
    """Files + the reference's texels committed by tests/golden/make_golden.py: pins the decoder where oracle/_ref is absent."""
    import vistrace_b200 as vt

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "vtf_small.npz"))
    names = sorted(k[5:] for k in g.files if k.startswith("file_"))
    assert len(names) >= 6
    for name in names:
        _, _, _, _, rgba, layout = vt.vtf_decode(g["file_" + name].tobytes())
        got = texels_as_floats(rgba, layout)
        assert got.tobytes() == g["want_" + name].tobytes(), name
