"""CPU tests of the host C mirror (c-ray_b200/host): tile quantisation and queue (reference
src/datatypes/tile.c), and the BMP/PNG encoders (reference src/utils/encoders)."""
import ctypes as C
import os
import struct
import zlib

import numpy as np
import pytest

from conftest import ROOT, GOLDEN


class IntCoord(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int)]


class RenderTile(C.Structure):                      # struct renderTile, tile.h:28-37
    _fields_ = [("width", C.c_uint), ("height", C.c_uint), ("begin", IntCoord), ("end", IntCoord),
                ("isRendering", C.c_bool), ("renderComplete", C.c_bool), ("networkRenderer", C.c_bool), ("tileNum", C.c_int)]


class Texture8(C.Structure):
    _fields_ = [("width", C.c_uint), ("height", C.c_uint), ("data", C.POINTER(C.c_uint8))]


@pytest.fixture(scope="module")
def host():
    L = C.CDLL(os.path.join(ROOT, "c-ray_b200", "libcrhost.so"))
    L.quantizeImage.argtypes = [C.POINTER(C.POINTER(RenderTile))] + [C.c_uint] * 4 + [C.c_int]
    L.quantizeImage.restype = C.c_uint
    L.newRenderer.restype = C.c_void_p
    L.loadSceneFile.argtypes = [C.c_void_p, C.c_char_p] + [C.c_int] * 4
    L.nextTile.argtypes = [C.c_void_p]
    L.nextTile.restype = RenderTile
    L.destroyRenderer.argtypes = [C.c_void_p]
    L.writeImage.argtypes = [C.POINTER(Texture8), C.c_char_p, C.c_int]
    return L


def tiles(host, W, H, tw, th, order):
    p = C.POINTER(RenderTile)()
    n = host.quantizeImage(C.byref(p), W, H, tw, th, order)
    return [(p[i].begin.x, p[i].begin.y, p[i].end.x, p[i].end.y, p[i].width, p[i].height, p[i].tileNum) for i in range(n)]


@pytest.mark.parametrize("W,H,tw,th", [(1920, 1080, 64, 64), (320, 200, 320, 200), (100, 70, 32, 16), (48, 32, 16, 16), (7, 5, 64, 64), (33, 1, 8, 8)])
@pytest.mark.parametrize("order", [0, 1, 2, 3, 4])
def test_quantize_covers_every_pixel_once(host, W, H, tw, th, order):
    ts = tiles(host, W, H, tw, th, order)
    cover = np.zeros((H, W), dtype=np.int32)
    for x0, y0, x1, y1, w, h, _ in ts:
        assert 0 <= x0 < x1 <= W and 0 <= y0 < y1 <= H and w == x1 - x0 and h == y1 - y0
        assert w <= min(tw, W) and h <= min(th, H)
        cover[y0:y1, x0:x1] += 1
    assert (cover == 1).all()
    tx, ty = -(-W // min(tw, W)), -(-H // min(th, H))
    assert len(ts) == tx * ty                                   # tile.c:76-80
    assert sorted(t[6] for t in ts) == list(range(len(ts)))     # a permutation of the row-major grid


def test_tile_orders_match_reference_sequences(host):
    # 6 tiles in a row; expected index sequences from tile.c:119-228
    seq = lambda order: [t[6] for t in tiles(host, 96, 16, 16, 16, order)]
    assert seq(3) == [0, 1, 2, 3, 4, 5]                 # normal
    assert seq(0) == [5, 4, 3, 2, 1, 0]                 # topToBottom: reversed
    assert seq(1) == [3, 2, 4, 1, 5, 0]                 # fromMiddle: right of the middle first, alternating
    assert seq(2) == [5, 0, 4, 1, 3, 2]                 # toMiddle: last, first, ...
    odd = [t[6] for t in tiles(host, 80, 16, 16, 16, 1)]
    assert odd == [2, 1, 3, 0, 4]
    rnd = seq(4)
    assert sorted(rnd) == [0, 1, 2, 3, 4, 5] and rnd == seq(4)  # deterministic shuffle (PCG seeded 3141592)


def test_next_tile_hands_out_each_tile_once(host):
    r = host.newRenderer()
    assert host.loadSceneFile(r, os.path.join(GOLDEN, "g_legacy.crscene").encode(), 0, 0, 0, 0) == 0
    seen = []
    while True:
        t = host.nextTile(r)
        if t.tileNum == -1:
            break
        seen.append((t.tileNum, t.begin.x, t.begin.y, t.end.x, t.end.y))
    assert [s[0] for s in seen] == list(range(6))          # 48x32 in 16x16 tiles, queue positions 0..5
    assert host.nextTile(r).tileNum == -1                  # exhausted stays exhausted (tile.c:27-31)
    assert len({s[1:] for s in seen}) == 6
    host.destroyRenderer(r)


def decode_png(path):
    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, W, H = 8, b"", 0, 0
    while pos < len(b):
        n, tag = struct.unpack(">I4s", b[pos:pos + 8])
        data = b[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", b[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + data) & 0xFFFFFFFF
        if tag == b"IHDR":
            W, H, depth, ctype = struct.unpack(">IIBB", data[:10])
            assert (depth, ctype) == (8, 2)
        elif tag == b"IDAT":
            idat += data
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(H, W * 3 + 1)
    assert (raw[:, 0] == 0).all()
    return raw[:, 1:].reshape(H, W, 3)


def test_encoders_roundtrip(host, tmp_path):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    t = Texture8(53, 37, img.ctypes.data_as(C.POINTER(C.c_uint8)))
    p = str(tmp_path / "a.png").encode()
    assert host.writeImage(C.byref(t), p, 1) == 0
    assert np.array_equal(decode_png(p.decode()), img)
    b = str(tmp_path / "a.bmp").encode()
    assert host.writeImage(C.byref(t), b, 0) == 0
    raw = open(b.decode(), "rb").read()
    assert raw[:2] == b"BM" and struct.unpack("<I", raw[2:6])[0] == len(raw)
    W, H = struct.unpack("<ii", raw[18:26])
    row = (W * 3 + 3) & ~3
    px = np.frombuffer(raw[54:], dtype=np.uint8).reshape(H, row)[:, :W * 3].reshape(H, W, 3)
    assert np.array_equal(px[::-1, :, ::-1], img)       # bottom-up, BGR (bmp.c:19-71)


def test_png_carries_the_reference_text_labels(host, tmp_path):
    """png.c:29-56: the reference labels its PNGs with uncompressed tEXt chunks (version, source, samples, bounces, render time,
    threads, uname); same keywords here, placed before the image data, and our own decoder still reads the file."""
    class Info(C.Structure):
        _fields_ = [("samples", C.c_int), ("bounces", C.c_int), ("threadCount", C.c_int), ("renderSeconds", C.c_double)]
    img = np.arange(5 * 7 * 3, dtype=np.uint8).reshape(5, 7, 3)
    t = Texture8(7, 5, img.ctypes.data_as(C.POINTER(C.c_uint8)))
    info = Info(1000, 32, 8, 83.4)
    p = str(tmp_path / "t.png")
    host.writeImageInfo.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
    assert host.writeImageInfo(C.byref(t), p.encode(), 1, C.byref(info)) == 0
    b = open(p, "rb").read()
    texts, pos, order = {}, 8, []
    while pos < len(b):
        n, tag = struct.unpack(">I4s", b[pos:pos + 8])
        data = b[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", b[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + data) & 0xFFFFFFFF
        order.append(tag)
        if tag == b"tEXt":
            k, v = data.split(b"\0", 1)
            texts[k.decode()] = v.decode()
        pos += 12 + n
    assert [k for k in texts] == ["C-ray Version", "C-ray Source", "C-ray Samples", "C-ray Bounces", "C-ray RenderTime", "C-ray Threads", "C-ray SysInfo"]
    assert (texts["C-ray Samples"], texts["C-ray Bounces"], texts["C-ray Threads"], texts["C-ray RenderTime"]) == ("1000", "32", "8", "1m 23s")
    assert texts["C-ray Source"] == "https://github.com/vkoskiv/c-ray"
    assert order[0] == b"IHDR" and order.index(b"IDAT") > max(i for i, t_ in enumerate(order) if t_ == b"tEXt") and order[-1] == b"IEND"
    assert np.array_equal(decode_png(p), img)


def test_png_encoder_multi_band_stream(host, tmp_path):
    """Images above ~1 MB of scanlines are deflated in parallel bands (one IDAT each); the concatenation must be ONE valid
    zlib stream (header, byte-aligned blocks, combined Adler-32) that any decoder accepts."""
    rng = np.random.default_rng(11)
    H, W = 1100, 700                                            # 2.3 MB of scanlines -> 3 bands
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.clip(((xx * 255 // W)[..., None] + rng.integers(-9, 10, (H, W, 3))), 0, 255).astype(np.uint8)
    t = Texture8(W, H, img.ctypes.data_as(C.POINTER(C.c_uint8)))
    p = str(tmp_path / "big.png")
    assert host.writeImage(C.byref(t), p.encode(), 1) == 0
    raw = open(p, "rb").read()
    assert raw.count(b"IDAT") >= 3
    assert np.array_equal(decode_png(p), img)                    # zlib.decompress checks the Adler-32 trailer
    # and through this repository's own PNG decoder (the scene loader's)
    import crscene, json
    sc = tmp_path / "s.json"
    sc.write_text(json.dumps({"renderer": {"width": 4, "height": 4}, "camera": {}, "scene": {"primitives": [
        {"type": "sphere", "radius": 1, "instances": [{}], "material": {"type": "diffuse", "color": {"type": "image", "path": p}}}]}}))
    s_ = crscene.load_json(str(sc))
    A = crscene.arrays(s_)
    tex = A["textures"][0]
    assert (tex["width"], tex["height"], tex["channels"]) == (W, H, 3)
    assert np.array_equal(A["texdata"][int(tex["data_offset"]):int(tex["data_offset"]) + H * W * 3].reshape(H, W, 3), img)
    crscene.free(s_)


def test_nccl_gather_library_is_loaded_on_demand():
    """libcrhost.so must not map NCCL by itself (a Python host with torch carries its own); the gather library is dlopen'ed
    by the first multi-GPU frame.  Checked in a subprocess so that this test process never maps the system NCCL."""
    import subprocess
    import sys
    code = ("import ctypes as C, sys\n"
            "L = C.CDLL(sys.argv[1])\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libnccl' not in maps and 'libcrgpu_nccl' not in maps, 'NCCL mapped by loading libcrhost.so'\n"
            "assert L.crhost_load_nccl() == 0\n"
            "assert 'libcrgpu_nccl' in open('/proc/self/maps').read()\n"
            "print('LAZY-NCCL-OK')\n")
    r = subprocess.run([sys.executable, "-c", code, os.path.join(ROOT, "c-ray_b200", "libcrhost.so")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=120)
    assert "LAZY-NCCL-OK" in r.stdout, r.stdout[-2000:]
