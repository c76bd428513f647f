"""Frame ingest (csrc/ingest.cu, betapose_b200/ingest.py) against the two decoders the reference
uses for the same files: PIL.Image.open (dataloader.py:162) and cv2.imread (yolo/preprocess.py:41).  Host-only code, so
these run without a GPU.  Byte-exact: PNG is lossless, there is no tolerance."""
import io
import os
import struct
import zlib

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from PIL import Image

from betapose_b200 import _lib
from betapose_b200.ingest import FrameIngest

H, W = 48, 64


def _chunk(kind: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xFFFFFFFF)


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def write_png(rows: np.ndarray, ctype: int, depth: int, filters, idat_split: int = 0, palette: bytes | None = None,
              interlace: int = 0, extra: bytes = b"") -> bytes:
    """Minimal PNG writer for the tests.  rows: uint8 [H, row_bytes] packed scan lines (already big-endian for 16 bit);
    filters: one filter type (0..4) per row, applied here so every un-filter branch of the decoder is reached."""
    h, row_bytes = rows.shape
    channels = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    bpp = max(1, channels * depth // 8)
    width = row_bytes * 8 // (channels * depth)
    raw = bytearray()
    prev = np.zeros(row_bytes, np.int32)
    for y in range(h):
        cur = rows[y].astype(np.int32)
        f = filters[y % len(filters)]
        left = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]]) if row_bytes > bpp else np.zeros(row_bytes, np.int32)
        left = left[:row_bytes]
        ul = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])[:row_bytes] if row_bytes > bpp else np.zeros(row_bytes, np.int32)
        if f == 0:
            out = cur
        elif f == 1:
            out = cur - left
        elif f == 2:
            out = cur - prev
        elif f == 3:
            out = cur - ((left + prev) >> 1)
        else:
            pred = np.array([_paeth(int(a), int(b), int(c)) for a, b, c in zip(left, prev, ul)], np.int32)
            out = cur - pred
        raw.append(f)
        raw += (out & 0xFF).astype(np.uint8).tobytes()
        prev = cur
    comp = zlib.compress(bytes(raw), 6)
    png = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", width, h, depth, ctype, 0, 0, interlace))
    png += extra
    if palette is not None:
        png += _chunk(b"PLTE", palette)
    if idat_split:
        for i in range(0, len(comp), idat_split):
            png += _chunk(b"IDAT", comp[i: i + idat_split])
    else:
        png += _chunk(b"IDAT", comp)
    return png + _chunk(b"IEND", b"")


def refs(png: bytes):
    rgb = np.asarray(Image.open(io.BytesIO(png)).convert("RGB"))
    bgr = cv2.imdecode(np.frombuffer(png, np.uint8), cv2.IMREAD_COLOR)
    return rgb, bgr


@pytest.fixture(scope="module")
def ing():
    with FrameIngest(4, H, W) as g:
        yield g


@pytest.fixture(scope="module")
def ing_bgr():
    with FrameIngest(2, H, W, order="bgr") as g:
        yield g


def _smooth(rng, c):
    """Image with structure (so that the encoders' filter heuristics pick every filter type)."""
    y, x = np.mgrid[0:H, 0:W]
    base = np.stack([(x * (3 + k) + y * (5 - k)) % 256 for k in range(c)], -1)
    return ((base + rng.integers(0, 24, (H, W, c))) % 256).astype(np.uint8)


@pytest.mark.parametrize("ctype,depth", [(2, 8), (6, 8), (0, 8), (4, 8), (2, 16), (6, 16), (0, 16), (4, 16)])
def test_every_filter_type_per_colour_type(ing, ing_bgr, ctype, depth):
    rng = np.random.default_rng(ctype * 100 + depth)
    channels = {0: 1, 2: 3, 4: 2, 6: 4}[ctype]
    rows = _smooth(rng, channels * depth // 8).reshape(H, -1)
    for filters in ([0], [1], [2], [3], [4], [4, 3, 2, 1, 0], [3, 4]):
        png = write_png(rows, ctype, depth, filters, idat_split=97)
        rgb, bgr = refs(png)
        if depth == 8 and ctype in (2, 6):  # the writer itself is right: Pillow reads back what was put in
            assert np.array_equal(rgb, rows.reshape(H, W, channels)[:, :, :3])
        got = ing.decode_bytes(png)
        assert np.array_equal(got[:, :, ::-1], bgr), (ctype, depth, filters)  # cv2: 16-bit samples -> high byte
        if (ctype, depth) != (0, 16):  # Pillow opens 16-bit grey as "I;16" and convert("RGB") saturates it: no reference there
            assert np.array_equal(got, rgb), (ctype, depth, filters)
        assert np.array_equal(ing_bgr.decode_bytes(png), got[:, :, ::-1])
        info = FrameIngest.png_info(png)
        assert (info["H"], info["W"], info["channels"], info["depth"]) == (H, W, channels, depth)


@pytest.mark.parametrize("depth", [1, 2, 4, 8])
def test_low_bit_depth_grey_and_palette(ing, depth):
    rng = np.random.default_rng(depth)
    idx = rng.integers(0, 1 << depth, (H, W)).astype(np.uint8)
    per = 8 // depth
    packed = np.zeros((H, W // per), np.uint8)
    for k in range(per):
        packed |= (idx[:, k::per] << ((per - 1 - k) * depth)).astype(np.uint8)
    pal = rng.integers(0, 256, (1 << depth) * 3, dtype=np.uint8).tobytes()
    for ctype, palette in ((0, None), (3, pal)):
        png = write_png(packed, ctype, depth, [0, 1, 2, 3, 4], palette=palette)
        rgb, bgr = refs(png)
        got = ing.decode_bytes(png)
        assert np.array_equal(got, rgb), (ctype, depth)
        assert np.array_equal(got[:, :, ::-1], bgr)


def test_encoders_of_the_reference_stack(ing, tmp_path):
    """Files written by Pillow and by OpenCV (all compression levels / libpng's adaptive filters), decoded from disk."""
    rng = np.random.default_rng(0)
    paths, want = [], []
    for i in range(12):
        im = _smooth(rng, 3)
        p = str(tmp_path / f"f{i:02d}.png")
        if i % 2:
            Image.fromarray(im).save(p, compress_level=i % 10, optimize=bool(i % 3))
        else:
            cv2.imwrite(p, im[:, :, ::-1], [cv2.IMWRITE_PNG_COMPRESSION, i % 10])
        paths.append(p)
        want.append(im)
    got = ing.decode_files(paths)
    assert np.array_equal(got, np.stack(want))
    rgba = np.concatenate([want[0], rng.integers(0, 256, (H, W, 1), dtype=np.uint8)], -1)
    p = str(tmp_path / "rgba.png")
    Image.fromarray(rgba, "RGBA").save(p)
    assert np.array_equal(ing.decode_files([p])[0], want[0])  # alpha dropped, not composited
    assert np.array_equal(ing.decode_files([p])[0], cv2.imread(p)[:, :, ::-1])


def test_ancillary_chunks_are_skipped(ing):
    rows = _smooth(np.random.default_rng(3), 3).reshape(H, -1)
    extra = _chunk(b"gAMA", struct.pack(">I", 45455)) + _chunk(b"tEXt", b"Comment\0hello")
    bad_crc_ancillary = struct.pack(">I", 4) + b"tIME" + b"abcd" + b"\0\0\0\0"  # libpng ignores CRC errors in ancillary chunks
    png = write_png(rows, 2, 8, [4], extra=extra + bad_crc_ancillary)
    assert np.array_equal(ing.decode_bytes(png), rows.reshape(H, W, 3))


def test_corrupt_streams_fail_loudly(ing):
    rows = _smooth(np.random.default_rng(4), 3).reshape(H, -1)
    png = write_png(rows, 2, 8, [1, 2])
    cases = {
        "truncated": png[: len(png) // 2],
        "crc": png[:60] + bytes([png[60] ^ 0x55]) + png[61:],
        "wrong size": write_png(rows[: H // 2], 2, 8, [0]),
        "no IDAT": png[:33] + _chunk(b"IEND", b""),
        "bad filter": None,
    }
    raw = bytearray(b"".join(bytes([0]) + rows[y].tobytes() for y in range(H)))
    raw[0] = 7
    cases["bad filter"] = (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 8, 2, 0, 0, 0)) +
                           _chunk(b"IDAT", zlib.compress(bytes(raw))) + _chunk(b"IEND", b""))
    for name, data in cases.items():
        with pytest.raises(_lib.BetaposeError):
            ing.decode_bytes(data)
    with pytest.raises(_lib.BetaposeError, match="not a PNG"):
        ing.decode_bytes(b"\xff\xd8\xff\xe0" + bytes(100))


def test_pool_falls_back_to_pillow_for_other_formats_and_reports_missing_files(ing, tmp_path):
    rng = np.random.default_rng(5)
    im = _smooth(rng, 3)
    p_png = str(tmp_path / "a.png")
    p_jpg = str(tmp_path / "b.jpg")
    p_int = str(tmp_path / "c.png")
    Image.fromarray(im).save(p_png)
    Image.fromarray(im).save(p_jpg, quality=95)
    # Adam7 file: written by hand is long; flag the header as interlaced and let Pillow be the judge of the content
    cv2.imwrite(p_int, im[:, :, ::-1])
    got = ing.decode_files([p_png, p_jpg, p_int])
    assert np.array_equal(got[0], im)
    assert np.array_equal(got[1], np.asarray(Image.open(p_jpg).convert("RGB")))
    assert np.array_equal(got[2], im)
    with pytest.raises(_lib.BetaposeError, match="nope.png"):
        ing.decode_files([p_png, str(tmp_path / "nope.png")])
    png = open(p_png, "rb").read()
    inter = bytearray(png)
    inter[28] = 1  # IHDR interlace byte
    inter[29:33] = struct.pack(">I", zlib.crc32(bytes(inter[12:29])) & 0xFFFFFFFF)
    rc = _lib.lib().bp_png_decode(bytes(inter), len(inter), H, W, 0, np.empty((H, W, 3), np.uint8).ctypes.data, 0)
    assert rc == _lib.ERR_UNSUPPORTED


def test_batches_generator_order_overlap_and_buffer_lifetime(tmp_path):
    """The stream BetaposeEngine.run_stream consumes: every batch in order, ragged last batch, and a yielded buffer
    untouched until two more batches have been requested (run_stream may still be uploading it)."""
    rng = np.random.default_rng(6)
    n, B = 23, 4
    ims = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    paths = []
    for i in range(n):
        p = str(tmp_path / f"{i:04d}.png")
        Image.fromarray(ims[i]).save(p, compress_level=1)
        paths.append(p)
    for depth in (0, 1, 2, 5):
        for in_flight in (1, 2, 3):  # lanes of the consumer (PipelinedEngine): the last `in_flight` batches may still be uploading
            with FrameIngest(3, H, W) as g:
                held = []
                for j, fr in enumerate(g.batches(paths, B, depth=depth, in_flight=in_flight)):
                    lo = j * B
                    assert np.array_equal(fr.numpy(), ims[lo: lo + B])
                    held.append((fr, lo))
                    for pf, plo in held[-1 - in_flight:-1]:  # the previous `in_flight` batches are still intact
                        assert np.array_equal(pf.numpy(), ims[plo: plo + B])
                assert j == (n + B - 1) // B - 1


def test_library_exports_the_ingest_symbols():
    L = _lib.lib()
    for s in ("bp_ingest_create", "bp_ingest_destroy", "bp_ingest_num_threads", "bp_png_info", "bp_png_decode", "bp_ingest_submit",
              "bp_ingest_wait"):
        assert hasattr(L, s)
    assert os.path.basename(_lib.LIB_PATH) == "libbetapose_b200.so"


# ---------------------------------------------------------------------------------------------- the ingest's own inflate
def _inflate(data: bytes, cap: int):
    import ctypes as C

    out = np.empty(cap + 1, np.uint8)
    n = C.c_size_t()
    rc = _lib.lib().bp_zlib_inflate(data, len(data), out.ctypes.data, cap, C.byref(n))
    return rc, bytes(out[: n.value]) if rc == 0 else b""


def _payloads():
    rng = np.random.default_rng(7)
    text = (b"the quick brown fox jumps over the lazy dog. " * 400)
    yield "empty", b""
    yield "one byte", b"x"
    yield "zeros", bytes(100000)                                   # distance-1 matches of length 258
    yield "period3", bytes([1, 2, 3]) * 30000                      # overlapping copies, distance 3 (Sub-filtered RGB rows)
    yield "period7", bytes(range(7)) * 9000
    yield "text", text
    yield "noise", rng.integers(0, 256, 200000, dtype=np.uint8).tobytes()          # literals only, long codes
    yield "skewed", rng.choice(np.arange(256, dtype=np.uint8), 300000, p=np.r_[[0.6], np.full(255, 0.4 / 255)]).tobytes()
    yield "far matches", (rng.integers(0, 256, 32768, dtype=np.uint8).tobytes()) * 4  # distance 32768
    yield "sparse", bytes(rng.choice([0, 0, 0, 0, 0, 0, 0, 255], 120000).astype(np.uint8))
    walk = np.cumsum(rng.integers(-2, 3, 250000)).astype(np.uint8).tobytes()
    yield "walk", walk


def test_inflate_matches_zlib_on_every_block_type():
    for name, data in _payloads():
        for level in (0, 1, 6, 9):                                  # 0 = stored blocks
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
                for mem in (1, 9):                                  # memLevel 1: tiny blocks, many table rebuilds
                    co = zlib.compressobj(level, zlib.DEFLATED, 15, mem, strategy)
                    comp = co.compress(data) + co.flush()
                    rc, got = _inflate(comp, len(data))
                    assert rc == 0 and got == data, (name, level, strategy, mem)
    # sync-flushed stream: empty stored blocks in between
    co = zlib.compressobj(6)
    comp = b"".join(co.compress(bytes([i]) * 1000) + co.flush(zlib.Z_SYNC_FLUSH) for i in range(20)) + co.flush()
    rc, got = _inflate(comp, 20000)
    assert rc == 0 and got == b"".join(bytes([i]) * 1000 for i in range(20))
    # small windows in the header
    for wbits in (9, 12):
        co = zlib.compressobj(6, zlib.DEFLATED, wbits)
        data = b"abcdefgh" * 5000
        rc, got = _inflate(co.compress(data) + co.flush(), len(data))
        assert rc == 0 and got == data


def test_inflate_rejects_bad_streams_without_touching_memory_outside_its_buffers():
    rng = np.random.default_rng(8)
    data = (b"betapose" * 3000) + rng.integers(0, 256, 20000, dtype=np.uint8).tobytes()
    comp = zlib.compress(data, 6)
    assert _inflate(comp, len(data))[0] == 0
    assert _inflate(comp, len(data) - 1)[0] == _lib.ERR_INVALID            # output too small
    assert _inflate(comp[:-1], len(data))[0] == _lib.ERR_INVALID           # trailer cut
    assert _inflate(comp[: len(comp) // 2], len(data))[0] == _lib.ERR_INVALID
    assert _inflate(comp[:-4] + b"\0\0\0\0", len(data))[0] == _lib.ERR_INVALID  # adler32
    assert _inflate(b"\x78\x9d" + comp[2:], len(data))[0] == _lib.ERR_INVALID   # header check bits
    assert _inflate(b"", 10)[0] == _lib.ERR_INVALID
    # fuzz: whatever zlib accepts we must decode identically, whatever it rejects we must reject (or, for damage the
    # checksum cannot see in time, at least never crash): 3000 single-byte corruptions + 500 random streams
    agree = 0
    for t in range(3000):
        bad = bytearray(comp)
        i = int(rng.integers(2, len(bad)))
        bad[i] ^= 1 << int(rng.integers(0, 8))
        try:
            want = zlib.decompress(bytes(bad))
        except zlib.error:
            want = None
        rc, got = _inflate(bytes(bad), len(data) + 64)
        if want is None:
            assert rc == _lib.ERR_INVALID, t
        else:
            assert rc == 0 and got == want, t
        agree += 1
    for t in range(500):
        junk = b"\x78\x9c" + rng.integers(0, 256, int(rng.integers(1, 400)), dtype=np.uint8).tobytes()
        rc, got = _inflate(junk, 4096)
        try:
            want = zlib.decompress(junk)
        except zlib.error:
            want = None
        assert (rc == 0 and got == want) if (want is not None and len(want) <= 4096) else rc == _lib.ERR_INVALID
    assert agree == 3000


def test_paeth_and_average_rows_on_noise(ing):
    """The branch-free pixel-wise Paeth / Average paths (3- and 4-byte pixels) on incompressible data, where every tie and
    sign case of the predictor occurs; first row (no row above) and every later row."""
    rng = np.random.default_rng(9)
    for ctype, ch in ((2, 3), (6, 4)):
        rows = rng.integers(0, 256, (H, W * ch), dtype=np.uint8)
        rows[5] = rows[4]          # predictor ties: a == b == c along a repeated row
        rows[7, :] = 0
        for filters in ([4], [3], [4, 3]):
            png = write_png(rows, ctype, 8, filters)
            assert np.array_equal(ing.decode_bytes(png), rows.reshape(H, W, ch)[:, :, :3])


def test_pre_decoded_containers_ppm_pgm_npy(ing, ing_bgr, tmp_path):
    """Binary PPM / PGM and .npy frames: no entropy coding, delivered by the same pool; convert_sequence writes them."""
    from betapose_b200.ingest import convert_sequence

    rng = np.random.default_rng(10)
    im = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    grey = rng.integers(0, 256, (H, W), dtype=np.uint8)
    ppm = f"P6\n# a comment line\n{W} {H}\n255\n".encode() + im.tobytes()
    pgm = f"P5 {W}\t{H} 255\n".encode() + grey.tobytes()
    assert np.array_equal(ing.decode_bytes(ppm), im) and np.array_equal(ing_bgr.decode_bytes(ppm), im[:, :, ::-1])
    assert np.array_equal(ing.decode_bytes(pgm), np.repeat(grey[..., None], 3, -1))
    assert np.array_equal(ing.decode_bytes(ppm), np.asarray(Image.open(io.BytesIO(ppm)).convert("RGB")))
    p_npy, p_npy_g, p_f = str(tmp_path / "a.npy"), str(tmp_path / "g.npy"), str(tmp_path / "f.npy")
    np.save(p_npy, im)
    np.save(p_npy_g, grey)
    np.save(p_f, np.asfortranarray(im))
    got = ing.decode_files([p_npy, p_npy_g, p_f])      # the Fortran-ordered one goes through the Python fallback
    assert np.array_equal(got[0], im) and np.array_equal(got[1], np.repeat(grey[..., None], 3, -1)) and np.array_equal(got[2], im)
    assert np.array_equal(ing_bgr.decode_files([p_npy])[0], im[:, :, ::-1])
    for bad in (ppm[:-5], f"P6\n{W} {H + 1}\n255\n".encode() + im.tobytes(), b"P6\n64 x\n255\n", f"P6\n{W} {H}\n255".encode()):
        with pytest.raises(_lib.BetaposeError):
            ing.decode_bytes(bad)
    rc = _lib.lib().bp_frame_decode(f"P6\n{W} {H}\n65535\n".encode() + bytes(2 * im.size), 15 + 2 * im.size, H, W, 0,
                                    np.empty((H, W, 3), np.uint8).ctypes.data, 0)
    assert rc == _lib.ERR_UNSUPPORTED
    np.save(str(tmp_path / "u16.npy"), im.astype(np.uint16))
    with pytest.raises(_lib.BetaposeError, match="uint8"):
        ing.decode_files([str(tmp_path / "u16.npy")])
    # convert a PNG sequence once, read it back through the pool
    pngs = []
    ims = rng.integers(0, 256, (5, H, W, 3), dtype=np.uint8)
    for i in range(5):
        p = str(tmp_path / f"s{i}.png")
        Image.fromarray(ims[i]).save(p)
        pngs.append(p)
    for fmt in ("ppm", "npy"):
        out = convert_sequence(pngs, str(tmp_path / fmt), fmt=fmt, frame_h=H, frame_w=W, n_threads=2, chunk=2)
        assert [os.path.basename(q) for q in out] == [f"s{i}.{fmt}" for i in range(5)]
        assert np.array_equal(ing.decode_files(out), ims)


def test_batches_reuses_its_ring_and_refuses_two_streams_at_once(tmp_path):
    rng = np.random.default_rng(12)
    ims = rng.integers(0, 256, (9, H, W, 3), dtype=np.uint8)
    paths = []
    for i in range(9):
        p = str(tmp_path / f"{i}.ppm")
        open(p, "wb").write(f"P6\n{W} {H}\n255\n".encode() + ims[i].tobytes())
        paths.append(p)
    with FrameIngest(2, H, W) as g:
        a = [fr.numpy().copy() for fr in g.batches(paths, 2, depth=1)]
        ring = [t.data_ptr() for t in g._ring]
        b = [fr.numpy().copy() for fr in g.batches(paths[::-1], 2, depth=1)]
        assert [t.data_ptr() for t in g._ring] == ring          # second stream: same (pinned) buffers
        assert np.array_equal(np.concatenate(a), ims) and np.array_equal(np.concatenate(b), ims[::-1])
        it = g.batches(paths, 2, depth=1)
        next(it)
        with pytest.raises(_lib.BetaposeError, match="still being consumed"):
            next(g.batches(paths, 2, depth=1))
        it.close()                                              # abandoning a stream waits for its queued decodes
        assert np.array_equal(np.concatenate([fr.numpy().copy() for fr in g.batches(paths, 4, depth=0)]), ims)
        assert g._ring[0].shape[0] == 4                         # other batch size: new ring
