"""Minimal raster file I/O for the CLI: GDAL when importable, else an own uncompressed-GeoTIFF / .npz codec.

"GDAL stays only for file I/O and geotransforms" (north star) -- but GDAL is absent from this image, so the
fallback reads/writes single-band, strip-organised, uncompressed little-endian (Big)TIFF with the GeoTIFF
ModelPixelScale (33550) / ModelTiepoint (33922) tags and GDAL_NODATA (42113).  ``.npz`` files
(``array``, ``gt``, ``ndv``) are also accepted.
"""
import struct

import numpy as np

_DT = {(3, 32): np.float32, (3, 64): np.float64, (1, 8): np.uint8, (1, 16): np.uint16, (2, 16): np.int16,
       (1, 32): np.uint32, (2, 32): np.int32}


def _read_tiff(fn):
    with open(fn, "rb") as f:
        buf = f.read()
    if buf[:2] != b"II":
        raise ValueError("only little-endian TIFF is supported by the fallback reader")
    magic = struct.unpack_from("<H", buf, 2)[0]
    big = magic == 43
    if big:
        off = struct.unpack_from("<Q", buf, 8)[0]
        n = struct.unpack_from("<Q", buf, off)[0]
        esz, base = 20, off + 8
    else:
        off = struct.unpack_from("<I", buf, 4)[0]
        n = struct.unpack_from("<H", buf, off)[0]
        esz, base = 12, off + 2
    tsz = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 13: 4, 16: 8, 17: 8, 18: 8}
    tfmt = {1: "B", 2: "c", 3: "H", 4: "I", 6: "b", 7: "B", 8: "h", 9: "i", 11: "f", 12: "d", 13: "I", 16: "Q", 17: "q",
            18: "Q"}
    tags = {}
    for k in range(n):
        e = base + k * esz
        tag, typ = struct.unpack_from("<HH", buf, e)
        cnt = struct.unpack_from("<Q" if big else "<I", buf, e + 4)[0]
        voff = e + (12 if big else 8)
        nbytes = tsz.get(typ, 1) * cnt
        if nbytes > (8 if big else 4):
            voff = struct.unpack_from("<Q" if big else "<I", buf, voff)[0]
        if typ == 2:
            tags[tag] = buf[voff:voff + cnt].split(b"\0")[0].decode(errors="replace")
        elif typ in tfmt:
            tags[tag] = struct.unpack_from("<%d%s" % (cnt, tfmt[typ]), buf, voff)
        # (RATIONAL / unknown field types -- XResolution and friends -- carry nothing this codec needs: skipped)
    w, h = tags[256][0], tags[257][0]
    if tags.get(259, (1,))[0] != 1:
        raise ValueError("compressed TIFF needs GDAL (not installed); write uncompressed strips")
    if 322 in tags:
        raise ValueError("tiled TIFF needs GDAL (not installed); write strips")
    dt = _DT[(tags.get(339, (1,))[0], tags[258][0])]
    offs, cnts = tags[273], tags[279]
    a = np.concatenate([np.frombuffer(buf, dtype=dt, count=c // np.dtype(dt).itemsize, offset=o) for o, c in zip(offs, cnts)])
    a = a.reshape(h, w).copy()
    gt = (0.0, 1.0, 0.0, float(h), 0.0, -1.0)
    if 33550 in tags and 33922 in tags:
        sx, sy = tags[33550][0], tags[33550][1]
        tp = tags[33922]
        gt = (tp[3] - tp[0] * sx, sx, 0.0, tp[4] + tp[1] * sy, 0.0, -sy)
        gk = tags.get(34735, ())
        for q in range(4, len(gk) - 3, 4):
            if gk[q] == 1025 and gk[q + 1] == 0 and gk[q + 3] == 2:   # GTRasterTypeGeoKey = RasterPixelIsPoint:
                gt = (gt[0] - 0.5 * sx, sx, 0.0, gt[3] + 0.5 * sy, 0.0, -sy)   # the tiepoint is a pixel CENTRE
    ndv = None
    if 42113 in tags:
        try:
            ndv = float(tags[42113])
        except ValueError:
            ndv = None
    return dict(array=a, gt=gt, ndv=ndv, proj=_pack_geokeys(tags))


_GK = "GEOKEYS:"


def _pack_geokeys(tags):
    """The CRS of a GeoTIFF as this codec carries it: the three GeoKey tags verbatim (34735 directory, 34736 doubles, 34737
    ASCII), packed into the ``proj`` string so that every product written from this raster gets them back."""
    import json
    if 34735 not in tags:
        return ""
    return _GK + json.dumps({"34735": list(tags[34735]), "34736": list(tags.get(34736, ())), "34737": tags.get(34737, "")})


def _unpack_geokeys(proj):
    import json
    if not proj or not proj.startswith(_GK):
        return None
    d = json.loads(proj[len(_GK):])
    gk = list(d["34735"])
    for q in range(4, len(gk) - 3, 4):      # products are written PixelIsArea (the geotransform is corner-based)
        if gk[q] == 1025 and gk[q + 1] == 0:
            gk[q + 3] = 1
    return gk, list(d["34736"]), d["34737"]


def _write_tiff(fn, a, gt, ndv, proj=""):
    a = np.ascontiguousarray(a)
    h, w = a.shape
    fmt = 3 if a.dtype.kind == "f" else (2 if a.dtype.kind == "i" else 1)
    bits = a.dtype.itemsize * 8
    data = a.tobytes()
    nd = (("%.17g" % ndv) + "\0").encode() if ndv is not None else None
    scale = struct.pack("<3d", gt[1], -gt[5], 0.0)
    tie = struct.pack("<6d", 0.0, 0.0, 0.0, gt[0], gt[3], 0.0)
    # BigTIFF always: one strip, offsets are 64-bit
    entries = []
    extra = b""
    hdr = 16
    geo = _unpack_geokeys(proj)
    n_entries = 12 + (1 if nd else 0) + ((1 + (1 if geo[1] else 0) + (1 if geo[2] else 0)) if geo else 0)
    ifd_size = 8 + n_entries * 20 + 8
    extra_off = hdr + ifd_size
    data_off = None

    def add(tag, typ, cnt, payload):
        nonlocal extra
        if len(payload) <= 8:
            entries.append(struct.pack("<HHQ", tag, typ, cnt) + payload.ljust(8, b"\0"))
        else:
            entries.append(struct.pack("<HHQQ", tag, typ, cnt, extra_off + len(extra)))
            extra += payload + (b"\0" * (len(payload) % 2))
    add(256, 4, 1, struct.pack("<I", w))
    add(257, 4, 1, struct.pack("<I", h))
    add(258, 3, 1, struct.pack("<H", bits))
    add(259, 3, 1, struct.pack("<H", 1))
    add(262, 3, 1, struct.pack("<H", 1))
    strip_idx = len(entries)
    add(273, 16, 1, struct.pack("<Q", 0))
    add(277, 3, 1, struct.pack("<H", 1))
    add(278, 4, 1, struct.pack("<I", h))
    add(279, 16, 1, struct.pack("<Q", len(data)))
    add(339, 3, 1, struct.pack("<H", fmt))
    add(33550, 12, 3, scale)
    add(33922, 12, 6, tie)
    if geo:
        add(34735, 3, len(geo[0]), struct.pack("<%dH" % len(geo[0]), *geo[0]))
        if geo[1]:
            add(34736, 12, len(geo[1]), struct.pack("<%dd" % len(geo[1]), *geo[1]))
        if geo[2]:
            asc = (geo[2] + "\0").encode()
            add(34737, 2, len(asc), asc)
    if nd:
        add(42113, 2, len(nd), nd)
    n_entries = len(entries)
    ifd_size = 8 + n_entries * 20 + 8
    assert extra_off == hdr + ifd_size, "entry count changed"
    data_off = extra_off + len(extra)
    data_off += (-data_off) % 16
    entries[strip_idx] = struct.pack("<HHQQ", 273, 16, 1, data_off)
    entries.sort(key=lambda e: struct.unpack_from("<H", e)[0])
    with open(fn, "wb") as f:
        f.write(b"II" + struct.pack("<HHHQ", 43, 8, 0, hdr))
        f.write(struct.pack("<Q", n_entries) + b"".join(entries) + struct.pack("<Q", 0))
        f.write(extra)
        f.write(b"\0" * (data_off - extra_off - len(extra)))
        f.write(data)


def read_array(fn):
    """-> dict(array, gt, ndv, proj)"""
    if fn.endswith(".npz"):
        z = np.load(fn)
        return dict(array=z["array"], gt=tuple(float(g) for g in z["gt"]),
                    ndv=float(z["ndv"]) if "ndv" in z and not np.isnan(z["ndv"]) else None, proj="")
    try:
        from osgeo import gdal  # noqa: F401
        ds = gdal.Open(fn)
        b = ds.GetRasterBand(1)
        return dict(array=b.ReadAsArray(), gt=ds.GetGeoTransform(), ndv=b.GetNoDataValue(), proj=ds.GetProjection())
    except ImportError:
        return _read_tiff(fn)


def read_raster(fn):
    from .raster import Raster
    r = read_array(fn)
    ndv = r["ndv"]
    if ndv is None:   # iolib.get_ndv_b: UL pixel if it equals LR (or is NaN), else 0
        a = r["array"]
        ndv = float(a[0, 0]) if (a[0, 0] == a[-1, -1] or np.isnan(a[0, 0])) else 0.0
    return Raster(np.ascontiguousarray(r["array"], dtype=np.float32), r["gt"], ndv, r["proj"] or 'LOCAL_CS["metres"]')


def write_raster(fn, a, gt, ndv=None, proj=""):
    if fn.endswith(".npz"):
        np.savez(fn, array=a, gt=np.array(gt, dtype=np.float64), ndv=np.float64(np.nan if ndv is None else ndv))
        return
    try:
        from osgeo import gdal
        drv = gdal.GetDriverByName("GTiff")
        ds = drv.Create(fn, a.shape[1], a.shape[0], 1, gdal.GDT_Float32 if a.dtype == np.float32 else gdal.GDT_Byte,
                        options=["COMPRESS=LZW", "TILED=YES", "BIGTIFF=IF_SAFER"])
        ds.SetGeoTransform(gt)
        if proj:
            ds.SetProjection(proj)
        b = ds.GetRasterBand(1)
        if ndv is not None:
            b.SetNoDataValue(ndv)
        b.WriteArray(a)
        ds = None
    except ImportError:
        _write_tiff(fn, a, gt, ndv, proj)
