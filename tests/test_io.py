"""File I/O of the CLI without GDAL (rasterio_lite): GeoTIFF round trip with the CRS tags, PixelIsPoint, odd tag types."""
import json
import struct

import numpy as np

from demcoreg_b200 import rasterio_lite as rio


def test_geotiff_roundtrip_keeps_crs_and_nodata(tmp_path):
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    gk = [1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 1, 3072, 0, 1, 32610]
    proj = rio._GK + json.dumps({"34735": gk, "34736": [], "34737": "WGS 84 / UTM zone 10N|"})
    fn = str(tmp_path / "x.tif")
    rio.write_raster(fn, a, (500000.0, 8.0, 0.0, 4100000.0, 0.0, -8.0), -9999.0, proj)
    r = rio.read_array(fn)
    assert np.array_equal(r['array'], a) and r['ndv'] == -9999.0
    assert r['gt'] == (500000.0, 8.0, 0.0, 4100000.0, 0.0, -8.0)
    assert r['proj'] == proj                                  # a product written from this raster carries the same CRS
    fn2 = str(tmp_path / "y.tif")
    rio.write_raster(fn2, a, r['gt'], r['ndv'], r['proj'])
    assert rio.read_array(fn2)['proj'] == proj
    # no CRS in, none out
    rio.write_raster(fn2, a, r['gt'], None, "")
    assert rio.read_array(fn2)['proj'] == "" and rio.read_array(fn2)['ndv'] is None


def test_pixel_is_point_shifts_the_origin_by_half_a_pixel(tmp_path):
    """GTRasterTypeGeoKey = 2: the tiepoint is a pixel centre, the geotransform origin is the corner half a pixel up-left
    (this tool measures sub-pixel offsets: the half pixel matters)."""
    a = np.zeros((3, 4), dtype=np.float32)
    gk = [1, 1, 0, 2, 1024, 0, 1, 1, 1025, 0, 1, 1]
    fn = str(tmp_path / "p.tif")
    rio.write_raster(fn, a, (100.0, 2.0, 0.0, 200.0, 0.0, -2.0), None, rio._GK + json.dumps({"34735": gk, "34736": [], "34737": ""}))
    buf = bytearray(open(fn, "rb").read())
    pat = struct.pack("<4H", 1025, 0, 1, 1)
    k = bytes(buf).index(pat)
    buf[k:k + 8] = struct.pack("<4H", 1025, 0, 1, 2)           # flip the file to PixelIsPoint
    open(fn, "wb").write(bytes(buf))
    assert rio.read_array(fn)['gt'] == (99.0, 2.0, 0.0, 201.0, 0.0, -2.0)


def test_unknown_tag_types_are_skipped(tmp_path):
    """RATIONAL XResolution / YResolution entries (type 5) used to raise KeyError in the fallback reader."""
    a = np.ones((2, 2), dtype=np.float32)
    fn = str(tmp_path / "r.tif")
    rio.write_raster(fn, a, (0.0, 1.0, 0.0, 2.0, 0.0, -1.0), -9999.0, "")
    buf = bytearray(open(fn, "rb").read())
    # BigTIFF: header 16 bytes, entry count (8), entries of 20 bytes: retag PlanarConfiguration... no: SamplesPerPixel (277)
    # entry -> tag 282 (XResolution) with type 5; the reader must ignore it
    off = struct.unpack_from("<Q", buf, 8)[0]
    n = struct.unpack_from("<Q", buf, off)[0]
    for e in range(n):
        p = off + 8 + e * 20
        if struct.unpack_from("<H", buf, p)[0] == 277:
            struct.pack_into("<HH", buf, p, 282, 5)
    open(fn, "wb").write(bytes(buf))
    assert np.array_equal(rio.read_array(fn)['array'], a)
