"""A small writer of VTK XML UnstructuredGrid files in the encodings Galaxy's datasets come in (test infrastructure for the
VTK-free reader of galaxy_b200/host/gxy_vtu.cpp; layout per the VTK file-format specification: [header][data] blocks,
header = byte count, or [#blocks][block size][last block size][compressed sizes...] with vtkZLibDataCompressor)."""
import base64
import struct
import zlib

import numpy as np

VTK_TYPE = {np.dtype("float32"): "Float32", np.dtype("float64"): "Float64", np.dtype("int32"): "Int32", np.dtype("int64"): "Int64",
            np.dtype("uint8"): "UInt8"}


def _block(raw, compressed, header64, block_size=32768):
    """(header bytes, data bytes) of one DataArray."""
    fmt = "<Q" if header64 else "<I"
    if not compressed:
        return struct.pack(fmt, len(raw)), raw
    blocks = [raw[i:i + block_size] for i in range(0, len(raw), block_size)] or [b""]
    comp = [zlib.compress(b) for b in blocks]
    last = len(blocks[-1]) if len(blocks[-1]) != block_size else 0
    hdr = struct.pack(fmt, len(blocks)) + struct.pack(fmt, block_size) + struct.pack(fmt, last) + b"".join(struct.pack(fmt, len(c)) for c in comp)
    return hdr, b"".join(comp)


def write_vtu(path, points, triangles=None, normals=None, scalars=None, scalars_name="data", mode="ascii", compressed=False, header64=False,
              normals_name="Normals", declare_scalars=True, polylines=None):
    """mode: ascii | binary (inline base64) | appended-raw | appended-base64.  triangles None -> one VTK_VERTEX cell per point;
    polylines: list of point-id lists -> VTK_POLY_LINE cells (what vtkStreamTracer output holds, tests/create_data_driven_datasets.vpy)."""
    points = np.ascontiguousarray(points, np.float32)
    n = len(points)
    if polylines is not None:
        conn = np.asarray([i for l in polylines for i in l], dtype=np.int64)
        offs = np.cumsum([len(l) for l in polylines], dtype=np.int64)
        types = np.full(len(polylines), 4, np.uint8)
    elif triangles is None:
        conn = np.arange(n, dtype=np.int64)
        offs = np.arange(1, n + 1, dtype=np.int64)
        types = np.full(n, 1, np.uint8)
    else:
        tri = np.ascontiguousarray(triangles, np.int64)
        conn = tri.ravel()
        offs = np.arange(3, 3 * len(tri) + 1, 3, dtype=np.int64)
        types = np.full(len(tri), 5, np.uint8)
    appended = bytearray()
    appended_b64 = []

    def data_array(arr, name=None, ncomp=None):
        arr = np.ascontiguousarray(arr)
        attrs = 'type="%s"' % VTK_TYPE[arr.dtype]
        if name:
            attrs += ' Name="%s"' % name
        if ncomp:
            attrs += ' NumberOfComponents="%d"' % ncomp
        if mode == "ascii":
            txt = " ".join(repr(float(v)) if arr.dtype.kind == "f" else str(int(v)) for v in arr.ravel())
            return '<DataArray %s format="ascii">\n%s\n</DataArray>\n' % (attrs, txt)
        hdr, data = _block(arr.tobytes(), compressed, header64)
        if mode == "binary":
            enc = base64.b64encode(hdr).decode() + base64.b64encode(data).decode() if compressed else base64.b64encode(hdr + data).decode()
            return '<DataArray %s format="binary">\n%s\n</DataArray>\n' % (attrs, enc)
        if mode == "appended-raw":
            off = len(appended)
            appended.extend(hdr + data)
            return '<DataArray %s format="appended" offset="%d"/>\n' % (attrs, off)
        if mode == "appended-base64":
            off = sum(len(x) for x in appended_b64)
            appended_b64.append(base64.b64encode(hdr).decode() + base64.b64encode(data).decode())  # header as its own unit
            return '<DataArray %s format="appended" offset="%d"/>\n' % (attrs, off)
        raise ValueError(mode)

    head = '<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" header_type="%s"%s>\n' % (
        "UInt64" if header64 else "UInt32", ' compressor="vtkZLibDataCompressor"' if compressed else "")
    s = head + '<UnstructuredGrid>\n<Piece NumberOfPoints="%d" NumberOfCells="%d">\n' % (n, len(offs))
    pd = ""
    if scalars is not None:
        pd += data_array(np.asarray(scalars), scalars_name)
    if normals is not None:
        pd += data_array(np.asarray(normals, np.float32), normals_name, 3)
    s += '<PointData%s>\n%s</PointData>\n' % (' Scalars="%s"' % scalars_name if scalars is not None and declare_scalars else "", pd)
    s += "<!-- cell data is ignored by Galaxy --><CellData></CellData>\n"
    s += "<Points>\n" + data_array(points, "Points", 3) + "</Points>\n"
    s += "<Cells>\n" + data_array(conn, "connectivity") + data_array(offs, "offsets") + data_array(types, "types") + "</Cells>\n"
    s += "</Piece>\n</UnstructuredGrid>\n"
    out = s.encode()
    if mode == "appended-raw":
        out += b'<AppendedData encoding="raw">\n_' + bytes(appended) + b"\n</AppendedData>\n"
    elif mode == "appended-base64":
        out += b'<AppendedData encoding="base64">\n_' + "".join(appended_b64).encode() + b"\n</AppendedData>\n"
    out += b"</VTKFile>\n"
    with open(path, "wb") as f:
        f.write(out)
