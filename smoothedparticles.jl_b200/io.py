"""ParaView output and input of particle systems: the counterpart of the reference's ``src/IO.jl``.

    new_pvd_file(path)            src/IO.jl:20-26    a frame collection in a directory
    save_frame(data, sys, *vars)  src/IO.jl:50-75    one ``frame<k>.vtp`` with the positions and the named fields
    save_pvd_file(data)           src/IO.jl:33-35    ``result.pvd`` indexing the frames
    import_particles(sys, path)   src/IO.jl:83-130   append the particles of a ``.vtp`` file to a system

The reference delegates the file format to WriteVTK.jl / ReadVTK.jl (third-party, versions unpinned).  What they
write — seen in the reference's own ``examples/init/cylinder.vtp`` — is VTK XML PolyData, version 1.0, little
endian, ``header_type="UInt64"``, one ``Verts`` cell per point (``connectivity`` 0..N-1, ``offsets`` 1..N, Int64),
Float64 point data with ``NumberOfComponents`` 1 / 3 / 9, everything in one ``<AppendedData encoding="raw">``
section, each array zlib-compressed in blocks behind a header ``[nblocks, blocksize, lastblocksize, csize...]``.
``write_vtp`` produces exactly that layout (or the uncompressed variant: header = byte count), ``read_vtp`` reads
both, with UInt32 or UInt64 headers.  The particle data come from the device through ``sp_download`` (one
SoA -> AoS transpose kernel per field); there is no per-particle host loop.

Fields are whatever the system was created with: scalars (n,), vectors (n, 3) and matrices (n, 9).  The nine numbers
of a matrix are kept in the order the reference writes them (the storage order of its ``RealMatrix``, an
``SMatrix{3,3}``: src/IO.jl:60-66 flattens with ``CartesianIndices``, :121-125 rebuilds from ``1:9``).
"""
from __future__ import annotations

import os
import re
import struct
import zlib
from typing import Dict, Sequence, Tuple

import numpy as np

_BLOCK = 1 << 15          # WriteVTK / VTK default compression block size
_NP_OF = {"Float64": "<f8", "Float32": "<f4", "Int64": "<i8", "Int32": "<i4", "UInt64": "<u8", "UInt32": "<u4",
          "UInt8": "u1", "Int8": "i1"}


# ----------------------------------------------------------------------------------------------- writer
def _encode(raw: bytes, compress: bool) -> bytes:
    if not compress:
        return struct.pack("<Q", len(raw)) + raw
    if len(raw) == 0:
        return struct.pack("<QQQ", 0, _BLOCK, 0)
    blocks = [raw[i:i + _BLOCK] for i in range(0, len(raw), _BLOCK)]
    comp = [zlib.compress(b) for b in blocks]
    last = len(blocks[-1]) if len(blocks[-1]) != _BLOCK else 0
    head = struct.pack("<QQQ", len(blocks), _BLOCK, last) + b"".join(struct.pack("<Q", len(c)) for c in comp)
    return head + b"".join(comp)


def write_vtp(path: str, points: np.ndarray, fields: Dict[str, np.ndarray], compress: bool = True) -> None:
    """One PolyData piece with a vertex per particle (src/IO.jl:37-47, capture_frame) and Float64 point data."""
    pts = np.ascontiguousarray(points, dtype="<f8").reshape(-1, 3)
    n = pts.shape[0]
    arrays = [("Points", "Float64", 3, pts.tobytes()),
              ("connectivity", "Int64", 1, np.arange(n, dtype="<i8").tobytes()),
              ("offsets", "Int64", 1, np.arange(1, n + 1, dtype="<i8").tobytes())]
    for name, a in fields.items():
        a = np.ascontiguousarray(a, dtype="<f8")
        if a.shape[0] != n:
            raise ValueError(f"field {name}: {a.shape[0]} rows for {n} points")
        nc = 1 if a.ndim == 1 else int(np.prod(a.shape[1:]))
        arrays.append((name, "Float64", nc, a.tobytes()))
    blobs, offsets, off = [], [], 0
    for _, _, _, raw in arrays:
        b = _encode(raw, compress)
        blobs.append(b)
        offsets.append(off)
        off += len(b)

    def tag(k, indent):
        name, typ, nc, _ = arrays[k]
        return (f'{indent}<DataArray type="{typ}" Name="{name}" NumberOfComponents="{nc}" format="appended" '
                f'offset="{offsets[k]}"/>\n')

    comp_attr = ' compressor="vtkZLibDataCompressor"' if compress else ""
    head = ('<?xml version="1.0" encoding="utf-8"?>\n'
            f'<VTKFile type="PolyData" version="1.0" byte_order="LittleEndian" header_type="UInt64"{comp_attr}>\n'
            '  <PolyData>\n'
            f'    <Piece NumberOfPoints="{n}" NumberOfVerts="{n}">\n'
            '      <Points>\n' + tag(0, "        ") + '      </Points>\n'
            '      <Verts>\n' + tag(1, "        ") + tag(2, "        ") + '      </Verts>\n'
            '      <PointData>\n' + "".join(tag(k, "        ") for k in range(3, len(arrays))) + '      </PointData>\n'
            '    </Piece>\n'
            '  </PolyData>\n'
            '  <AppendedData encoding="raw">\n_')
    with open(path, "wb") as f:
        f.write(head.encode())
        for b in blobs:
            f.write(b)
        f.write(b"\n  </AppendedData>\n</VTKFile>\n")


# ----------------------------------------------------------------------------------------------- reader
def _decode(buf: bytes, off: int, hdr: str, compressed: bool) -> bytes:
    hs = struct.calcsize(hdr)
    if not compressed:
        (nbytes,) = struct.unpack_from(hdr, buf, off)
        return buf[off + hs: off + hs + nbytes]
    nblocks, _bs, _last = struct.unpack_from("<" + hdr[1] * 3, buf, off)
    sizes = struct.unpack_from("<" + hdr[1] * nblocks, buf, off + 3 * hs) if nblocks else ()
    p = off + (3 + nblocks) * hs
    out = []
    for cs in sizes:
        out.append(zlib.decompress(buf[p:p + cs]))
        p += cs
    return b"".join(out)


def read_vtp(path: str) -> Tuple[np.ndarray, Dict[str, np.ndarray]]:
    """Points (n, 3) and the point-data arrays of a PolyData file with appended raw data (what WriteVTK.jl writes
    and what ``write_vtp`` writes).  Arrays with 3 or 9 components come back as (n, 3) / (n, 9)."""
    blob = open(path, "rb").read()
    m = re.search(rb'<AppendedData[^>]*encoding="raw"[^>]*>\s*_', blob)
    if not m:
        raise ValueError(f"{path}: only appended raw data is supported")
    xml = blob[:m.start()].decode("utf-8", errors="replace")
    data0 = m.end()
    vf = re.search(r"<VTKFile([^>]*)>", xml).group(1)
    if 'type="PolyData"' not in vf:
        raise ValueError(f"{path}: not a PolyData file")
    if 'byte_order="BigEndian"' in vf:
        raise ValueError(f"{path}: big-endian files are not supported")
    hdr = "<Q" if 'header_type="UInt64"' in vf else "<I"
    compressed = "vtkZLibDataCompressor" in vf
    n = int(re.search(r'NumberOfPoints="(\d+)"', xml).group(1))

    def arrays_in(section):
        sm = re.search(rf"<{section}[^>]*>(.*?)</{section}>", xml, flags=re.S)
        out = []
        if not sm:
            return out
        for am in re.finditer(r"<DataArray([^>]*)/?>", sm.group(1)):
            attrs = dict(re.findall(r'(\w+)="([^"]*)"', am.group(1)))
            out.append(attrs)
        return out

    def load(attrs):
        if attrs.get("format") != "appended":
            raise ValueError(f"{path}: array {attrs.get('Name')} is not in the appended section")
        raw = _decode(blob, data0 + int(attrs["offset"]), hdr, compressed)
        a = np.frombuffer(raw, dtype=_NP_OF[attrs["type"]])
        nc = int(attrs.get("NumberOfComponents", "1"))
        return a.reshape(-1, nc) if nc > 1 else a

    pts = load(arrays_in("Points")[0]).astype(np.float64).reshape(-1, 3)
    if pts.shape[0] != n:
        raise ValueError(f"{path}: {pts.shape[0]} points, header says {n}")
    fields = {}
    for attrs in arrays_in("PointData"):
        a = load(attrs).astype(np.float64)
        if a.shape[0] != n:
            raise ValueError(f"{path}: point data {attrs.get('Name')} has {a.shape[0]} rows for {n} points")
        fields[attrs["Name"]] = a
    return pts, fields


# ----------------------------------------------------------------------------------------------- the IO.jl surface
class DataStorage:
    """``DataStorage`` of src/IO.jl:9-13: a directory of frames plus the ``.pvd`` collection that lists them."""

    def __init__(self, path: str, compress: bool = True):
        self.path = path
        self.frame = 0
        self.compress = compress
        self.entries = []   # (timestep, file name)


def new_pvd_file(path: str, compress: bool = True) -> DataStorage:
    os.makedirs(path, exist_ok=True)
    return DataStorage(path, compress)


def save_frame(data: DataStorage, sys, *vars: str) -> str:
    """``save_frame!(data, sys, vars...)``: positions and the named fields of every particle, in the reference's
    particle order, into ``<path>/frame<k>.vtp``; the frame is registered under time step k (src/IO.jl:72-73)."""
    name = f"frame{data.frame}.vtp"
    fields = {v: sys.get(v) for v in vars}
    write_vtp(os.path.join(data.path, name), sys.get("x"), fields, compress=data.compress)
    data.entries.append((float(data.frame), name))
    data.frame += 1
    return os.path.join(data.path, name)


def save_pvd_file(data: DataStorage) -> str:
    """``save_pvd_file(data)``: writes ``<path>/result.pvd`` (what ``vtk_save`` of the collection does)."""
    out = os.path.join(data.path, "result.pvd")
    with open(out, "w") as f:
        f.write('<?xml version="1.0" encoding="utf-8"?>\n'
                '<VTKFile type="Collection" version="1.0" byte_order="LittleEndian">\n  <Collection>\n')
        for t, name in data.entries:
            f.write(f'    <DataSet timestep="{t!r}" part="0" file="{name}"/>\n')
        f.write("  </Collection>\n</VTKFile>\n")
    return out


def read_pvd(path: str) -> Sequence[Tuple[float, str]]:
    text = open(path).read()
    return [(float(t), f) for t, f in re.findall(r'<DataSet[^>]*timestep="([^"]*)"[^>]*file="([^"]*)"', text)]


def import_particles(sys, path: str, constructor=None, **constants) -> int:
    """``import_particles!(sys, path, particle_constructor)`` (src/IO.jl): the points of the file become new particles at
    the end of the reference order.  The reference builds every particle with the caller's constructor — which sets
    non-zero defaults such as ``rho0`` or ``m`` — and then overwrites the fields present in the file.  Here the defaults
    come either as ``**constants`` (field name -> scalar or per-component tuple, as in ``generate_particles``) or from
    ``constructor(x)``, a callable that gets the (n, 3) positions and returns a ``{field: array or scalar}`` dict; every
    point-data array of the file whose name and width match a field of the system is then laid over them.  Fields named
    by neither stay zero."""
    pts, fields = read_vtp(path)
    n = pts.shape[0]
    arrays = {}
    defaults = dict(constants)
    if constructor is not None:
        defaults.update(constructor(pts))
    for name, val in defaults.items():
        if name == "x":
            continue
        if name not in sys.fields:
            raise KeyError(f"import_particles: the system has no field '{name}'")
        nc = sys.fields[name]
        a = np.asarray(val, dtype=np.float64)
        arrays[name] = np.broadcast_to(a, (n,) if nc == 1 else (n, nc)).copy()
    for name, a in fields.items():
        if name in sys.fields and name != "x":
            nc = sys.fields[name]
            if (a.ndim == 1 and nc == 1) or (a.ndim == 2 and a.shape[1] == nc):
                arrays[name] = a
    arrays["x"] = pts
    sys.add_particles(**arrays)
    return n
