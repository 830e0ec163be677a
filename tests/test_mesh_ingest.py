"""Host logic: the VTK XML unstructured-grid reader (vgposp_b200/mesh_ingest.py) against files written here in every
encoding the format allows -- the accessors of vtktools.vtu that gp_pvtk.py:24-75 uses must return the same arrays
whatever the container looks like.  (The `vtk` package is not available, so the files are produced by a small writer
in this test that follows the VTK XML layout: [header][data] per array, base64 or raw, optional zlib blocks.)"""
import base64
import os
import struct
import zlib

import numpy as np
import pytest

from vgposp_b200 import mesh_ingest

TYPES = {np.dtype("f4"): "Float32", np.dtype("f8"): "Float64", np.dtype("i4"): "Int32", np.dtype("i8"): "Int64",
         np.dtype("u1"): "UInt8"}


def payload(a, order, header, compress, block=1 << 10):
    """(header bytes, body bytes) of one array as VTK writes it."""
    raw = np.ascontiguousarray(a).astype(a.dtype.newbyteorder(order)).tobytes()
    fmt = order + ("I" if header == "UInt32" else "Q")
    if not compress:
        return struct.pack(fmt, len(raw)), raw
    blocks = [raw[i:i + block] for i in range(0, len(raw), block)] or [b""]
    comp = [zlib.compress(b) for b in blocks]
    last = len(blocks[-1]) if len(blocks[-1]) != block else 0
    head = struct.pack(fmt, len(blocks)) + struct.pack(fmt, block) + struct.pack(fmt, last)
    head += b"".join(struct.pack(fmt, len(c)) for c in comp)
    return head, b"".join(comp)


def write_vtu(path, points, point_data, cells, cell_data=None, mode="ascii", order="<", header="UInt32",
              compress=False, split_header=False):
    """mode: ascii | binary | appended_raw | appended_b64."""
    conn, offsets, types = cells
    appended, chunks = mode.startswith("appended"), []

    def data_array(a, name=None, nc=1):
        a = np.asarray(a)
        attrs = 'type="%s"' % TYPES[a.dtype]
        if name:
            attrs += ' Name="%s"' % name
        if nc > 1:
            attrs += ' NumberOfComponents="%d"' % nc
        if mode == "ascii":
            return '<DataArray %s format="ascii">%s</DataArray>' % (attrs, " ".join(repr(v.item()) for v in a.ravel()))
        head, body = payload(a, order, header, compress)
        if mode == "binary":
            if compress or split_header:
                text = base64.b64encode(head).decode() + base64.b64encode(body).decode()
            else:
                text = base64.b64encode(head + body).decode()
            return '<DataArray %s format="binary">%s</DataArray>' % (attrs, text)
        if mode == "appended_raw":
            off = sum(len(c) for c in chunks)
            chunks.append(head + body)
        else:
            off = sum(len(c) for c in chunks)
            enc = (base64.b64encode(head) + base64.b64encode(body)) if compress else base64.b64encode(head + body)
            chunks.append(enc)
        return '<DataArray %s format="appended" offset="%d"/>' % (attrs, off)

    xml = ['<?xml version="1.0"?>',
           '<VTKFile type="UnstructuredGrid" version="1.0" byte_order="%s" header_type="%s"%s>' %
           ("LittleEndian" if order == "<" else "BigEndian", header,
            ' compressor="vtkZLibDataCompressor"' if compress else ""),
           "<UnstructuredGrid>",
           '<Piece NumberOfPoints="%d" NumberOfCells="%d">' % (len(points), len(offsets)),
           "<PointData>"]
    for name, a in point_data.items():
        xml.append(data_array(a, name, 1 if a.ndim == 1 else a.shape[1]))
    xml.append("</PointData><CellData>")
    for name, a in (cell_data or {}).items():
        xml.append(data_array(a, name, 1 if a.ndim == 1 else a.shape[1]))
    xml.append("</CellData><Points>")
    xml.append(data_array(points, "Points", 3))
    xml.append("</Points><Cells>")
    xml.append(data_array(conn, "connectivity"))
    xml.append(data_array(offsets, "offsets"))
    xml.append(data_array(types, "types"))
    xml.append("</Cells></Piece></UnstructuredGrid>")
    out = "\n".join(xml).encode()
    if appended:
        out += b'\n<AppendedData encoding="%s">\n_' % (b"raw" if mode == "appended_raw" else b"base64")
        out += b"".join(chunks) + b"\n</AppendedData>"
    out += b"\n</VTKFile>\n"
    with open(path, "wb") as f:
        f.write(out)


def room(seed, n=257):
    """A room-simulation-like time step: tetrahedra over n nodes with the reference's feature names."""
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    pd = {"Tracer": rng.uniform(0, 1, n), "Pressure": rng.normal(size=n).astype(np.float32),
          "Temperature": 290 + rng.uniform(0, 10, n), "Time": np.full(n, 0.5 * seed), "Density": np.ones(n),
          "Velocity": rng.normal(size=(n, 3)), "Stress": rng.normal(size=(n, 9)), "Id": np.arange(n, dtype=np.int32)}
    ncell = 100
    conn = rng.integers(0, n, 4 * ncell).astype(np.int64)
    cells = (conn, (4 * np.arange(1, ncell + 1)).astype(np.int64), np.full(ncell, 10, dtype=np.uint8))
    cd = {"CellVolume": rng.uniform(0, 1, ncell)}
    return pts, pd, cells, cd


VARIANTS = [dict(mode="ascii"), dict(mode="binary"), dict(mode="binary", split_header=True),
            dict(mode="binary", header="UInt64"), dict(mode="binary", compress=True),
            dict(mode="binary", compress=True, header="UInt64", order=">"), dict(mode="binary", order=">"),
            dict(mode="appended_raw"), dict(mode="appended_raw", compress=True, header="UInt64"),
            dict(mode="appended_b64"), dict(mode="appended_b64", compress=True)]


@pytest.mark.parametrize("variant", VARIANTS, ids=lambda v: "-".join("%s=%s" % kv for kv in v.items()))
def test_every_encoding_reads_back_the_same_arrays(tmp_path, variant):
    pts, pd, cells, cd = room(3)
    fn = str(tmp_path / "step.vtu")
    write_vtu(fn, pts, pd, cells, cd, **variant)
    g = mesh_ingest.vtu(fn)
    loc = g.GetLocations()
    assert loc.dtype == np.float64 and loc.shape == (len(pts), 3)
    np.testing.assert_array_equal(loc, pts.astype(np.float64))
    assert g.GetFieldNames() == list(pd)
    for name in ("Tracer", "Temperature", "Time", "Density"):
        np.testing.assert_array_equal(g.GetScalarField(name), pd[name])
    np.testing.assert_array_equal(g.GetScalarField("Pressure"), pd["Pressure"].astype(np.float64))
    np.testing.assert_array_equal(g.GetVectorField("Velocity"), pd["Velocity"])
    np.testing.assert_array_equal(g.GetVectorNorm("Velocity"), np.sqrt((pd["Velocity"] ** 2).sum(1)))
    assert g.GetField("Stress").shape == (len(pts), 3, 3)
    np.testing.assert_array_equal(g.GetField("Stress").reshape(len(pts), 9), pd["Stress"])
    assert g.GetField("Id").dtype == np.int32 and g.GetField("Id").shape == (len(pts), 1)
    assert [g.GetFieldRank(n) for n in ("Tracer", "Velocity", "Stress")] == [0, 1, 2]
    np.testing.assert_array_equal(g.GetScalarField("CellVolume"), cd["CellVolume"])       # cell data as the fallback
    assert g.GetScalarRange("Tracer") == (pd["Tracer"].min(), pd["Tracer"].max())
    for c in (0, 1, 99):
        np.testing.assert_array_equal(g.GetCellPoints(c), cells[0][4 * c:4 * c + 4])


def test_errors_follow_the_reference(tmp_path):
    pts, pd, cells, cd = room(1, n=8)
    fn = str(tmp_path / "a.vtu")
    write_vtu(fn, pts, pd, cells, cd)
    g = mesh_ingest.vtu(fn)
    with pytest.raises(Exception, match="couldn't find point or cell scalar field data with name Nope"):
        g.GetScalarField("Nope")
    with pytest.raises(Exception, match="don't recognise file extension"):
        mesh_ingest.vtu(str(tmp_path / "a.vtk"))
    empty = str(tmp_path / "empty.vtu")
    write_vtu(empty, np.zeros((0, 3), np.float32), {}, (np.zeros(0, np.int64), np.zeros(0, np.int64),
                                                         np.zeros(0, np.uint8)))
    with pytest.raises(Exception, match="No points or cells found"):
        mesh_ingest.vtu(empty)
    assert mesh_ingest.vtu().GetLocations().shape == (0, 3)                               # vtu() = empty grid


def test_pvtu_concatenates_its_pieces(tmp_path):
    a, b = room(5, n=40), room(6, n=25)
    write_vtu(str(tmp_path / "p_0.vtu"), *a, mode="binary", compress=True)
    write_vtu(str(tmp_path / "p_1.vtu"), *b, mode="appended_raw")
    with open(str(tmp_path / "p.pvtu"), "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="PUnstructuredGrid" version="0.1" byte_order="LittleEndian">\n'
                '<PUnstructuredGrid GhostLevel="0"><PPointData><PDataArray type="Float64" Name="Tracer"/></PPointData>\n'
                '<Piece Source="p_0.vtu"/><Piece Source="p_1.vtu"/></PUnstructuredGrid></VTKFile>\n')
    g = mesh_ingest.vtu(str(tmp_path / "p.pvtu"))
    np.testing.assert_array_equal(g.GetLocations(), np.vstack([a[0], b[0]]).astype(np.float64))
    np.testing.assert_array_equal(g.GetScalarField("Tracer"), np.concatenate([a[1]["Tracer"], b[1]["Tracer"]]))
    np.testing.assert_array_equal(g.GetCellPoints(100), b[2][0][:4] + 40)                  # second piece, renumbered


def test_time_steps_stack_like_gp_pvtk(tmp_path):
    """gp_pvtk.py:24-75: walk the file indices, skip missing / empty files, vstack coordinates and features."""
    files, want_xyz, want_tr = [], [], []
    for i in range(620, 626):
        fn = str(tmp_path / ("room_selection_0001_%d.vtu" % i))
        files.append(fn)
        if i == 622:
            continue                                         # missing
        if i == 623:
            open(fn, "w").close()                            # empty
            continue
        pts, pd, cells, cd = room(i, n=30 + i % 7)
        write_vtu(fn, pts, pd, cells, cd, mode=("binary", "appended_raw", "ascii")[i % 3], compress=(i % 2 == 0) and
                  i % 3 != 2)
        want_xyz.append(pts.astype(np.float64))
        want_tr.append(pd["Tracer"])
    xyz, feats, used = mesh_ingest.load_time_steps(files)
    assert [os.path.basename(u) for u in used] == ["room_selection_0001_%d.vtu" % i for i in (620, 621, 624, 625)]
    np.testing.assert_array_equal(xyz, np.vstack(want_xyz))
    np.testing.assert_array_equal(feats["Tracer"], np.concatenate(want_tr).reshape(-1, 1))
    assert feats["Velocity"].shape == (len(xyz), 3) and feats["Pressure"].shape == (len(xyz), 1)
    xyz2, _, used2 = mesh_ingest.load_time_steps(files, max_steps=2)
    assert len(used2) == 2 and len(xyz2) == len(want_xyz[0]) + len(want_xyz[1])


def test_several_pieces_in_one_file(tmp_path):
    """A .vtu may hold more than one <Piece>; points, point data and cells are concatenated (cells renumbered)."""
    a, b = room(11, n=12), room(12, n=9)
    fa, fb = str(tmp_path / "a.vtu"), str(tmp_path / "b.vtu")
    write_vtu(fa, *a, mode="ascii")
    write_vtu(fb, *b, mode="ascii")

    def piece(fn):
        text = open(fn).read()
        return text[text.index("<Piece"):text.index("</Piece>") + len("</Piece>")]

    both = str(tmp_path / "both.vtu")
    with open(both, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" '
                'header_type="UInt32">\n<UnstructuredGrid>\n' + piece(fa) + "\n" + piece(fb) +
                "\n</UnstructuredGrid>\n</VTKFile>\n")
    g = mesh_ingest.vtu(both)
    np.testing.assert_array_equal(g.GetLocations(), np.vstack([a[0], b[0]]).astype(np.float64))
    np.testing.assert_array_equal(g.GetVectorField("Velocity"), np.vstack([a[1]["Velocity"], b[1]["Velocity"]]))
    np.testing.assert_array_equal(g.GetCellPoints(101), b[2][0][4:8] + 12)
