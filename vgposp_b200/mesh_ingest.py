"""Mesh ingest: VTK XML unstructured grids (.vtu / .pvtu) -> coordinate and feature arrays.

The reference reads its room-simulation time steps through the `vtk` package (`vtktools.py:11-30` the `vtu` class,
`:32-45` GetScalarField, `:62-75` GetVectorField, `:97-118` GetField, `:120-143` GetFieldRank, `:277-284`
GetLocations, `:286-290` GetCellPoints, `:292-295` GetFieldNames) and stacks points and features over the time steps
(`gp_pvtk.py:24-75`, `data_readers.py:50-140`).  `vtk` is not a dependency here: this module parses the XML container
itself -- ascii, inline base64 ("binary") and appended (raw or base64) arrays, UInt32 / UInt64 block headers, either
byte order, zlib-compressed blocks, several pieces per file and .pvtu piece lists -- and offers the same accessors
with the same names, shapes and error messages.  Host-side only; what it returns (float64 [n, 3] coordinates,
[n] / [n, c] features) is what the kernel-matrix builder and the VGP trainer take.
"""
import base64
import os
import re
import xml.etree.ElementTree as ET
import zlib

import numpy as np

_DTYPES = {"Int8": "i1", "UInt8": "u1", "Int16": "i2", "UInt16": "u2", "Int32": "i4", "UInt32": "u4",
           "Int64": "i8", "UInt64": "u8", "Float32": "f4", "Float64": "f8"}


def _b64_chars(nbytes):
    return (nbytes + 2) // 3 * 4


class _Container:
    """Decoding state of one VTKFile element: byte order, header type, compressor, appended payload."""

    def __init__(self, root, appended_raw):
        self.order = "<" if root.get("byte_order", "LittleEndian") == "LittleEndian" else ">"
        self.header = np.dtype(self.order + _DTYPES[root.get("header_type", "UInt32")])
        comp = root.get("compressor")
        if comp not in (None, "", "vtkZLibDataCompressor"):
            raise Exception("ERROR: unsupported compressor " + comp)
        self.zlib = bool(comp)
        self.appended, self.appended_b64 = None, False
        node = root.find("AppendedData")
        if appended_raw is not None:
            self.appended = appended_raw
        elif node is not None:
            text = (node.text or "").strip()
            self.appended = text[text.index("_") + 1:] if "_" in text else text
            self.appended_b64 = True
            self.appended = "".join(self.appended.split())

    # -- one array's bytes from a base64 stream (header [+ block table] first, then the blocks)
    def _from_b64(self, text):
        hb = self.header.itemsize
        if not self.zlib:
            first = text[:_b64_chars(hb)]
            nbytes = int(np.frombuffer(base64.b64decode(first)[:hb], self.header)[0])
            if first.endswith("="):                       # header encoded on its own, data in a second stream
                return base64.b64decode(text[_b64_chars(hb):_b64_chars(hb) + _b64_chars(nbytes)])[:nbytes]
            return base64.b64decode(text[:_b64_chars(hb + nbytes)])[hb:hb + nbytes]     # one stream for both
        head = np.frombuffer(base64.b64decode(text[:_b64_chars(3 * hb)])[:3 * hb], self.header)
        nblocks = int(head[0])
        table_chars = _b64_chars((3 + nblocks) * hb)
        table = np.frombuffer(base64.b64decode(text[:table_chars])[:(3 + nblocks) * hb], self.header)
        sizes = [int(v) for v in table[3:]]
        body = base64.b64decode(text[table_chars:table_chars + _b64_chars(sum(sizes))])
        return self._inflate(body, sizes)

    def _from_raw(self, buf):
        hb = self.header.itemsize
        if not self.zlib:
            nbytes = int(np.frombuffer(buf[:hb], self.header)[0])
            return buf[hb:hb + nbytes]
        nblocks = int(np.frombuffer(buf[:hb], self.header)[0])
        table = np.frombuffer(buf[:(3 + nblocks) * hb], self.header)
        sizes = [int(v) for v in table[3:]]
        return self._inflate(buf[(3 + nblocks) * hb:], sizes)

    @staticmethod
    def _inflate(body, sizes):
        out, pos = [], 0
        for s in sizes:
            out.append(zlib.decompress(body[pos:pos + s]))
            pos += s
        return b"".join(out)

    def array(self, node):
        """The DataArray element as a numpy array [tuples, components] (components squeezed when 1)."""
        dtype = np.dtype(self.order + _DTYPES[node.get("type")])
        fmt = node.get("format", "ascii")
        if fmt == "ascii":
            a = np.array((node.text or "").split(), dtype=np.float64 if dtype.kind == "f" else np.int64).astype(dtype)
        elif fmt == "binary":
            a = np.frombuffer(self._from_b64("".join((node.text or "").split())), dtype)
        elif fmt == "appended":
            if self.appended is None:
                raise Exception("ERROR: appended array without an AppendedData section")
            off = int(node.get("offset", "0"))
            a = np.frombuffer(self._from_b64(self.appended[off:]) if self.appended_b64
                              else self._from_raw(self.appended[off:]), dtype)
        else:
            raise Exception("ERROR: unknown DataArray format " + fmt)
        nc = int(node.get("NumberOfComponents", "1"))
        a = a.astype(dtype.newbyteorder("="))
        return a.reshape(-1, nc) if nc > 1 else a


def _parse_file(filename):
    """-> list of pieces: dict(points [n, 3], point_data {name: array}, cell_data {...}, cells (conn, offsets, types))."""
    with open(filename, "rb") as f:
        blob = f.read()
    appended_raw = None
    m = re.search(rb"<AppendedData[^>]*encoding\s*=\s*\"raw\"[^>]*>", blob)
    if m:                                                  # raw bytes are not XML: cut them out before parsing
        start = blob.index(b"_", m.end()) + 1
        end = blob.rindex(b"</AppendedData>")
        appended_raw = blob[start:end]
        blob = blob[:m.end()] + b"_" + blob[end:]
    root = ET.fromstring(blob)
    if root.tag != "VTKFile":
        raise Exception("ERROR: not a VTK XML file: " + filename)
    kind = root.get("type")
    if kind == "PUnstructuredGrid":
        pieces = []
        here = os.path.dirname(os.path.abspath(filename))
        for p in root.find("PUnstructuredGrid").findall("Piece"):
            pieces.extend(_parse_file(os.path.join(here, p.get("Source"))))
        return pieces
    if kind != "UnstructuredGrid":
        raise Exception("ERROR: not an unstructured grid: " + filename)
    box = _Container(root, appended_raw)
    pieces = []
    for p in root.find("UnstructuredGrid").findall("Piece"):
        npts = int(p.get("NumberOfPoints", "0"))
        piece = {"points": np.zeros((0, 3)), "point_data": {}, "cell_data": {}, "cells": None}
        pts = p.find("Points")
        if pts is not None and pts.find("DataArray") is not None and npts:
            piece["points"] = np.asarray(box.array(pts.find("DataArray")), dtype=np.float64).reshape(npts, -1)
        for tag, key in (("PointData", "point_data"), ("CellData", "cell_data")):
            sec = p.find(tag)
            if sec is not None:
                for node in sec.findall("DataArray"):
                    piece[key][node.get("Name")] = box.array(node)
        cells = p.find("Cells")
        if cells is not None:
            named = {node.get("Name"): box.array(node) for node in cells.findall("DataArray")}
            if "connectivity" in named and "offsets" in named:
                piece["cells"] = (named["connectivity"].astype(np.int64), named["offsets"].astype(np.int64),
                                  named.get("types"))
        pieces.append(piece)
    return pieces


class vtu:
    """Unstructured grid read from a .vtu / .pvtu file; the accessors of `vtktools.vtu` that the reference's data path
    uses (`vtktools.py:11-30`), with numpy arrays in place of vtk objects."""

    def __init__(self, filename=None):
        self.filename = filename
        self.points = np.zeros((0, 3))
        self.point_data, self.cell_data = {}, {}
        self.connectivity, self.offsets = np.zeros(0, np.int64), np.zeros(0, np.int64)
        if filename is None:
            return
        if not (filename[-4:] == ".vtu" or filename[-5:] == ".pvtu"):
            raise Exception("ERROR: don't recognise file extension" + filename)
        pieces = _parse_file(filename)
        if pieces:
            self.points = np.vstack([p["points"] for p in pieces]) if any(len(p["points"]) for p in pieces) \
                else np.zeros((0, 3))
            for key in ("point_data", "cell_data"):
                names = []
                for p in pieces:
                    names += [n for n in p[key] if n not in names]
                merged = {n: np.concatenate([p[key][n] for p in pieces if n in p[key]]) for n in names}
                setattr(self, key, merged)
            conn, offs, base_pt, base_off = [], [], 0, 0
            for p in pieces:
                if p["cells"] is not None:
                    conn.append(p["cells"][0] + base_pt)
                    offs.append(p["cells"][1] + base_off)
                    base_off += len(p["cells"][0])
                base_pt += len(p["points"])
            if conn:
                self.connectivity, self.offsets = np.concatenate(conn), np.concatenate(offs)
        if len(self.points) + len(self.offsets) == 0:
            raise Exception("ERROR: No points or cells found after loading vtu " + filename)

    # -- lookups: point data first, then cell data (vtktools.py:34-45)
    def _find(self, name, what):
        if name in self.point_data:
            return self.point_data[name]
        if name in self.cell_data:
            return self.cell_data[name]
        raise Exception("ERROR: couldn't find point or cell %s data with name %s in file %s." %
                        (what, name, self.filename))

    def GetScalarField(self, name):
        """[n] values of a one-component field (first component otherwise, like GetTuple1) -- vtktools.py:32-45."""
        a = self._find(name, "scalar field")
        return np.array(a if a.ndim == 1 else a[:, 0], dtype=np.float64)

    def GetScalarRange(self, name):
        a = self.GetScalarField(name)
        return (float(a.min()), float(a.max()))

    def GetVectorField(self, name):
        """[n, 3] values of a three-component field -- vtktools.py:62-75."""
        a = self._find(name, "vector field")
        if a.ndim != 2 or a.shape[1] != 3:
            raise Exception("ERROR: field %s does not have three components" % name)
        return np.array(a, dtype=np.float64)

    def GetVectorNorm(self, name):
        return np.sqrt(np.sum(self.GetVectorField(name) ** 2, axis=1))

    def GetField(self, name):
        """[n, c] in the field's own type; 9 components -> [n, 3, 3], 4 -> [n, 2, 2] -- vtktools.py:97-118."""
        a = self._find(name, "field")
        a = a.reshape(len(a), -1)
        nc = a.shape[1]
        if nc == 9:
            return a.reshape(-1, 3, 3).copy()
        if nc == 4:
            return a.reshape(-1, 2, 2).copy()
        return a.copy()

    def GetFieldRank(self, name):
        a = self._find(name, "field")
        comps = 1 if a.ndim == 1 else a.shape[1]
        if comps == 1:
            return 0
        if comps in (2, 3):
            return 1
        if comps in (4, 9):
            return 2
        raise Exception("Field rank > 2 encountered")

    def GetFieldNames(self):
        """Names of the point-data arrays, in file order -- vtktools.py:292-295."""
        return list(self.point_data)

    def GetLocations(self):
        """[n, 3] node coordinates, float64 -- vtktools.py:277-284."""
        return np.array(self.points[:, :3], dtype=np.float64)

    def GetCellPoints(self, id):
        """Node numbers of cell `id` -- vtktools.py:286-290."""
        lo = 0 if id == 0 else int(self.offsets[id - 1])
        return np.array(self.connectivity[lo:int(self.offsets[id])])


def load_time_steps(filenames, scalar_fields=("Tracer", "Pressure", "Temperature", "Time", "Density"),
                    vector_fields=("Velocity",), max_steps=None):
    """Stack the points and the named features of a sequence of time-step files the way `gp_pvtk.py:24-75` /
    `data_readers.py:97-140` do: missing or empty files are skipped, coordinates are vstacked to [sum n_i, 3], every
    scalar field to [sum n_i, 1], vector fields to [sum n_i, 3].  Returns (xyz, {name: array}, files_used)."""
    xyz, feats, used = [], {n: [] for n in tuple(scalar_fields) + tuple(vector_fields)}, []
    for fn in filenames:
        if max_steps is not None and len(used) >= max_steps:
            break
        if (not os.path.isfile(fn)) or os.stat(fn).st_size == 0:
            continue
        g = vtu(fn)
        xyz.append(g.GetLocations())
        for n in scalar_fields:
            feats[n].append(g.GetScalarField(n).reshape(-1, 1))
        for n in vector_fields:
            feats[n].append(g.GetVectorField(n))
        used.append(fn)
    if not used:
        return np.zeros((0, 3)), {n: np.zeros((0, 1)) for n in feats}, used
    return np.vstack(xyz), {n: np.vstack(v) for n, v in feats.items()}, used
