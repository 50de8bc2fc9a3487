"""Gmsh MSH 2.2 ASCII reader / writer (numpy).

Test / bench infrastructure and mesh tooling only -- the product's own reader is
the C++ one in navier-stokes_equations_b200/host/mesh.cpp.

Semantics follow what the reference feeds to deal.II's GridIn::read_msh
(reference src/classes/NavierStokes.cpp:7-53):
  * '\r' stripped from every line (cpp:25-26),
  * a `$ParametricNodes` block is treated as `$Nodes`, keeping only the first
    four fields `id x y z` of every node line (cpp:28-47),
  * vertices get consecutive 0-based indices in file order of $Nodes,
  * elements of type 2 (triangle) / 4 (tetrahedron) become cells in file order,
    material id = first tag; lower dimensional elements (type 1 lines in 2-D,
    type 2 triangles in 3-D) carry the boundary id = first tag.
"""
from __future__ import annotations

import io
import numpy as np

# gmsh element type -> number of nodes
_NODES_PER_TYPE = {1: 2, 2: 3, 4: 4, 15: 1}


class Mesh:
    """Plain container.

    dim       : 2 or 3
    points    : (V, dim) float64
    cells     : (C, dim+1) int32, 0-based vertex indices, file order
    cell_tag  : (C,) int32 physical tag (201)
    faces     : (F, dim) int32 boundary elements (lines / triangles), file order
    face_tag  : (F,) int32 physical tag = boundary id (101..104)
    """

    def __init__(self, dim, points, cells, cell_tag, faces, face_tag, names=None):
        self.dim = int(dim)
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.cell_tag = np.ascontiguousarray(cell_tag, dtype=np.int32)
        self.faces = np.ascontiguousarray(faces, dtype=np.int32)
        self.face_tag = np.ascontiguousarray(face_tag, dtype=np.int32)
        self.names = names or []

    @property
    def n_vertices(self):
        return self.points.shape[0]

    @property
    def n_cells(self):
        return self.cells.shape[0]


def _normalised_lines(text: str):
    """Apply the reference's pre-pass (cpp:18-51) and yield lines."""
    in_param = False
    first = False
    for line in text.split("\n"):
        if line.endswith("\r"):
            line = line[:-1]
        if line == "$ParametricNodes":
            in_param, first = True, True
            yield "$Nodes"
        elif line == "$EndParametricNodes":
            in_param = False
            yield "$EndNodes"
        elif in_param:
            if first:
                first = False
                yield line
            else:
                f = line.split()
                yield " ".join(f[:4])
        else:
            yield line


def read_msh(path_or_text, dim=None) -> Mesh:
    if isinstance(path_or_text, str) and "\n" not in path_or_text:
        with open(path_or_text, "r") as fh:
            text = fh.read()
    else:
        text = path_or_text
    lines = list(_normalised_lines(text))
    i = 0
    names = []
    node_ids = None
    xyz = None
    elems = []
    while i < len(lines):
        s = lines[i].strip()
        if s == "$MeshFormat":
            ver = lines[i + 1].split()
            if not ver[0].startswith("2"):
                raise ValueError("only MSH 2.x ASCII is supported, got " + ver[0])
            if int(ver[1]) != 0:
                raise ValueError("binary MSH is not supported")
            i += 3
        elif s == "$PhysicalNames":
            n = int(lines[i + 1])
            for k in range(n):
                f = lines[i + 2 + k].split(None, 2)
                names.append((int(f[0]), int(f[1]), f[2].strip().strip('"')))
            i += n + 3
        elif s == "$Nodes":
            n = int(lines[i + 1])
            arr = np.loadtxt(io.StringIO("\n".join(lines[i + 2:i + 2 + n])), dtype=np.float64, ndmin=2)
            node_ids = arr[:, 0].astype(np.int64)
            xyz = arr[:, 1:4].copy()
            i += n + 3
        elif s == "$Elements":
            n = int(lines[i + 1])
            for k in range(n):
                f = lines[i + 2 + k].split()
                etype = int(f[1])
                ntags = int(f[2])
                tag = int(f[3]) if ntags > 0 else 0
                nn = _NODES_PER_TYPE.get(etype)
                if nn is None:
                    raise ValueError("unsupported gmsh element type %d" % etype)
                conn = [int(v) for v in f[3 + ntags:3 + ntags + nn]]
                elems.append((etype, tag, conn))
            i += n + 3
        else:
            i += 1
    if xyz is None:
        raise ValueError("no $Nodes section")
    # gmsh node ids -> consecutive 0-based indices in file order
    id2idx = {int(g): k for k, g in enumerate(node_ids)}
    has_tet = any(e[0] == 4 for e in elems)
    if dim is None:
        dim = 3 if has_tet else 2
    ctype, ftype = (4, 2) if dim == 3 else (2, 1)
    cells = [[id2idx[v] for v in c] for (t, tag, c) in elems if t == ctype]
    ctag = [tag for (t, tag, c) in elems if t == ctype]
    faces = [[id2idx[v] for v in c] for (t, tag, c) in elems if t == ftype]
    ftag = [tag for (t, tag, c) in elems if t == ftype]
    cells = np.array(cells, dtype=np.int32).reshape(-1, dim + 1)
    faces = np.array(faces, dtype=np.int32).reshape(-1, dim)
    return Mesh(dim, xyz[:, :dim], cells, np.array(ctag, np.int32), faces, np.array(ftag, np.int32), names)


_DEFAULT_NAMES_2D = [(1, 101, "inlet"), (1, 102, "outlet"), (1, 103, "walls"), (1, 104, "cylinder"), (2, 201, "fluid")]
_DEFAULT_NAMES_3D = [(2, 101, "inlet"), (2, 102, "outlet"), (2, 103, "cylinder"), (2, 104, "walls"), (3, 201, "fluid")]


def write_msh(path, mesh: Mesh):
    """MSH 2.2 ASCII with the tag conventions of reference meshes/*.geo."""
    dim = mesh.dim
    names = mesh.names or (_DEFAULT_NAMES_2D if dim == 2 else _DEFAULT_NAMES_3D)
    ctype, ftype = (4, 2) if dim == 3 else (2, 1)
    pts = np.zeros((mesh.n_vertices, 3))
    pts[:, :dim] = mesh.points
    with open(path, "w") as fh:
        fh.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n")
        fh.write("$PhysicalNames\n%d\n" % len(names))
        for d, tag, nm in names:
            fh.write('%d %d "%s"\n' % (d, tag, nm))
        fh.write("$EndPhysicalNames\n$Nodes\n%d\n" % mesh.n_vertices)
        buf = io.StringIO()
        for k in range(mesh.n_vertices):
            buf.write("%d %.17g %.17g %.17g\n" % (k + 1, pts[k, 0], pts[k, 1], pts[k, 2]))
        fh.write(buf.getvalue())
        fh.write("$EndNodes\n$Elements\n%d\n" % (mesh.faces.shape[0] + mesh.n_cells))
        buf = io.StringIO()
        eid = 1
        for f, t in zip(mesh.faces, mesh.face_tag):
            buf.write("%d %d 2 %d %d %s\n" % (eid, ftype, t, t, " ".join(str(int(v) + 1) for v in f)))
            eid += 1
        for c, t in zip(mesh.cells, mesh.cell_tag):
            buf.write("%d %d 2 %d %d %s\n" % (eid, ctype, t, t, " ".join(str(int(v) + 1) for v in c)))
            eid += 1
        fh.write(buf.getvalue())
        fh.write("$EndElements\n")


def save_npz(path, mesh: Mesh):
    np.savez_compressed(path, dim=np.int32(mesh.dim), points=mesh.points, cells=mesh.cells,
                        cell_tag=mesh.cell_tag, faces=mesh.faces, face_tag=mesh.face_tag)


def load_npz(path) -> Mesh:
    z = np.load(path)
    return Mesh(int(z["dim"]), z["points"], z["cells"], z["cell_tag"], z["faces"], z["face_tag"])


def write_bin(path, mesh: Mesh):
    """Little-endian binary dump read by the C++ host (mesh.cpp: read_bin) -- avoids
    a ~300 MB ASCII file for the mesh-3D-20-equivalent.  Layout:
    magic 'NSBMESH1' | int32 dim, V, C, F | f64 points[V][dim] | i32 cells[C][dim+1]
    | i32 cell_tag[C] | i32 faces[F][dim] | i32 face_tag[F]."""
    with open(path, "wb") as fh:
        fh.write(b"NSBMESH1")
        np.array([mesh.dim, mesh.n_vertices, mesh.n_cells, mesh.faces.shape[0]], dtype="<i4").tofile(fh)
        mesh.points.astype("<f8").tofile(fh)
        mesh.cells.astype("<i4").tofile(fh)
        mesh.cell_tag.astype("<i4").tofile(fh)
        mesh.faces.astype("<i4").tofile(fh)
        mesh.face_tag.astype("<i4").tofile(fh)
