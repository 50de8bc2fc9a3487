"""Convert the reference's shipped 2-D meshes (/root/reference/meshes/*.msh) into compact
.npz fixtures under tests/golden/ (the reference tree does not exist on the GPU box).
Run once in the build container:  python -m tools.make_golden_meshes
Input data only -- no reference source code is copied."""
import os, sys
import numpy as np
from tools import msh

REF = "/root/reference/meshes"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

def main():
    os.makedirs(OUT, exist_ok=True)
    for name in ("mesh-2D", "mesh-2D-40", "mesh-2D-100"):
        m = msh.read_msh(os.path.join(REF, name + ".msh"))
        p = m.points[m.cells]
        area2 = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 2, 0] - p[:, 0, 0]) * (p[:, 1, 1] - p[:, 0, 1])
        print(name, "V", m.n_vertices, "C", m.n_cells, "F", m.faces.shape[0], "tags", np.unique(m.face_tag),
              "min 2*area", area2.min(), "sum area", 0.5 * area2.sum())
        msh.save_npz(os.path.join(OUT, name + ".npz"), m)

if __name__ == "__main__":
    main()
