"""Caches the graded 2-D cross-section triangulations of the mesh-3D-<level>-equivalents (the DistMesh iteration is the
slow part of tools/meshgen.mesh_3d; the extrusion is a few numpy calls).  tools/meshgen.cached_cross_section() picks the
files up; the meshes are bit-identical with or without the cache (the generator is deterministic).
  python tools/make_cross_sections.py 20 40"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import meshgen  # noqa: E402

if __name__ == "__main__":
    out = os.path.join(ROOT, "tools", "cross_sections")
    os.makedirs(out, exist_ok=True)
    for lv in [int(a) for a in sys.argv[1:]] or [20, 40]:
        t0 = time.time()
        lc_cyl, lc_global = meshgen.LEVELS_3D[lv]
        p2, t2, be2, tag2 = meshgen.cross_section(lc_cyl, lc_global)
        np.savez_compressed(os.path.join(out, "level%d.npz" % lv), p2=p2, t2=t2, be2=be2, tag2=tag2)
        print("level %d: %d points, %d triangles, %.0f s" % (lv, p2.shape[0], t2.shape[0], time.time() - t0), flush=True)
