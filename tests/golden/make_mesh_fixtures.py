#!/usr/bin/env python
"""Extract the geometry of the two OBJ assets present in the reference checkout into small binary
fixtures (positions + polygon indices only; no normals / texcoords / materials):

    /root/reference/RT_Metal/coatball/coatball.obj  -> tests/golden/meshes/coatball.npz
    /root/reference/RT_Metal/meshes/teapot.obj      -> tests/golden/meshes/teapot.npz

These are INPUT fixtures for BASELINE configs C2-C4: /root/reference does not
exist on the GPU box, so the bench and the GPU tests read the fixtures instead. Run here, once:
    python tests/golden/make_mesh_fixtures.py
"""
import os
import sys

import numpy as np

REF = "/root/reference/RT_Metal"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "meshes")


def read_obj(path):
    verts, faces = [], []
    with open(path, "r", errors="replace") as f:
        for line in f:
            if line.startswith("v "):
                p = line.split()
                verts.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("f "):
                idx = [int(tok.split("/")[0]) for tok in line.split()[1:]]
                faces.append([i - 1 if i > 0 else len(verts) + i for i in idx])
    v = np.asarray(verts, dtype=np.float32)
    tris = []
    for f in faces:                      # fan triangulation (0,1,2),(0,2,3),... as ModelIO does for quads
        for k in range(1, len(f) - 1):
            tris.append((f[0], f[k], f[k + 1]))
    t = np.asarray(tris, dtype=np.int32)
    # drop unreferenced vertices, keep order
    used = np.zeros(len(v), dtype=bool)
    used[t.reshape(-1)] = True
    remap = np.cumsum(used) - 1
    return v[used], remap[t].astype(np.int32), len(faces)


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, rel in (("coatball", "coatball/coatball.obj"), ("teapot", "meshes/teapot.obj")):
        v, t, nf = read_obj(os.path.join(REF, rel))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), v=v, tris=t)
        print(f"{name}: {len(v)} vertices, {nf} faces -> {len(t)} triangles")


if __name__ == "__main__":
    sys.exit(main())
