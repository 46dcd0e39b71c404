"""Synthetic tetrahedral bar meshes + the .msh dialect DOT reads.

The benchmark configs of BASELINE.json name meshes that only exist under the
reference's `input/` directory (absent on the GPU box), so every workload here is
a structured bar: nx*ny*nz unit cells, each split into 6 Kuhn tets (SURVEY.md
section 8(d), config C4).  `bar17K_like` (64x15x15 cells -> 86,400 tets / 16,640
nodes) stands in for `bar17K.msh` (86,058 tets / 17,315 nodes); `bar1M`
(140x35x35 -> 1,029,000 tets / 182,736 nodes) is config C4 itself.

File format written by `write_msh` = what IglUtils::readTetMesh parses
(reference src/Utils/IglUtils.cpp:680-749).
"""
from __future__ import annotations

import itertools
import numpy as np

PRESETS = {
    # name: (nx, ny, nz)
    "bar_tiny": (6, 2, 2),        # 144 tets, 63 nodes   (golden fixtures)
    "bar_small": (12, 3, 3),      # 648 tets, 208 nodes  (golden fixtures)
    "bar2K": (24, 6, 6),          # 5,184 tets
    "bar5K_like": (32, 10, 10),   # 19,200 tets (bunny5K-sized: 19,379)
    "bar17K_like": (64, 15, 15),  # 86,400 tets (bar17K-sized: 86,058)
    "bar136K_like": (128, 29, 29),  # 645,888 tets (horse136K-sized: 641,852)
    "bar1M": (140, 35, 35),       # 1,029,000 tets (config C4)
}


def kuhn_bar(nx: int, ny: int, nz: int):
    """Return (V [nV,3] float64, T [nT,4] int32) of an nx*ny*nz-cell bar with unit cells
    of edge 1/ny (so the bar spans [0,nx/ny] x [0,1] x [0,nz/ny]).

    Cells are enumerated i-major (i, j, k), tets cell-major, the 6 tets of a cell in
    lexicographic order of the axis permutation; node id = (i*(ny+1)+j)*(nz+1)+k.
    Every tet has det(Dm) > 0.
    """
    e = np.eye(3, dtype=np.int64)
    locs = []
    for perm in itertools.permutations(range(3)):
        p0 = np.zeros(3, dtype=np.int64)
        p1 = e[perm[0]]
        p2 = p1 + e[perm[1]]
        p3 = np.ones(3, dtype=np.int64)
        Dm = np.stack([p1 - p0, p2 - p0, p3 - p0], axis=1).astype(float)
        if np.linalg.det(Dm) < 0:
            p1, p2 = p2, p1
        locs.append(np.stack([p0, p1, p2, p3]))
    locs = np.stack(locs)  # [6,4,3]
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cell = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1)  # i-major
    corner = cell[:, None, None, :] + locs[None]  # [nC,6,4,3]
    nid = (corner[..., 0] * (ny + 1) + corner[..., 1]) * (nz + 1) + corner[..., 2]
    T = nid.reshape(-1, 4).astype(np.int32)
    gi, gj, gk = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    V = np.stack([gi.ravel(), gj.ravel(), gk.ravel()], axis=1).astype(np.float64) / float(ny)
    return V, T


def surface_tris(T: np.ndarray) -> np.ndarray:
    """Boundary triangles with the vertex order buildSTri2Tet looks up
    (reference IglUtils.cpp:596-619): faces (v0,v2,v1),(v0,v3,v2),(v0,v1,v3),(v1,v2,v3)
    that occur exactly once."""
    f = np.concatenate([T[:, [0, 2, 1]], T[:, [0, 3, 2]], T[:, [0, 1, 3]], T[:, [1, 2, 3]]], axis=0)
    # keep the per-tet grouping order: tet-major
    order = np.arange(T.shape[0] * 4).reshape(4, -1).T.ravel()
    f = f[order]
    key = np.sort(f, axis=1)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    return f[cnt[inv.ravel()] == 1].astype(np.int32)


def normalise_like_loader(V: np.ndarray, size: float = 1.0) -> np.ndarray:
    """main.cpp:709-710: scale so the longest bbox edge is `size`, move min corner to 0."""
    V = V * (size / (V.max(axis=0) - V.min(axis=0)).max())
    return V - V.min(axis=0)


def write_msh(path: str, V: np.ndarray, T: np.ndarray, SF: np.ndarray | None = None) -> None:
    if SF is None:
        SF = surface_tris(T)
    lo, hi = V.min(axis=0), V.max(axis=0)
    with open(path, "w") as f:
        f.write("$MeshFormat\n4 0 8\n$EndMeshFormat\n$Entities\n0 0 0 1\n")
        f.write("0 %.17g %.17g %.17g %.17g %.17g %.17g 0 0\n" % (lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]))
        f.write("$EndEntities\n$Nodes\n1 %d\n0 3 0 %d\n" % (V.shape[0], V.shape[0]))
        f.write("".join("%d %.17g %.17g %.17g\n" % (i + 1, v[0], v[1], v[2]) for i, v in enumerate(V)))
        f.write("$EndNodes\n$Elements\n1 %d\n0 3 4 %d\n" % (T.shape[0], T.shape[0]))
        f.write("".join("%d %d %d %d %d\n" % (i + 1, t[0] + 1, t[1] + 1, t[2] + 1, t[3] + 1) for i, t in enumerate(T)))
        f.write("$EndElements\n$Surface\n%d\n" % SF.shape[0])
        f.write("".join("%d %d %d\n" % (s[0] + 1, s[1] + 1, s[2] + 1) for s in SF))
        f.write("$EndSurface\n")


def write_script(path: str, msh_path: str, energy: str = "SNH", parts: int = 8, anim: str = "twist",
                 duration: float = 5.0, dt: float = 0.025, density: float = 1000.0,
                 youngs: float = 1e5, poisson: float = 0.4, tol: float | None = None, stepper: str = "DOT") -> None:
    """A DOT script with the keys the shipped input/*.txt scripts use (reference Config.cpp:43-200)."""
    with open(path, "w") as f:
        # `timeStepper Newton` = Projected Newton; DOT and LBFGSJH take the partition count (Config.cpp:60-80)
        ts = "%s %d" % (stepper, parts) if stepper in ("DOT", "LBFGSJH") else stepper
        f.write("energy %s\ntimeStepper %s\ninexactSolve 0\nwarmStart 2\nresolution 1000\nsize 1\n" % (energy, ts))
        f.write("time %.17g %.17g\ndensity %.17g\nstiffness %.17g %.17g\nscript %s\n" % (duration, dt, density, youngs, poisson, anim))
        f.write("shape input %s\n" % msh_path)
        if tol is not None:
            f.write("tol 1\n%.17g\n" % tol)


def preset(name: str):
    nx, ny, nz = PRESETS[name]
    return kuhn_bar(nx, ny, nz)


def load_mesh_npz(path: str):
    """(V [nV,3] float64 as parsed from the .msh, T [nT,4] int32) of a mesh fixture written by oracle/gen_golden.py --meshes
    (the reference's own input/tetMeshes/*.msh, e.g. bar17K, bunny5K, horse38K)."""
    z = np.load(path)
    return np.ascontiguousarray(z["V"], dtype=np.float64), np.ascontiguousarray(z["T"], dtype=np.int32)
