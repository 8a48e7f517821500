"""Synthetic meshes for the large configurations (BASELINE.json configs 3-5): the reference ships
nothing beyond 5 054 triangles, so the 100 k / 1 M-triangle cells and the 4 M-node substrate are generated.

Both generators return (xyz [N,3] float64 with z == 0, tri [M,3] int32, counter-clockwise like the shipped meshes)
numbered along a Morton (Z-order) curve so that one-ring gathers on the GPU stay local in L2.
"""
import numpy as np


def _morton_key(ix, iy):
    def spread(v):
        v = v.astype(np.uint64)
        v = (v | (v << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x3333333333333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x5555555555555555)
        return v
    return spread(ix) | (spread(iy) << np.uint64(1))


def _renumber(xyz, tri, order="morton"):
    if order is None:
        return xyz, tri
    lo = xyz[:, :2].min(axis=0); span = np.maximum(xyz[:, :2].max(axis=0) - lo, 1e-300)
    q = np.minimum(((xyz[:, :2] - lo) / span * 65535.0).astype(np.int64), 65535)
    perm = np.argsort(_morton_key(q[:, 0], q[:, 1]), kind="stable")       # new -> old
    if order == "morton_rows":
        # inside every chunk of `chunk` consecutive Morton nodes (= one GPU patch) re-sort by rows (y band, then x): neighbouring
        # nodes then sit in neighbouring slots of the patch's shared-memory node table, which is what keeps the one-ring gathers
        # of the patch kernels free of bank conflicts
        chunk = 128
        p2 = perm.copy()
        for b in range(0, len(perm), chunk):
            ids = perm[b:b + chunk]
            pts = xyz[ids, :2]
            hrow = np.sqrt(max((pts[:, 0].max() - pts[:, 0].min()) * (pts[:, 1].max() - pts[:, 1].min()), 1e-300) / max(len(ids), 1))
            row = np.floor((pts[:, 1] - pts[:, 1].min()) / max(hrow, 1e-300) + 0.5)
            p2[b:b + chunk] = ids[np.lexsort((pts[:, 0], row))]
        perm = p2
    inv = np.empty_like(perm); inv[perm] = np.arange(len(perm))
    xyz = xyz[perm]
    tri = inv[tri].astype(np.int32)
    cen = xyz[tri].mean(axis=1)
    qc = np.minimum(((cen[:, :2] - lo) / span * 65535.0).astype(np.int64), 65535)
    tri = tri[np.argsort(_morton_key(qc[:, 0], qc[:, 1]), kind="stable")]
    return xyz, np.ascontiguousarray(tri)


def square_mesh(n_side, half=50.0, order="morton", jitter=0.0, seed=1234, diagonals="alternate"):
    """Structured right-triangle mesh of [-half, half]^2 with n_side x n_side nodes (2 (n_side-1)^2 triangles).
    diagonals="alternate" (default): union-jack pattern, interior valences 4 and 8; "same": every diagonal the same way,
    interior valence 6 everywhere (a diagnostic: what the node-centric kernels do without valence divergence inside a warp)."""
    g = np.linspace(-half, half, n_side)
    X, Y = np.meshgrid(g, g, indexing="xy")
    xyz = np.stack([X.ravel(), Y.ravel(), np.zeros(n_side * n_side)], axis=1)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        h = 2 * half / (n_side - 1)
        inner = (np.abs(xyz[:, 0]) < half - 1e-9) & (np.abs(xyz[:, 1]) < half - 1e-9)
        xyz[inner, :2] += jitter * h * rng.uniform(-1, 1, size=(int(inner.sum()), 2))
    i, j = np.meshgrid(np.arange(n_side - 1), np.arange(n_side - 1), indexing="xy")
    a = (j * n_side + i).ravel(); b = a + 1; c = a + n_side; d = c + 1
    flip = ((i + j) % 2 == 0).ravel() if diagonals == "alternate" else np.zeros(i.size, dtype=bool)   # alternate the diagonal
    t1 = np.where(flip[:, None], np.stack([a, b, d], 1), np.stack([a, b, c], 1))
    t2 = np.where(flip[:, None], np.stack([a, d, c], 1), np.stack([b, d, c], 1))
    tri = np.concatenate([t1, t2]).astype(np.int32)
    return _renumber(xyz, tri, order)


def disk_mesh(n_tri_target, radius=10.0, center=(0.05, 0.05), order="morton"):
    """Disk of the reference cell's size (20 um diameter, centred like scripts/disk_20um_diam_edited.txt) with about
    n_tri_target triangles: concentric rings, ring k has 6k nodes, fan-triangulated (all counter-clockwise)."""
    nr = max(2, int(round(np.sqrt(n_tri_target / 6.0))))
    pts = [(0.0, 0.0)]
    start = [0]
    for k in range(1, nr + 1):
        start.append(len(pts))
        r = radius * k / nr
        th = 2 * np.pi * np.arange(6 * k) / (6 * k)
        pts.extend(zip(r * np.cos(th), r * np.sin(th)))
    tris = []
    for k in range(1, nr + 1):
        n_out = 6 * k
        n_in = 6 * (k - 1)
        o0 = start[k]
        i0 = start[k - 1]
        for s in range(6):                      # six sectors: k outer segments over k-1 inner segments
            for j in range(k):
                inner = i0 if k == 1 else i0 + (s * (k - 1) + j) % n_in
                tris.append((inner, o0 + (s * k + j) % n_out, o0 + (s * k + j + 1) % n_out))
            for j in range(k - 1):
                tris.append((i0 + (s * (k - 1) + j) % n_in, o0 + (s * k + j + 1) % n_out, i0 + (s * (k - 1) + j + 1) % n_in))
    xyz = np.zeros((len(pts), 3)); xyz[:, :2] = np.array(pts); xyz[:, 0] += center[0]; xyz[:, 1] += center[1]
    tri = np.array(tris, dtype=np.int32)
    return _renumber(xyz, tri, order)
