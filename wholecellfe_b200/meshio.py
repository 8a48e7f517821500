"""Host-side mesh I/O in the reference's "edited COMSOL" format.

Format (reference reader: cellBP/iomesh.cpp:491-565): a line containing `Nodes`, a line with N, N lines
`x y [z]` (z is forced to 0 for planar problems, iomesh.cpp:523), a line containing `Elements`, a line
with M, M lines `i j k` (0-based vertex ids).
"""
import numpy as np


def read_ecomsol(path):
    """Returns (xyz float64 [N,3] with z == 0, tri int32 [M,3])."""
    with open(path) as f:
        lines = f.read().split("\n")
    xyz = tri = None
    i = 0
    while i < len(lines):
        if "Nodes" in lines[i]:
            n = int(lines[i + 1].split()[0])
            xyz = np.zeros((n, 3))
            for k in range(n):
                t = lines[i + 2 + k].split()
                xyz[k, 0] = float(t[0]); xyz[k, 1] = float(t[1])
            i += 2 + n
            continue
        if "Elements" in lines[i]:
            m = int(lines[i + 1].split()[0])
            tri = np.zeros((m, 3), dtype=np.int32)
            for k in range(m):
                t = lines[i + 2 + k].split()
                tri[k] = (int(t[0]), int(t[1]), int(t[2]))
            i += 2 + m
            continue
        i += 1
    if xyz is None or tri is None:
        raise ValueError("%s: not an ecomsol mesh (Nodes/Elements blocks missing)" % path)
    return xyz, tri


def write_ecomsol(path, xyz, tri):
    xyz = np.asarray(xyz); tri = np.asarray(tri)
    with open(path, "w") as f:
        f.write("Nodes\n%d\n" % len(xyz))
        for p in xyz:
            f.write("%.17g %.17g \n" % (p[0], p[1]))
        f.write("Elements\n%d\n" % len(tri))
        for t in tri:
            f.write("%d %d %d \n" % (t[0], t[1], t[2]))
