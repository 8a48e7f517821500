"""The C++ host shell's script parser, ecomsol reader and .vtk writers (wholecellfe_b200/csrc/host) against the reference's
formats: scripts written like cellBP/scripts/cell_example.txt, meshes in the ecomsol layout, and -- byte for byte -- .vtk
frames written by the reference's own writers (golden, tests/golden/make_golden.py).  CPU only."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from conftest import GOLDEN
from wholecellfe_b200.meshio import read_ecomsol, write_ecomsol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "wholecellfe_b200", "csrc", "host")
_dp = C.POINTER(C.c_double); _ip = C.POINTER(C.c_int)

SCRIPT = """relative_output_path cell_subs_example

frame_dt 0.04
sim_dt 0.005    # a comment
end_sim_t 4.0

surftrack_parameters
(
   use_fraction 0
   min_edge_length 0.025
   max_edge_length 2
   perform_improvement %d
   allow_topology_changes 0
   collision_safety 0
   subdivision_scheme butterfly
)

ecomsolfile
(
   filepath %s
)

camera
(
   target [0.0 0.0 0.0]
   distance 1.5
)
"""


@pytest.fixture(scope="module")
def host():
    subprocess.check_call(["make", "-s", "-C", HOST, os.path.join(HOST, "libcellbp_host.so")])
    L = C.CDLL(os.path.join(HOST, "libcellbp_host.so"))
    L.cellbp_last_error.restype = C.c_char_p
    return L


def d(a):
    return a.ctypes.data_as(_dp)


def test_script_parser_reads_the_reference_keys(host, tmp_path):
    mesh = tmp_path / "m.txt"
    p = tmp_path / "cell.txt"
    p.write_text(SCRIPT % (1, mesh))
    out = np.zeros(9); op = C.create_string_buffer(512); mp = C.create_string_buffer(512)
    assert host.cellbp_read_script(str(p).encode(), d(out), op, 512, mp, 512) == 0
    assert list(out) == [0.04, 0.005, 4.0, 0.0, 0.0, 1.0, 1.0, 0.0, 0.0]
    assert op.value == b"cell_subs_example" and mp.value == str(mesh).encode()
    # num_substeps alone, absolute output path, curr_t
    p.write_text("frame_dt 0.1\nnum_substeps 4\noutput_path /tmp/xyz\ncurr_t 0.3\necomsolfile ( filepath a.txt )\n")
    assert host.cellbp_read_script(str(p).encode(), d(out), op, 512, mp, 512) == 0
    assert out[0] == 0.1 and out[1] == 0.025 and out[3] == 0.3 and out[4] == 1.0 and out[5] == 0.0 and out[2] == 1e30
    assert op.value == b"/tmp/xyz"
    # malformed: missing ')' and missing frame_dt
    p.write_text("frame_dt 0.1\nsim_dt 0.01\nsurftrack_parameters ( perform_improvement 0\n")
    assert host.cellbp_read_script(str(p).encode(), d(out), op, 512, mp, 512) != 0
    p.write_text("sim_dt 0.01\n")
    assert host.cellbp_read_script(str(p).encode(), d(out), op, 512, mp, 512) != 0
    assert host.cellbp_read_script(b"/nonexistent/script.txt", d(out), op, 512, mp, 512) != 0


def test_ecomsol_reader_matches_python_reader_and_forces_z_to_zero(host, meshes, tmp_path):
    xc, tc, _, _ = meshes
    path = str(tmp_path / "disk.txt")
    write_ecomsol(path, xc, tc)
    nn = C.c_int(); ne = C.c_int()
    assert host.cellbp_read_ecomsol(path.encode(), C.byref(nn), C.byref(ne)) == 0
    assert (nn.value, ne.value) == (len(xc), len(tc))
    xyz = np.zeros((nn.value, 3)); tri = np.zeros((ne.value, 3), dtype=np.int32)
    host.cellbp_get_mesh(d(xyz), tri.ctypes.data_as(_ip))
    px, pt = read_ecomsol(path)
    assert np.array_equal(xyz, px) and np.array_equal(tri, pt) and np.array_equal(xyz, xc) and np.array_equal(tri, tc)
    bad = tmp_path / "bad.txt"; bad.write_text("hello\n")
    assert host.cellbp_read_ecomsol(str(bad).encode(), C.byref(nn), C.byref(ne)) != 0


@pytest.mark.parametrize("variant,kint", [("k1", 1000.0), ("kMD", -1.0)])
def test_vtk_frames_byte_identical_to_the_reference_writers(host, meshes, tmp_path, variant, kint):
    xc, tc, xs, ts = meshes
    g = np.load(os.path.join(GOLDEN, "vtk_frames.npz"))
    x = np.ascontiguousarray(g[variant + "_c_x"]); cA = np.ascontiguousarray(g[variant + "_c_cA"]); cC = np.ascontiguousarray(g[variant + "_c_cC"])
    fa = np.ascontiguousarray(g[variant + "_c_fa_u"]); ea = np.ascontiguousarray(g[variant + "_c_ea"]); eas = np.ascontiguousarray(g[variant + "_s_ea"])
    t = float(g[variant + "_t"])
    tc32 = np.ascontiguousarray(tc, dtype=np.int32); ts32 = np.ascontiguousarray(ts, dtype=np.int32)
    pc = str(tmp_path / "c.vtk"); ps = str(tmp_path / "s.vtk")
    assert host.cellbp_write_vtk_cell(pc.encode(), C.c_double(t), len(x), d(x), len(tc32), tc32.ctypes.data_as(_ip), d(cA), d(cC), d(fa), d(ea), C.c_double(kint)) == 0
    xs_c = np.ascontiguousarray(xs)
    assert host.cellbp_write_vtk_subs(ps.encode(), C.c_double(t), len(xs_c), d(xs_c), len(ts32), ts32.ctypes.data_as(_ip), d(eas), len(fa), d(fa)) == 0
    want_c = bytes(g[variant + "_cell_vtk"]); want_s = bytes(g[variant + "_subs_vtk"])
    got_c = open(pc, "rb").read(); got_s = open(ps, "rb").read()
    assert got_s == want_s
    assert got_c == want_c      # theta_e in Eigen's own operation order (3x3 inverse by cofactors, then its determinant): every digit
    assert host.cellbp_write_vtk_cell(b"/nonexistent/dir/c.vtk", C.c_double(t), len(x), d(x), len(tc32), tc32.ctypes.data_as(_ip), d(cA), d(cC), d(fa), d(ea), C.c_double(kint)) != 0
