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


REFAPP = os.path.join(ROOT, "oracle", "_ref", "cellBP_ref_k1")


@pytest.mark.skipif(not os.path.exists(REFAPP), reason="oracle/_ref/cellBP_ref_k1 not built here (oracle/build_app.sh)")
def test_shipped_example_with_eltopo_on_equals_the_eltopo_free_run_for_k1(host, meshes, tmp_path):
    """What `cellbp_gpu --no-eltopo` rests on: the UNMODIFIED reference application run on the shipped example with the shipped
    surftrack_parameters (perform_improvement 1, allow_topology_changes 1, collision_safety 1: El Topo fully on) writes, for k1, the
    same frames byte for byte as the El-Topo-free step sequence -- El Topo never touches that mesh (296 / 542 for all 800
    steps, actual_dt == dt).  20 frames = 160 steps here; all 100 frames were compared when the flag was added."""
    from oracle.oracle import OracleSim
    xc, tc, xs, ts = meshes
    (tmp_path / "scripts").mkdir()
    write_ecomsol(str(tmp_path / "scripts" / "disk.txt"), xc, tc); write_ecomsol(str(tmp_path / "scripts" / "subs.txt"), xs, ts)
    tpl = open(os.path.join(GOLDEN, "reference_script_template.txt")).read()
    (tmp_path / "cell.txt").write_text(tpl.format(end_t=0.8, max_edge=2, improve=1, topo=1, collide=1, mesh="scripts/disk.txt"))
    (tmp_path / "subs.txt").write_text(tpl.format(end_t=0.8, max_edge=8, improve=1, topo=1, collide=1, mesh="scripts/subs.txt"))
    (tmp_path / "out").mkdir()
    r = subprocess.run([REFAPP, "cell.txt", "subs.txt", "out"], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    run = tmp_path / "out" / "cell_subs_example" / "run0001"
    o = OracleSim("k1", xc, tc, xs, ts); o.srand(1)
    tc32 = np.ascontiguousarray(tc, dtype=np.int32)
    t = 0.0
    for f in range(20):
        cur = 0.0
        while not (cur >= 0.04 - 1e-7):                         # framestepper.h:14-76
            dt = 0.04 - cur if cur + 0.005 > 0.04 else 0.005
            o.step(dt, t); t += dt; cur += dt
        if f in (0, 7, 19):
            x = o.get(0, "x"); cA = o.get(0, "cA"); cC = o.get(0, "cC"); fa = o.get(0, "fa_u"); ea = o.get(0, "ea")
            mine = str(tmp_path / ("mine%d.vtk" % f))
            assert host.cellbp_write_vtk_cell(mine.encode(), C.c_double(t), len(x), d(x), len(tc32), tc32.ctypes.data_as(_ip), d(cA), d(cC), d(fa), d(ea), C.c_double(1000.0)) == 0
            assert open(mine, "rb").read() == open(run / ("cell_frame%04d.vtk" % f), "rb").read(), f


def test_shell_accepts_eltopo_scripts_only_when_told_to(meshes, tmp_path):
    """cellbp_gpu refuses a script that leaves El Topo's switches on unless --no-eltopo is given; then it says what it skips.  Without a
    GPU (here) the accepted run must stop at wcfe_create with the no-CPU-fallback message, not compute anything."""
    exe = os.path.join(HOST, "cellbp_gpu")
    subprocess.check_call(["make", "-s", "-C", HOST])
    xc, tc, xs, ts = meshes
    write_ecomsol(str(tmp_path / "disk.txt"), xc, tc); write_ecomsol(str(tmp_path / "subs.txt"), xs, ts)
    tpl = open(os.path.join(GOLDEN, "reference_script_template.txt")).read()
    (tmp_path / "cell.txt").write_text(tpl.format(end_t=0.04, max_edge=2, improve=1, topo=1, collide=1, mesh=str(tmp_path / "disk.txt")))
    (tmp_path / "subs_script.txt").write_text(tpl.format(end_t=0.04, max_edge=8, improve=1, topo=1, collide=1, mesh=str(tmp_path / "subs.txt")))
    (tmp_path / "out").mkdir()
    args = [exe, str(tmp_path / "cell.txt"), str(tmp_path / "subs_script.txt"), str(tmp_path / "out"), "--quiet"]
    r = subprocess.run(args, capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "does not link El Topo" in r.stderr and "--no-eltopo" in r.stderr
    r = subprocess.run(args + ["--no-eltopo"], capture_output=True, text=True, timeout=300)
    assert "does not link El Topo" not in r.stderr and r.stderr.count("SKIPPED") == 2
    import torch
    if torch.cuda.is_available():
        assert r.returncode == 0 and "8 steps" in r.stdout
    else:
        assert r.returncode != 0 and "no CPU fallback" in r.stderr
