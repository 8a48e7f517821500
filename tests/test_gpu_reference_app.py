"""The literal drop-in, whole application: oracle/_ref/cellBP_gpu_<v> is the UNMODIFIED reference application (cellBP/main.cpp, script
parser, VTK writer, El Topo with collision safety and remeshing ON) linked against wholecellfe_b200/csrc/ul_adapter.cpp + libwcfe.so
INSTEAD of UL_growth_explicit.o; oracle/_ref/cellBP_ref_<v> is the same application with the reference's own UL_growth_explicit.o.
Both are built by oracle/build_app.sh in the container that has the reference sources; the test runs them side by side on the GPU box
on the shipped example meshes and compares the .vtk frames they write (6 significant digits, iomesh.cpp:854).  GPU only."""
import os
import subprocess
import numpy as np
import pytest
from wholecellfe_b200.meshio import write_ecomsol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def have(variant):
    return all(os.path.exists(os.path.join(REFDIR, "cellBP_%s_%s" % (k, variant))) for k in ("ref", "gpu"))


pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have("k1"), reason="oracle/_ref/cellBP_{ref,gpu}_* not built here (oracle/build_app.sh)")]


def sections(path):
    """vtk file -> {"POINTS": array, "concentrations": array, "fa_u": array, "text": [non-numeric lines]}"""
    out = {"text": []}
    key, rows = None, []

    def flush():
        if key is not None:
            out[key] = np.array(rows, dtype=float)
    for line in open(path).read().split("\n"):
        w = line.split()
        if not w:
            continue
        try:
            vals = [float(v) for v in w]
            rows.append(vals)
        except ValueError:
            if w[0] in ("POINTS", "CELLS", "CELL_TYPES", "LOOKUP_TABLE", "VECTORS"):
                flush()
                key = {"LOOKUP_TABLE": "scalars"}.get(w[0], w[0] if w[0] != "VECTORS" else w[1]); rows = []
            if w[0] != "CellBP":                     # the title line carries the time in default precision: compared numerically below
                out["text"].append(line)
    flush()
    return out


def run_pair(tmp_path, variant, end_t, cell_max_edge=2.0):
    meshes = np.load(os.path.join(ROOT, "tests", "golden", "example_meshes.npz"))
    w = tmp_path / variant
    (w / "scripts").mkdir(parents=True)
    write_ecomsol(str(w / "scripts" / "disk.txt"), meshes["cell_xyz"], meshes["cell_tri"])
    write_ecomsol(str(w / "scripts" / "subs.txt"), meshes["subs_xyz"], meshes["subs_tri"])
    tpl = open(os.path.join(ROOT, "tests", "golden", "reference_script_template.txt")).read()
    (w / "cell.txt").write_text(tpl.format(end_t=end_t, max_edge=cell_max_edge, improve=1, topo=1, collide=1, mesh="scripts/disk.txt"))
    (w / "subs.txt").write_text(tpl.format(end_t=end_t, max_edge=8, improve=1, topo=1, collide=1, mesh="scripts/subs.txt"))
    runs = {}
    for kind in ("ref", "gpu"):
        out = w / ("out_" + kind); out.mkdir()
        r = subprocess.run([os.path.join(REFDIR, "cellBP_%s_%s" % (kind, variant)), "cell.txt", "subs.txt", str(out)], cwd=str(w),
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (kind, r.stdout[-2000:], r.stderr[-2000:])
        runs[kind] = out / "cell_subs_example" / "run0001"
    return runs


N_CELL_AT_START = 296      # vertices of the shipped example cell (tests/golden/example_meshes.npz)


def compare_frames(runs, frames, rtol, atol, surfaces=("cell", "subs")):
    npoints = []
    for f in frames:
        ncell = None
        for s in surfaces:
            name = "%s_frame%04d.vtk" % (s, f)
            a, b = sections(str(runs["ref"] / name)), sections(str(runs["gpu"] / name))
            assert a["text"] == b["text"], name                       # headers, counts, section layout
            for k in a:
                if k == "text":
                    continue
                assert a[k].shape == b[k].shape, (name, k)
                x, y = a[k], b[k]
                if s == "subs" and k == "fa_u":
                    # the substrate's SurfTrack is CONSTRUCTED with the cell's fa_u vector (main.cpp:609: as many
                    # entries as the cell had vertices at start-up, 296 for the example) and its writer indexes that copy by
                    # substrate node (iomesh.cpp:831-833 via main.cpp:451): rows from 296 on are out-of-bounds heap reads, whatever
                    # the cell's vertex count has become since (equal between two builds only by accident of heap layout)
                    x, y = x[:N_CELL_AT_START], y[:N_CELL_AT_START]
                assert np.allclose(x, y, rtol=rtol, atol=atol), (name, k, float(np.abs(x - y).max()))
            if s == "cell":
                ncell = len(a["POINTS"])
                npoints.append(ncell)
    return npoints


def test_k1_whole_application_with_eltopo_matches_the_reference(tmp_path):
    """config 1: 64 steps of the example, El Topo's integrate() with collision safety and improve_mesh() every step"""
    runs = run_pair(tmp_path, "k1", 0.32)
    frames = sorted(p.name for p in runs["gpu"].glob("cell_frame*.vtk"))
    assert frames == ["cell_frame%04d.vtk" % i for i in range(8)]
    assert (runs["gpu"] / "aaa-script_cell.txt").exists()
    n = compare_frames(runs, range(8), rtol=2e-5, atol=2e-7)
    assert set(n) == {296}                                            # the example does not remesh under k1


def test_k1_remeshing_in_the_loop(tmp_path):
    """max_edge_length 1.3 makes El Topo split edges from the first step on: the vertex count grows 296 -> 452 -> ... while the
    adapter notices the change, re-uploads connectivity and state, and the frames stay equal to the reference's"""
    runs = run_pair(tmp_path, "k1", 0.32, cell_max_edge=1.3)
    n = compare_frames(runs, range(8), rtol=2e-5, atol=2e-7)
    assert n[0] > 296 and n[-1] > n[0], n


@pytest.mark.skipif(not have("k31"), reason="k31 application not built")
def test_k31_whole_application_first_frames(tmp_path):
    """k31 amplifies rounding differences about tenfold every four steps once bonds rupture (chaos): the first three
    frames (24 steps) agree to the printed precision, later ones cannot, between ANY two builds"""
    runs = run_pair(tmp_path, "k31", 0.16)
    compare_frames(runs, range(3), rtol=1e-4, atol=1e-6)


@pytest.mark.skipif(not have("kMD"), reason="kMD application not built")
def test_kmd_whole_application_runs(tmp_path):
    """kMD's force law is discontinuous at zero extension, so from rest the directions of the first integrin forces are rounding
    noise in both builds; the application-level check is structural: same frames, same mesh sizes, positions close"""
    runs = run_pair(tmp_path, "kMD", 0.16)

    def area(sec):
        x, t = sec["POINTS"], sec["CELLS"][:, 1:].astype(int)
        e1, e2 = x[t[:, 1]] - x[t[:, 0]], x[t[:, 2]] - x[t[:, 0]]
        return 0.5 * np.abs(e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]).sum()
    for f in range(4):
        a = sections(str(runs["ref"] / ("cell_frame%04d.vtk" % f))); b = sections(str(runs["gpu"] / ("cell_frame%04d.vtk" % f)))
        assert a["text"] == b["text"]
        assert np.abs(a["POINTS"] - b["POINTS"]).max() < 0.5          # node spacing is ~1.2 um
        assert abs(area(a) - area(b)) < 2e-3 * area(a), (f, area(a), area(b))
        assert np.abs(a["POINTS"].mean(0) - b["POINTS"].mean(0)).max() < 0.01
