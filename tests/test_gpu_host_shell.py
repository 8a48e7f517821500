"""cellbp_gpu end to end on the GPU: reference CLI, script format, frame loop and .vtk output.  GPU only."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from oracle.oracle import OracleSim
from wholecellfe_b200.meshio import write_ecomsol

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "wholecellfe_b200", "csrc", "host")
SCRIPT = """relative_output_path cell_subs_example
frame_dt 0.04
sim_dt 0.005
end_sim_t %g
surftrack_parameters
(
   perform_improvement %d
   allow_topology_changes 0
   collision_safety 0
)
ecomsolfile
(
   filepath %s
)
"""


def numbers(path):
    out = []
    for line in open(path).read().split("\n"):
        try:
            out.append([float(v) for v in line.split()])
        except ValueError:
            out.append(line)
    return out


def test_cellbp_gpu_runs_the_example_and_writes_reference_frames(meshes, tmp_path):
    exe = os.path.join(HOST, "cellbp_gpu")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", HOST])
    xc, tc, xs, ts = meshes
    write_ecomsol(str(tmp_path / "disk.txt"), xc, tc); write_ecomsol(str(tmp_path / "subs.txt"), xs, ts)
    (tmp_path / "cell.txt").write_text(SCRIPT % (0.16, 0, tmp_path / "disk.txt"))
    (tmp_path / "subs_script.txt").write_text(SCRIPT % (0.16, 0, tmp_path / "subs.txt"))
    out = tmp_path / "out"; out.mkdir()
    r = subprocess.run([exe, str(tmp_path / "cell.txt"), str(tmp_path / "subs_script.txt"), str(out), "--variant", "k1", "--quiet"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    run = out / "cell_subs_example" / "run0001"
    assert (run / "aaa-script_cell.txt").exists() and (run / "aaa-script_subs.txt").exists()
    frames = sorted(p.name for p in run.glob("cell_frame*.vtk"))
    assert frames == ["cell_frame%04d.vtk" % i for i in range(4)], frames          # 0.16 / 0.04 frames, last one is frame 3 (quirk q8)
    assert "32 steps" in r.stdout
    # the same 32 steps on the oracle, written by the same writer
    o = OracleSim("k1", xc, tc, xs, ts); o.srand(1)
    t = 0.0
    for f in range(4):
        cur = 0.0
        while not (cur >= 0.04 - 1e-7):
            dt = 0.04 - cur if cur + 0.005 > 0.04 else 0.005
            o.step(dt, t); t += dt; cur += dt
    L = C.CDLL(os.path.join(HOST, "libcellbp_host.so"))
    dp = C.POINTER(C.c_double); ip = C.POINTER(C.c_int)
    d = lambda a: a.ctypes.data_as(dp)
    x = o.get(0, "x"); cA = o.get(0, "cA"); cC = o.get(0, "cC"); fa = o.get(0, "fa_u"); ea = o.get(0, "ea")
    tc32 = np.ascontiguousarray(tc, dtype=np.int32)
    ref = str(tmp_path / "ref.vtk")
    assert L.cellbp_write_vtk_cell(ref.encode(), C.c_double(t), len(x), d(x), len(tc32), tc32.ctypes.data_as(ip), d(cA), d(cC), d(fa), d(ea), C.c_double(1000.0)) == 0
    a = numbers(str(run / "cell_frame0003.vtk")); b = numbers(ref)
    assert len(a) == len(b)
    for la, lb in zip(a, b):
        if isinstance(la, str) or isinstance(lb, str):
            assert la == lb
        else:
            assert np.allclose(la, lb, rtol=2e-5, atol=1e-9), (la, lb)
    # second invocation lands in run0002; El Topo flags left on are refused
    r2 = subprocess.run([exe, str(tmp_path / "cell.txt"), str(tmp_path / "subs_script.txt"), str(out), "--quiet"], capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0 and (out / "cell_subs_example" / "run0002").exists()
    (tmp_path / "cell_on.txt").write_text(SCRIPT % (0.16, 1, tmp_path / "disk.txt"))
    r3 = subprocess.run([exe, str(tmp_path / "cell_on.txt"), str(tmp_path / "subs_script.txt"), str(out)], capture_output=True, text=True, timeout=60)
    assert r3.returncode != 0 and "El Topo" in r3.stderr
