"""Host-side file formats (dot_b200/io.py): the reference's script format, .msh dialect, restart files, label.obj, and the
fallback partitioner.  Pure CPU."""
import glob
import os

import numpy as np
import pytest

from dot_b200 import io, meshgen

REF = "/root/reference"


def test_script_parser_keys(tmp_path):
    p = tmp_path / "s.txt"
    p.write_text("energy SNH\ntimeIntegration BE\ntimeStepper DOT 16\ninexactSolve 0\nwarmStart 2\nresolution 1000\nsize 2.5\n"
                 "time 10 0.0416667\ndensity 1200\nstiffness 2e5 0.45\nturnOffGravity\nscript twistnsns_old\nhandleRatio 0.02\n"
                 "rotateModel 0 1 0 90\ntuning 2\n1 2\nshape input input/tetMeshes/x.msh\ntol 3\n1e-3 1e-4\n1e-5\nview orthographic\nzoom 0.8\n")
    s = io.parse_script(str(p))
    assert (s.energy, s.time_stepper, s.partitions) == ("SNH", "DOT", 16)
    assert (s.size, s.duration, s.dt, s.rho, s.YM, s.PR) == (2.5, 10.0, 0.0416667, 1200.0, 2e5, 0.45)
    assert not s.with_gravity and s.script == "twistnsns_old" and s.handle_ratio == 0.02
    assert s.rot_axis == (0.0, 1.0, 0.0) and s.rot_deg == 90.0
    assert s.input_shape_path == "input/tetMeshes/x.msh"
    assert s.tol == [1e-3, 1e-4, 1e-5] and s.rel_tol(0) == 1e-3 and s.rel_tol(7) == 1e-5
    assert s.num_frames() == 240 and s.unknown == []
    # k < 2 is rewritten to 4 (Config.cpp:76-80); negative k means block-size mode
    p.write_text("timeStepper DOT 1\n")
    assert io.parse_script(str(p)).partitions == 4
    p.write_text("timeStepper DOT -1 900\n")
    assert io.parse_script(str(p)).block_size == 900
    p.write_text("timeStepper Newton\nenergy FCR\n")
    assert io.parse_script(str(p)).time_stepper == "Newton" and io.parse_script(str(p)).rel_tol(3) == 1e-5
    p.write_text("timeStepper ADMM 10\n")
    with pytest.raises(ValueError):
        io.parse_script(str(p))


def test_msh_and_status_round_trip(tmp_path):
    V, T = meshgen.preset("bar_tiny")
    meshgen.write_msh(str(tmp_path / "m.msh"), V, T)
    V2, T2, SF = io.read_msh(str(tmp_path / "m.msh"))
    assert np.array_equal(V, V2) and np.array_equal(T, T2)          # %.17g round-trips doubles exactly
    assert np.array_equal(SF, meshgen.surface_tris(T))
    s2t = io.surface_to_tet(T, SF)
    assert all(set(SF[i]).issubset(set(T[s2t[i]])) for i in range(SF.shape[0]))
    io.write_label_obj(str(tmp_path / "label.obj"), SF, s2t, np.arange(T.shape[0]) % 3)
    lines = open(tmp_path / "label.obj").read().splitlines()
    assert len(lines) == SF.shape[0] and lines[0].split()[0] == "v"
    rng = np.random.default_rng(0)
    x, v = rng.standard_normal(V.shape), rng.standard_normal(3 * V.shape[0])
    io.write_status(str(tmp_path / "status7"), 7, x, v)
    st = io.read_status(str(tmp_path / "status7"))
    assert st["timestep"] == 7
    assert np.allclose(st["position"], x, rtol=1e-6) and np.allclose(st["velocity"], v, rtol=1e-6)   # %le keeps 7 digits, like the reference
    assert st["dx_Elastic"].shape == x.shape
    w = io.IterStatsWriter(str(tmp_path / "iterStats.txt"), "Newton")
    w.frame(0, np.array([[0.0, 1.5, 2.5], [1.0, 1.25, 1e-9]]))
    w.close()
    assert open(tmp_path / "iterStats.txt").read().splitlines() == ["0 0 1.5 2.5 0", "0 1 1.25 1e-09 0"]
    w = io.IterStatsWriter(str(tmp_path / "iterStats_dot.txt"))
    w.frame(0, np.array([[0.0, 1.5, 2.5], [1.0, 1.25, 1e-9]]))
    w.close()
    assert open(tmp_path / "iterStats_dot.txt").read().splitlines() == ["0 0 1.5 2.5", "0 1 1.25 1e-09"]


def test_rotate_model_matches_axis_angle():
    V = np.array([[1.0, 0.0, 0.0], [0.0, 2.0, 0.0]])
    R = io.rotate_model(V, (0.0, 0.0, 1.0), 90.0)
    assert np.allclose(R, [[0.0, 1.0, 0.0], [-2.0, 0.0, 0.0]], atol=1e-15)
    assert io.rotate_model(V, (0, 0, 1), 0.0) is V


def test_rcb_partition_is_balanced_and_complete():
    V, T = meshgen.preset("bar2K")
    for k in (1, 3, 8, 13):
        ep = io.partition_rcb(V, T, k)
        cnt = np.bincount(ep, minlength=k)
        assert ep.min() == 0 and ep.max() == k - 1 and cnt.min() > 0
        assert cnt.max() - cnt.min() <= max(2, 0.02 * T.shape[0])


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "input")), reason="reference inputs only exist in the build container")
def test_reference_scripts_and_meshes_parse():
    ok = 0
    for f in sorted(glob.glob(os.path.join(REF, "input", "*.txt")) + glob.glob(os.path.join(REF, "input", "tb*", "*.txt"))):
        try:
            s = io.parse_script(f)
        except ValueError:
            continue                                   # steppers / energies outside the GPU path
        assert s.unknown == [], (f, s.unknown)
        assert s.dt > 0 and s.YM > 0
        ok += 1
    assert ok >= 40                                 # every shipped DOT script except the rubberBandPull ones
    V, T, SF = io.read_msh(os.path.join(REF, "input", "tetMeshes", "bunny5K.msh"))
    assert (V.shape[0], T.shape[0]) == (4670, 19379) and SF.shape[0] > 0       # bunny5K sizes
    Dm = np.stack([V[T[:, 1]] - V[T[:, 0]], V[T[:, 2]] - V[T[:, 0]], V[T[:, 3]] - V[T[:, 0]]], axis=2)
    assert (np.linalg.det(Dm) > 0).all()


def _io_fixture():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io_small.npz"))


def test_status_file_is_byte_compatible_with_the_reference(tmp_path):
    """`status<n>` as Optimizer::saveStatus writes it (Optimizer.cpp:1096-1130; fixture written by the unmodified
    reference, oracle/gen_golden.py --io).  Reading it gives the reference's state; writing that state back gives the same BYTES."""
    z = _io_fixture()
    ref_txt = str(z["status_txt"])
    p = tmp_path / "status_ref"
    p.write_text(ref_txt)
    st = io.read_status(str(p))
    assert st["timestep"] == int(z["status_timestep"])
    # %le keeps 7 significant digits: the parsed fields are the reference's binary state to that precision
    assert np.allclose(st["position"], z["V"], rtol=1e-6, atol=1e-12)
    assert np.allclose(st["velocity"], z["velocity"].reshape(-1), rtol=1e-6, atol=1e-12)
    assert st["dx_Elastic"].shape == z["V"].shape
    q = tmp_path / "status_mine"
    io.write_status(str(q), st["timestep"], z["V"], z["velocity"], st["dx_Elastic"])   # positions / velocity from the BINARY state
    assert q.read_text() == ref_txt


def test_iterstats_rows_match_the_reference_format(tmp_path):
    """iterStats.txt of `timeStepper DOT` has 4 columns, of `timeStepper Newton` a trailing ` 0`: re-writing the parsed
    rows of the reference's own files reproduces them byte for byte (default ostream formatting == %g)."""
    z = _io_fixture()
    for kind, key in (("DOT", "dot/iterStats_txt"), ("Newton", "newton/iterStats_txt")):
        ref_txt = str(z[key])
        rows = np.array([[float(v) for v in line.split()] for line in ref_txt.strip().splitlines()])
        assert rows.shape[1] == (4 if kind == "DOT" else 5)
        p = tmp_path / ("iter_%s.txt" % kind)
        w = io.IterStatsWriter(str(p), kind)
        for f in np.unique(rows[:, 0]).astype(int):
            w.frame(int(f), rows[rows[:, 0] == f][:, 1:4])
        w.close()
        assert p.read_text() == ref_txt


def test_msh_writer_matches_saveTetMesh(tmp_path):
    """.msh written by IglUtils::saveTetMesh (IglUtils.cpp:627-679) for the reference's final configuration vs
    io.write_msh_reference for the same arrays: identical bytes; and the reader recovers the arrays from the reference's file."""
    z = _io_fixture()
    ref_txt = str(z["msh_txt"])
    p = tmp_path / "ref.msh"
    p.write_text(ref_txt)
    V, T, SF = io.read_msh(str(p))
    assert np.array_equal(T, z["T"]) and np.array_equal(SF, z["SF"])
    assert np.allclose(V, z["V"], rtol=1e-6, atol=1e-12)
    q = tmp_path / "mine.msh"
    io.write_msh_reference(str(q), z["V"], z["T"], z["SF"])
    assert q.read_text() == ref_txt


def test_script_defaults_are_the_reference_defaults(tmp_path):
    p = tmp_path / "s.txt"
    p.write_text("energy FCR\n")
    s = io.parse_script(str(p))
    assert (s.rho, s.YM, s.PR, s.dt, s.duration, s.warm_start, s.handle_ratio) == (1.0, 100.0, 0.4, 0.025, 10.0, 2, 0.01)   # Config.cpp:33-37


def test_info_txt_has_the_layout_of_saveInfoForPresent(tmp_path):
    """info.txt (main.cpp:338-358): '<nV> <nT>', '<iterNum> <innerIterAmt> 0 0 0', Timer::print of the descent timer, of the 14
    timer_step activities (main.cpp:867-880, same names, same order) and of the 7 ADMM activities, then '0 0'.  Timer::print
    (Utils/Timer.hpp:58-69): '<n> activities:' + per activity a width-10 right-aligned default-formatted double, ' s: <name>', + Total."""
    p = tmp_path / "info.txt"
    io.write_info_txt(str(p), 17315, 86058, 200, 3613, 1.4321, {"numericalFactorization": 0.516, "backSolve": 0.495, "lineSearch_eVal": 0.367})
    L = p.read_text().splitlines()
    assert L[0] == "17315 86058" and L[1] == "200 3613 0 0 0"
    assert L[2] == "1 activities:" and L[3] == "    1.4321 s: descent" and L[4] == "    1.4321 s: Total"
    assert L[5] == "14 activities:"
    names = [l.split(" s: ")[1] for l in L[6:20]]
    assert names == ["matrixComputation", "matrixAssembly", "symbolicFactorization", "numericalFactorization", "backSolve", "lineSearch_other",
                     "modifyGrad", "modifySearchDir", "updateHistory", "lineSearch_eVal", "fullyImplicit_eComp", "solve_extraComp", "compGrad", "CCD"]
    assert L[9] == "     0.516 s: numericalFactorization" and L[6] == "         0 s: matrixComputation"
    assert L[20] == "     1.378 s: Total"
    assert L[21] == "7 activities:" and L[29] == "         0 s: Total" and L[30] == "0 0" and len(L) == 31
    with pytest.raises(ValueError):
        io.write_info_txt(str(p), 1, 1, 1, 1, 0.0, {"notATimer": 1.0})
