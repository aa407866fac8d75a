"""Parameter-study driver: file formats, append check, reader round trip (CPU, with a
stub solver -- the sweep logic itself needs no GPU) and, on the GPU, the batched alpha sweep
against per-point runs."""
import datetime
import os

import numpy as np
import pytest

from nsol_b200.observer import Observer
from nsol_b200.reader_parameter_study import ReaderParameterStudy
from nsol_b200.solver_parameter_study import SolverParameterStudy
import nsol_b200.primal_dual_solver as pd
from nsol_b200.primal_dual_solver_parameter_study import PrimalDualSolverParameterStudy


class _StubSolver(object):
    """Records the set_<key>/run/set_x0 protocol of nsol/solver_parameter_study.py:175-221."""

    def __init__(self, n=12):
        self._alpha, self._rho, self._x0 = 0.01, 0.5, np.arange(float(n))
        self._observer = None
        self.calls = []

    def set_alpha(self, v):
        self._alpha = v

    def get_alpha(self):
        return self._alpha

    def set_rho(self, v):
        self._rho = v

    def get_rho(self):
        return self._rho

    def get_iterations(self):
        return 3

    def set_observer(self, o):
        self._observer = o

    def get_x0(self):
        return self._x0

    def set_x0(self, x0):
        self.calls.append("set_x0")

    def run(self):
        self.calls.append(("run", self._alpha, self._rho))
        for i in range(4):
            self._observer.add_x(self._x0 * self._alpha + i * self._rho)
        self._observer.set_computational_time(datetime.timedelta(seconds=1.5))


class _StubStudy(SolverParameterStudy):
    def _get_fileheader(self):
        return self._header_from_keys(["alpha", "rho", "iterations"])


def _make(tmp_path, append=False, alphas=(0.1, 0.2), rhos=(1.0, 2.0, 3.0)):
    solver = _StubSolver()
    obs = Observer()
    obs.set_measures({"SUM": lambda x: float(np.sum(x)), "MAX": lambda x: float(np.max(x))})
    study = _StubStudy(solver=solver, parameters={"alpha": list(alphas), "rho": list(rhos)}, observer=obs,
                       dir_output=str(tmp_path), name="Stub", reconstruction_info={"shape": (3, 4)}, append=append)
    return solver, study


def test_study_files_and_reader_roundtrip(tmp_path):
    solver, study = _make(tmp_path)
    study.run()
    assert [c for c in solver.calls if c != "set_x0"] == [("run", a, r) for a in (0.1, 0.2) for r in (1.0, 2.0, 3.0)]
    assert solver.calls.count("set_x0") == 6
    lines = open(os.path.join(str(tmp_path), "Stub_parameters.txt")).read().split("\n")
    assert lines[0].startswith("## Stub, iterations=3 (") and lines[1] == "## alpha\trho"
    assert lines[2] == "0.1\t1.0" and lines[7] == "0.2\t3.0"
    reader = ReaderParameterStudy(str(tmp_path), "Stub")
    reader.read_study()
    assert sorted(reader.get_measures()) == ["MAX", "SUM"]
    assert reader.get_parameters() == {"alpha": [0.1, 0.2], "rho": [1.0, 2.0, 3.0]}
    res = reader.get_results("SUM")
    assert res.shape == (6, 4)          # one row per run, one column per stored iterate
    assert abs(res[0, 0] - np.sum(np.arange(12.0) * 0.1)) < 1e-8
    rec = reader.get_reconstructions()
    assert tuple(rec["shape"]) == (3, 4) and rec["5"].dtype == np.float16 and len(rec.files) == 7
    assert list(reader.get_lines_to_parameters({"alpha": 0.2, "rho": [1.0, 2.0, 3.0]})) == [3, 4, 5]
    assert reader.get_line_to_parameter_labels()[4] == "alpha=0.2, rho=2.0"
    times = open(os.path.join(str(tmp_path), "Stub_computational_time.txt")).read().split("\n")
    assert times[2] == "0:00:01.500000"


def test_study_append_continues_and_checks_header(tmp_path):
    _, study = _make(tmp_path)
    study.run()
    _, study2 = _make(tmp_path, append=True, alphas=(0.3,), rhos=(1.0,))
    study2.run()
    reader = ReaderParameterStudy(str(tmp_path), "Stub")
    reader.read_study()
    assert len(reader.get_parameters_to_line()) == 7 and "6" in reader.get_reconstructions().files

    class Other(_StubStudy):
        def _get_fileheader(self):
            return self._header_from_keys(["rho", "iterations"])
    solver = _StubSolver()
    obs = Observer()
    obs.set_measures({"SUM": lambda x: 0.0})
    bad = Other(solver=solver, parameters={"alpha": [0.5]}, observer=obs, dir_output=str(tmp_path), name="Stub",
                reconstruction_info={}, append=True)
    with pytest.raises(RuntimeError):
        bad.run()


def test_study_type_check():
    with pytest.raises(TypeError):     # nsol/primal_dual_solver_parameter_study.py:48-49
        PrimalDualSolverParameterStudy(solver=_StubSolver(), observer=Observer(), dir_output="/tmp")


@pytest.mark.gpu
def test_primal_dual_study_batched_sweep_matches_sequential(tmp_path, golden):
    """BASELINE config 5 in miniature: alpha sweep; the batched fast path (observer without
    measures) stores the same float16 reconstructions as point-by-point runs with measures."""
    from test_gpu_parity import make_pd
    obs = golden("pd", "in/bw2d")
    alphas = np.linspace(0.01, 0.05, 5)
    out = {}
    for tag, measures in (("batched", {}), ("sequential", {"SUM": lambda x: float(np.sum(x))})):
        solver = make_pd(obs, reg="TV", data="L2", alpha=0.01, L2=8, iterations=25)
        o = Observer()
        o.set_measures(measures)
        d = tmp_path / tag
        d.mkdir()
        PrimalDualSolverParameterStudy(solver, o, str(d), name="PD", parameters={"alpha": alphas},
                                       reconstruction_info={"shape": obs.shape}).run()
        out[tag] = np.load(str(d / "PD_reconstructions.npz"))
    for i in range(5):
        assert np.array_equal(out["batched"][str(i)], out["sequential"][str(i)])
    from oracle import nsol_oracle as orc
    ref = orc.primal_dual_denoise(obs.reshape(-1), obs.shape, reg="TV", data="L2", alpha=alphas[3], L2=8, iterations=25,
                                  x_scale=float(obs.max()))
    assert np.array_equal(out["batched"]["3"], ref.astype(np.float16))


def test_parallel_npz_writer_is_read_like_savez_compressed(tmp_path):
    """The reconstructions file (nsol/solver_parameter_study.py:320-321: np.savez_compressed) assembled from
    members deflated in parallel: np.load / zipfile see the same arrays as in numpy's own file."""
    import zipfile
    from nsol_b200.parameter_study import npz_members, write_npz_members
    rng = np.random.RandomState(0)
    dic = {"shape": (33, 20), "0": (rng.rand(660) * 255).astype(np.float16), "1": rng.rand(3, 4), "empty": np.zeros(0)}
    mine, theirs = str(tmp_path / "mine.npz"), str(tmp_path / "numpy.npz")
    assert write_npz_members(mine, npz_members(dic, threads=3))
    np.savez_compressed(theirs, **dic)
    a, b = np.load(mine), np.load(theirs)
    assert a.files == b.files == list(dic.keys())
    for k in dic:
        assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k])
    assert zipfile.ZipFile(mine).testzip() is None
    assert all(i.compress_type == zipfile.ZIP_DEFLATED for i in zipfile.ZipFile(mine).infolist())
