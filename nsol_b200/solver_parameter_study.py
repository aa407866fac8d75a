"""Parameter-study driver (``nsol.solver_parameter_study.SolverParameterStudy``,
nsol/solver_parameter_study.py:28-323): for every tuple of ``itertools.product`` over the
parameter lists set the solver parameters, run, evaluate the observer's measures and append
one row per file; the last iterate of every run is stored as float16 in
``*_reconstructions.npz`` keyed by the run index.  ``append=True`` continues a previous
study after checking that the headers match.

B200 additions (the reference runs the points one after another on one CPU thread):
  * a primal-dual study that sweeps only ``alpha`` and whose observer has no measures is
    batched -- one fused launch per iteration advances every alpha (``run_sweep``);
  * when ``torch.distributed`` is initialised the points are dealt round-robin to the ranks
    (one GPU each, no communication on the data path) and rank 0 writes the files.
"""
import datetime
import itertools
import re
import time
from abc import ABCMeta, abstractmethod

import numpy as np

from nsol_b200.parameter_study import (ParameterStudy, get_time_stamp, is_float, npz_members, write_array_to_file,
                                       write_npz_members, write_to_file)
from nsol_b200.reader_parameter_study import ReaderParameterStudy

MAX_SWEEP_BATCH = 32


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist
    except ImportError:
        pass
    return None


class SolverParameterStudy(ParameterStudy):
    __metaclass__ = ABCMeta

    def __init__(self, solver, parameters, observer, dir_output, name, reconstruction_info, append):
        ParameterStudy.__init__(self, directory=dir_output, name=name)
        self._solver = solver
        self._parameters = parameters
        self._observer = observer
        self._reconstruction_info = reconstruction_info
        self._append = append
        self._computational_time = datetime.timedelta(seconds=0)

    def run(self):
        import os
        self._observer.set_name(self._name)
        self._observer.clear_x_list()
        self._solver.set_observer(self._observer)
        dist = _dist()
        rank = dist.get_rank() if dist else 0
        # Rank 0 alone looks at the files, decides between "new study" and "append", creates / checks them;
        # the decision (or its error) is broadcast, so no rank ever reads a file another one is writing.
        state = [None]
        if rank == 0:
            try:
                previous = os.path.isfile(self._get_path_to_file_parameters())
                if not self._append or not previous:
                    self._create_file_parameters()
                    self._create_files_measures()
                    self._create_file_computational_time()
                    state[0] = ("new", None)
                else:
                    self._check_that_studies_match()
                    state[0] = ("append", None)
            except Exception as e:     # every rank must leave run() together
                state[0] = ("error", e)
        if dist:
            dist.broadcast_object_list(state, src=0)
        kind, err = state[0]
        if kind == "error":
            raise err
        self._append = kind == "append"
        t0 = time.time()
        self._run()
        self._computational_time = datetime.timedelta(seconds=time.time() - t0)

    def get_computational_time(self):
        return self._computational_time

    def get_parameters(self):
        return self._parameters

    # ------------------------------------------------------------------ sweep
    def _run(self):
        import os
        keys = list(self._parameters.keys())
        points = list(itertools.product(*self._parameters.values()))
        dist = _dist()
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
        offset, members = 0, []
        if rank == 0:
            if self._append:       # only the writer reads the study being continued
                reader = ReaderParameterStudy(directory=self._directory, name=self._name)
                reader.read_study()
                offset = len(reader.get_parameters_to_line().keys())
                dic_x = dict(reader.get_reconstructions())
            else:
                dic_x = {k: v for k, v in self._reconstruction_info.items()}
            members = npz_members(dic_x)     # reconstruction_info / the runs of the study being appended to
        if dist:
            box = [offset]
            dist.broadcast_object_list(box, src=0)
            offset = box[0]
        threads = max(1, (os.cpu_count() or 1) // world)
        # Rounds of world x MAX_SWEEP_BATCH points: inside a round the points are dealt round-robin to the ranks
        # (no data-path communication); after every round the rows of its points are appended to the text files
        # in index order, as the reference appends them point by point (nsol/solver_parameter_study.py:194-205),
        # so an interrupted study keeps what it finished.  The reconstructions archive is written once at the
        # end -- also when a later point fails.
        per_round = world * MAX_SWEEP_BATCH
        failure = None
        try:
            for r0 in range(0, len(points), per_round):
                idx = list(range(r0, min(len(points), r0 + per_round)))
                mine = idx[rank::world]
                try:
                    results = self._run_points(keys, points, mine)    # {index: (params, measures, time, x_last)}
                    # the float16 reconstruction of every run is deflated on the rank that computed it
                    packed = npz_members({str(i + offset): np.array(results[i][3], dtype=np.float16) for i in sorted(results)}, threads)
                    for i, member in zip(sorted(results), packed):
                        results[i] = results[i][:3] + (member,)
                except Exception as e:
                    results, failure = {}, e
                if dist:
                    gathered = [None] * world if rank == 0 else None
                    dist.gather_object((results, failure), gathered, dst=0)
                    flag = [None]
                    if rank == 0:
                        results = {}
                        for part, err in gathered:
                            results.update(part)
                            failure = failure or err
                        flag[0] = failure
                    dist.broadcast_object_list(flag, src=0)
                    failure = flag[0]
                if rank == 0:
                    for i in idx:
                        if i not in results:
                            break              # rows stay contiguous: stop at the first missing point
                        params, measures, ctime, member = results[i]
                        for measure, values in measures.items():
                            self._add_to_file_measures(measure, np.asarray(values).reshape(1, -1))
                        self._add_to_file_computational_time(ctime)
                        self._add_to_file_parameters(params)
                        members.append(member)
                if failure is not None:
                    break
        finally:
            if rank == 0:
                self._write_to_file_reconstructions(members)
        if failure is not None:
            raise failure

    def _apply_point(self, keys, vals):
        params = {}
        for j, key in enumerate(keys):
            getattr(self._solver, "set_%s" % key)(vals[j])
            params[key] = str(getattr(self._solver, "get_%s" % key)())
        return params

    def _can_batch(self, keys):
        has_measures = len(self._observer.get_measures()) > 0
        return (keys == ["alpha"] and not has_measures and hasattr(self._solver, "run_sweep")
                and getattr(self._solver, "_probe", None) is not None and self._solver._probe()["kind"] == "denoise")

    def _run_points(self, keys, points, mine):
        results = {}
        if mine and self._can_batch(keys):
            for c in range(0, len(mine), MAX_SWEEP_BATCH):
                chunk = mine[c:c + MAX_SWEEP_BATCH]
                t0 = time.time()
                xs = self._solver.run_sweep([points[i][0] for i in chunk])
                per_point = datetime.timedelta(seconds=(time.time() - t0) / len(chunk))
                for row, i in enumerate(chunk):
                    params = self._apply_point(keys, points[i])
                    results[i] = (params, {}, per_point, xs[row])
            self._solver.set_x0(self._solver.get_x0())
            return results
        for i in mine:
            params = self._apply_point(keys, points[i])
            self._solver.run()
            self._observer.compute_measures()
            measures = {m: np.array(v) for m, v in self._observer.get_measures().items()}
            results[i] = (params, measures, self._observer.get_computational_time(),
                          np.array(self._observer.get_x_list()[-1]))
            self._observer.clear_x_list()
            self._solver.set_x0(self._solver.get_x0())      # nsol/solver_parameter_study.py:221
        return results

    # ------------------------------------------------------------------ append check
    def _check_that_studies_match(self):
        reader = ReaderParameterStudy(directory=self._directory, name=self._name)
        reader.read_study()
        new = self._get_fileheader().split(" ")[1:-2]
        old = reader.get_file_header().split(" ")[1:-2]

        def fail(h1, h2, info=""):
            raise RuntimeError("Study cannot be appended as parameter settings do not match: %s != %s%s"
                               % (h1, h2, info))
        if len(new) != len(old):
            fail(new, old)
        for a, b in zip(new, old):
            a, b = re.sub(",", "", a), re.sub(",", "", b)
            if a == b:
                continue
            if "=" in a and "=" in b:
                (ka, va), (kb, vb) = a.split("="), b.split("=")
                if ka != kb:
                    fail(a, b)
                if is_float(va) and is_float(vb) and abs(float(va) - float(vb)) < 1.5e-6:
                    continue
            fail(a, b)

    # ------------------------------------------------------------------ files
    def _create_file_parameters(self):
        header = self._get_fileheader() + "## " + "\t".join(self._parameters.keys()) + "\n"
        write_to_file(self._get_path_to_file_parameters(), header, "w")

    def _create_files_measures(self):
        for measure in self._observer.get_measures().keys():
            header = self._get_fileheader() + "## " + measure + " for iteration 0 to n\n"
            write_to_file(self._get_path_to_file_measures(measure), header, "w")

    def _create_file_computational_time(self):
        header = self._get_fileheader() + "## Computational time measured for n iterations\n"
        write_to_file(self._get_path_to_file_computational_time(), header, "w")

    def _add_to_file_parameters(self, dic_parameters):
        write_to_file(self._get_path_to_file_parameters(), "\t".join(dic_parameters.values()) + "\n", "a")

    def _add_to_file_measures(self, measure, nda):
        write_array_to_file(self._get_path_to_file_measures(measure), nda)

    def _add_to_file_computational_time(self, computational_time):
        write_to_file(self._get_path_to_file_computational_time(), str(computational_time) + "\n", "a")

    def _write_to_file_reconstructions(self, members):
        """Same file as the reference's np.savez_compressed(path, **dic) (nsol/solver_parameter_study.py:320-321),
        assembled from members that were deflated in parallel."""
        path = self._get_path_to_file_reconstructions()
        if not write_npz_members(path, members):
            # archive needs zip64: let numpy write it (single-threaded)
            import io
            import zlib
            dic = {}
            for name, _, _, comp in members:
                dic[name[:-4]] = np.lib.format.read_array(io.BytesIO(zlib.decompress(comp, -15)), allow_pickle=False)
            np.savez_compressed(path, **dic)

    def _header_from_keys(self, keys):
        header = "## " + self._name
        for key in keys:
            if key not in self._parameters.keys():
                header += ", %s=%s" % (key, str(getattr(self._solver, "get_" + key)()))
        return header + " (%s)\n" % (get_time_stamp())

    @abstractmethod
    def _get_fileheader(self):
        pass
