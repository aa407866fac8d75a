"""Reader of a parameter study written by ``SolverParameterStudy``
(``nsol.reader_parameter_study.ReaderParameterStudy``, nsol/reader_parameter_study.py:21-347).
Same file formats: first line of ``*_parameters.txt`` is the header, second the
tab-separated parameter names, then one tab-separated row per run."""
import os
import re

import numpy as np

from nsol_b200.parameter_study import (FILENAME_EXTENSION, REGEX_FILENAMES, ParameterStudy,
                                       read_file_line_by_line)


def _natural_key(text):
    return [int(t) if t.isdigit() else t.lower() for t in re.split(r"(\d+)", text)]


class ReaderParameterStudy(ParameterStudy):

    def __init__(self, directory, name):
        ParameterStudy.__init__(self, directory=directory, name=name)
        self._lines_params = None

    def read_study(self):
        if not os.path.isdir(self._directory):
            raise ValueError("Directory '%s' does not exist" % (self._directory))
        self._measures = self._get_measure_names()
        if len(self._measures) == 0:
            raise RuntimeError("No measures to study '%s' found in '%s'" % (self._name, self._directory))
        lines = read_file_line_by_line(self._get_path_to_file_parameters())[1:]
        self._lines_params = [re.sub("\n", "", re.sub("## ", "", line)) for line in lines]
        self._parameters_dic = self._get_parameters()
        for k in self._parameters_dic.keys():
            if len(self._parameters_dic[k]) == 0:
                raise RuntimeError("Directory '%s' does not contain suitable parameter study info" % self._directory)

    def get_reconstructions(self):
        path = self._get_path_to_file_reconstructions()
        if not os.path.isfile(path):
            raise IOError("File '%s' not available" % path)
        return np.load(path)

    def get_measures(self):
        self._check_that_study_was_read()
        return self._measures

    def get_file_header(self):
        self._check_that_study_was_read()
        return read_file_line_by_line(self._get_path_to_file_parameters())[0]

    def get_results(self, measure):
        return np.loadtxt(self._get_path_to_file_measures(measure), skiprows=2)

    def get_parameters(self):
        self._check_that_study_was_read()
        return self._parameters_dic

    def get_parameters_to_line(self):
        self._check_that_study_was_read()
        return {tuple(line.split("\t")): i for i, line in enumerate(self._lines_params[1:])}

    def get_lines_to_parameters(self, parameters):
        self._check_that_study_was_read()
        if parameters.keys() != self._parameters_dic.keys():
            raise ValueError("Provided dictionary keys must match. Required keys for this study are "
                             + str(self._parameters_dic.keys()))
        varying, rows = None, None
        for key, val in parameters.items():
            if type(val) in (tuple, list, np.ndarray):
                if len(val) == 1:
                    raise ValueError("Single entry in key '%s' must not be a list" % (key))
                if len(val) > 1:
                    if varying is not None:
                        raise ValueError("Provided dictionary can only vary in a single key")
                    varying, rows = key, len(val)
        to_line = self.get_parameters_to_line()
        lines = np.zeros(rows, dtype=int)
        for i in range(rows):
            key = tuple(str(parameters[k][i]) if k == varying else str(parameters[k]) for k in parameters.keys())
            lines[i] = to_line[key]
        return lines

    def get_line_to_parameter_labels(self, separator=", ", compact=False):
        labels = {}
        for i, line in enumerate(self._lines_params[1:]):
            vals = line.split("\t")
            if compact:
                labels[i] = separator.join(vals)
            else:
                labels[i] = separator.join(k + "=" + v for k, v in zip(self._parameters_dic.keys(), vals))
        return labels

    def _get_measure_names(self):
        pattern = re.compile(self._name + "_measure_(" + REGEX_FILENAMES + ")[.]" + FILENAME_EXTENSION)
        return [pattern.match(f).group(1) for f in os.listdir(self._directory) if pattern.match(f)]

    def _get_parameters(self):
        names = self._lines_params[0].split("\t")
        rows = self._lines_params[1:]
        out = {}
        for i, name in enumerate(names):
            vals = sorted(set(row.split("\t")[i] for row in rows), key=_natural_key)
            try:
                vals = sorted(float(v) for v in vals)
            except ValueError:
                pass
            out[name] = vals
        return out

    def _check_that_study_was_read(self):
        if self._lines_params is None:
            raise UnboundLocalError("Execute 'read_study' first to get information on parameters.")
