"""File naming shared by the parameter-study writer and reader
(``nsol.parameter_study.ParameterStudy``, nsol/parameter_study.py:23-91):
``<dir>/<name>_parameters.txt``, ``<name>_measure_<m>.txt``,
``<name>_computational_time.txt`` and ``<name>_reconstructions.npz``."""
import os
import time
from abc import ABCMeta

FILENAME_EXTENSION = "txt"                 # nsol/definitions.py:14
REGEX_FILENAMES = "[A-Za-z0-9+-_]+"        # nsol/definitions.py:13


class ParameterStudy(object):
    __metaclass__ = ABCMeta

    def __init__(self, directory, name):
        self._directory = directory
        self._name = name

    def get_parameter_study_name(self):
        return self._name

    def _path(self, specifier, extension=FILENAME_EXTENSION):
        return os.path.join(self._directory, self._name + specifier + "." + extension)

    def _get_path_to_file_parameters(self, specifier="_parameters"):
        return self._path(specifier)

    def _get_path_to_file_measures(self, measure, specifier="_measure_"):
        return self._path(specifier + measure)

    def _get_path_to_file_computational_time(self, specifier="_computational_time"):
        return self._path(specifier)

    def _get_path_to_file_reconstructions(self, specifier="_reconstructions"):
        return self._path(specifier, "npz")


# ---- small file helpers (the reference takes these from pysitk.python_helper) ---------
def write_to_file(path, text, access_mode="w"):
    directory = os.path.dirname(path)
    if directory and not os.path.isdir(directory):
        os.makedirs(directory)
    with open(path, access_mode) as fh:
        fh.write(text)


def write_array_to_file(path, array, fmt="%.10e", delimiter="\t"):
    import numpy as np
    with open(path, "a") as fh:
        np.savetxt(fh, np.atleast_2d(array), fmt=fmt, delimiter=delimiter)


def read_file_line_by_line(path):
    with open(path) as fh:
        return fh.readlines()


def get_time_stamp():
    return time.strftime("%Y-%m-%d %H:%M:%S")


def is_float(text):
    try:
        float(text)
        return True
    except ValueError:
        return False
