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


# ---- compressed .npz written in parallel ------------------------------------------------------
# The reference stores the reconstructions with np.savez_compressed (nsol/solver_parameter_study.py:
# 320-321): one deflate stream per array, single-threaded -- for BASELINE config 5 (64 x 1024^2 float16
# per study) that is ~5 s per study against ~0.2 s of GPU time.  The same file (a zip of deflated .npy
# members, read back by np.load) is produced here with the members compressed concurrently: zlib
# releases the GIL, and in a multi-rank study every rank compresses the runs it computed.
def npz_member(name, array, level=6):
    """(member name, crc32, uncompressed size, raw-deflate bytes) of one array of an .npz file."""
    import io
    import zlib
    import numpy as np
    bio = io.BytesIO()
    np.lib.format.write_array(bio, np.asanyarray(array), allow_pickle=False)
    raw = bio.getvalue()
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = co.compress(raw) + co.flush()
    return (name + ".npy", zlib.crc32(raw) & 0xFFFFFFFF, len(raw), comp)


def npz_members(dic, threads=None):
    """npz_member for every item of ``dic`` (insertion order kept), compressed by a thread pool."""
    from concurrent.futures import ThreadPoolExecutor
    items = list(dic.items())
    if threads is None:
        threads = min(len(items), max(1, (os.cpu_count() or 1)))
    if threads <= 1 or len(items) <= 1:
        return [npz_member(k, v) for k, v in items]
    with ThreadPoolExecutor(max_workers=threads) as pool:
        return list(pool.map(lambda kv: npz_member(kv[0], kv[1]), items))


def write_npz_members(path, members):
    """Assemble a zip archive (PKZIP 2.0, deflate) from pre-compressed members.  Returns False -- nothing
    written -- if the archive would need zip64 (the caller then falls back to np.savez_compressed)."""
    import struct
    total = sum(len(m[3]) + 30 + 46 + 2 * len(m[0]) for m in members) + 22
    if len(members) >= 0xFFFF or total >= 0xFFFFFFFF or any(m[2] >= 0xFFFFFFFF for m in members):
        return False
    t = time.localtime()
    dostime = (t.tm_hour << 11) | (t.tm_min << 5) | (t.tm_sec // 2)
    dosdate = ((max(t.tm_year, 1980) - 1980) << 9) | (t.tm_mon << 5) | t.tm_mday
    directory = os.path.dirname(path)
    if directory and not os.path.isdir(directory):
        os.makedirs(directory)
    central = []
    with open(path, "wb") as fh:
        for name, crc, usize, comp in members:
            fn = name.encode("utf-8")
            offset = fh.tell()
            fh.write(struct.pack("<IHHHHHIIIHH", 0x04034B50, 20, 0x800, 8, dostime, dosdate, crc, len(comp), usize, len(fn), 0))
            fh.write(fn)
            fh.write(comp)
            central.append(struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 20, 20, 0x800, 8, dostime, dosdate, crc, len(comp), usize,
                                       len(fn), 0, 0, 0, 0, 0x01800000, offset) + fn)
        start = fh.tell()
        for rec in central:
            fh.write(rec)
        size = fh.tell() - start
        fh.write(struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, len(members), len(members), size, start, 0))
    return True
