"""pysitk.python_helper stand-in: timing + silent printing (test infrastructure)."""
import datetime
import time


def start_timing():
    return time.time()


def stop_timing(t0):
    return datetime.timedelta(seconds=time.time() - t0)


def print_info(*args, **kwargs):
    pass


def print_title(*args, **kwargs):
    pass


def print_subtitle(*args, **kwargs):
    pass


# ---- file helpers the reference's parameter-study writer / reader call (oracle/gen_golden_r2.py) ----------
# pysitk itself is not installed; these follow pysitk.python_helper's documented behaviour (text written as
# given; arrays appended with np.savetxt(fmt="%.10e", delimiter="\t")).
def file_exists(path, extension=""):
    import os
    return os.path.isfile(path + extension)


def directory_exists(path):
    import os
    return os.path.isdir(path)


def create_directory(path, delete_files=False):
    import os
    if not os.path.isdir(path):
        os.makedirs(path)


def is_float(text):
    try:
        float(text)
        return True
    except ValueError:
        return False


def write_to_file(path, text, access_mode="w", verbose=True):
    import os
    directory = os.path.dirname(path)
    if directory and not os.path.isdir(directory):
        os.makedirs(directory)
    with open(path, access_mode) as fh:
        fh.write(text)


def write_array_to_file(path, array, format="%.10e", delimiter="\t", access_mode="a", verbose=True):
    import numpy as np
    with open(path, access_mode) as fh:
        np.savetxt(fh, array, fmt=format, delimiter=delimiter)


def read_file_line_by_line(path):
    with open(path) as fh:
        return fh.readlines()


def get_time_stamp():
    # two space-separated tokens: the reference's append check drops them with header.split(" ")[1:-2]
    # (nsol/solver_parameter_study.py:117-118, "ignore date info")
    return time.strftime("%Y-%m-%d %H:%M:%S")
