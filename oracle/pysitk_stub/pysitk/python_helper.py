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
