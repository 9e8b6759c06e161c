"""print_time, the one helper of lavis/compression/pruners/utils.py (:6-18) the hot path keeps."""
import functools
import time


def print_time(func):
    @functools.wraps(func)
    def timed(*args, **kwargs):
        t0 = time.time()
        out = func(*args, **kwargs)
        print(f"{func.__qualname__} took {time.time() - t0:.4f} s")
        return out
    return timed
