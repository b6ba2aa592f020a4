"""
tess_b200._lib -- locates and loads libtess_b200.so.  There is NO fallback: if the CUDA library
is missing the import of any compute entry point raises, by design (a silent CPU path would void
every parity and performance claim of this package).
"""
import os
import threading

from ._capi import CApi

HERE = os.path.dirname(os.path.abspath(__file__))
_LOCK = threading.Lock()
_API = None


def library_path():
    return os.environ.get("TESS_B200_LIB", os.path.join(HERE, "libtess_b200.so"))


def get_api():
    """The process-wide CApi.  Raises OSError with build instructions if the library is absent."""
    global _API
    with _LOCK:
        if _API is None:
            path = library_path()
            if not os.path.exists(path):
                raise OSError(
                    "tess_b200: CUDA library %s not found.  Build it with `python -m tess_b200.build` "
                    "(needs nvcc; sm_100a).  tess_b200 has no CPU fallback." % path)
            _API = CApi(path)
        return _API
