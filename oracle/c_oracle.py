"""ctypes wrapper of ``oracle/loglr_ref.c`` (the plain-C restatement of the
reference's per-walker chain).  TEST / BASELINE INFRASTRUCTURE ONLY -- see the
header of ``gwin_oracle.py``.  PARITY UNPINNED upstream."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle_loglr.so")
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return _LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        lib = C.CDLL(_LIB)
        lib.oracle_create.restype = C.c_void_p
        lib.oracle_create.argtypes = [C.c_int, C.c_long, C.c_double, C.c_double,
                                      C.c_double, C.c_double, C.c_double,
                                      C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_double]
        lib.oracle_destroy.argtypes = [C.c_void_p]
        lib.oracle_lognl.restype = C.c_double
        lib.oracle_lognl.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_cutoffs.argtypes = [C.c_void_p, C.POINTER(C.c_long),
                                       C.POINTER(C.c_long)]
        lib.oracle_loglr.restype = C.c_int
        lib.oracle_loglr.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = lib
    return _lib


class COracle(object):
    """Same inputs as ``gwin_oracle.GaussianNoiseOracle``; evaluates batches on
    ``nthreads`` host threads."""

    def __init__(self, data, delta_f, epoch, f_lower, psds=None, f_upper=None,
                 norm=None, waveform_f_lower=None):
        import gwin_oracle as O
        lib = load()
        self.dets = list(data.keys())
        arr = np.ascontiguousarray(np.stack([np.asarray(data[d], dtype=np.complex128)
                                             for d in self.dets]))
        nd, nf = arr.shape
        psd = None
        if psds is not None:
            psd = np.ascontiguousarray(np.stack([np.asarray(psds[d], dtype=np.float64)
                                                 for d in self.dets]))
        D = [O.Detector(d) for d in self.dets]
        resp = np.ascontiguousarray(np.stack([d.response for d in D]))
        loc = np.ascontiguousarray(np.stack([d.location for d in D]))
        with np.errstate(divide="ignore"):
            self._h = lib.oracle_create(
                nd, nf, float(delta_f), float(epoch), float(f_lower),
                float(f_upper or 0.0),
                float(f_lower if waveform_f_lower is None else waveform_f_lower),
                resp.ctypes.data, loc.ctypes.data, arr.ctypes.data,
                psd.ctypes.data if psd is not None else None,
                float(norm or 0.0))
        if not self._h:
            raise ValueError("oracle_create failed (kmax <= kmin?)")
        self._lib = lib
        self.n_det = nd

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.oracle_destroy(self._h)
            self._h = None

    def lognl(self):
        return self._lib.oracle_lognl(self._h, None)

    def batch_loglr(self, params, mode=0, nthreads=1):
        P = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 13)
        W = P.shape[0]
        lr = np.empty(W)
        hd = np.empty((W, self.n_det), dtype=np.complex128)
        hh = np.empty((W, self.n_det))
        self._lib.oracle_loglr(self._h, int(mode), W, P.ctypes.data,
                               lr.ctypes.data, hd.ctypes.data, hh.ctypes.data,
                               int(nthreads))
        return lr, hd, hh
