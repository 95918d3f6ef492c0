"""ctypes driver of r/gpb200_r_mock.so: r/shim.c compiled against the functional mock of the R C API
(r/mock/mock_r.c) and linked against libgpb200.so.  `MockR.call(name, *args)` does what R's
.Call(name, ...) does: looks the entry point up in the table the shim registered and invokes it; an
Rf_error inside the shim becomes a Python RError."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SO = os.path.join(ROOT, "r", "gpb200_r_mock.so")
REALSXP, INTSXP, VECSXP, STRSXP, NILSXP = 14, 13, 19, 16, 0


class RError(RuntimeError):
    pass


def build(force=False):
    srcs = [os.path.join(ROOT, "r", "shim.c"), os.path.join(ROOT, "r", "mock", "mock_r.c"),
            os.path.join(ROOT, "r", "mock", "Rinternals.h"), os.path.join(ROOT, "include", "gpb200.h")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.check_call(["bash", os.path.join(ROOT, "r", "mock", "build_mock.sh")], stdout=subprocess.DEVNULL)
    return SO


class MockR:
    def __init__(self):
        self.lib = C.CDLL(build())
        L = self.lib
        for name in ("mock_real_vector", "mock_int_vector", "mock_real_matrix", "mock_int_matrix", "mock_list", "mock_nil",
                     "mock_list_get", "mock_call", "mock_data"):
            getattr(L, name).restype = C.c_void_p
        L.mock_real_vector.argtypes = [C.c_void_p, C.c_long]
        L.mock_int_vector.argtypes = [C.c_void_p, C.c_long]
        L.mock_real_matrix.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.mock_int_matrix.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.mock_list.argtypes = [C.c_long]
        L.mock_list_set.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        L.mock_list_get.argtypes = [C.c_void_p, C.c_long]
        L.mock_call.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        L.mock_data.argtypes = [C.c_void_p]
        for name in ("mock_type", "mock_nrow", "mock_ncol"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = C.c_int
        L.mock_length.argtypes = [C.c_void_p]
        L.mock_length.restype = C.c_long
        L.mock_name.argtypes = [C.c_void_p, C.c_long]
        L.mock_name.restype = C.c_char_p
        L.mock_last_error.restype = C.c_char_p
        L.mock_registered.argtypes = [C.c_char_p]
        L.mock_registered.restype = C.c_int
        L.R_init_gpb200_r.argtypes = [C.c_void_p]
        L.R_init_gpb200_r(None)   # what R does after dyn.load

    # -- R values ---------------------------------------------------------------------------------
    def sexp(self, v):
        """Python value -> SEXP the way R would hold it: float arrays are double, int arrays are integer
        storage, 2-D arrays are matrices (column-major), lists are R lists, None is NULL."""
        L = self.lib
        if v is None:
            return L.mock_nil()
        if isinstance(v, (list, tuple)):
            out = L.mock_list(len(v))
            for i, e in enumerate(v):
                L.mock_list_set(out, i, self.sexp(e))
            return out
        a = np.asarray(v)
        integer = np.issubdtype(a.dtype, np.integer) or a.dtype == bool
        a = np.asfortranarray(a, dtype=np.int32 if integer else np.float64)
        if a.ndim == 2:
            f = L.mock_int_matrix if integer else L.mock_real_matrix
            return f(a.ctypes.data, a.shape[0], a.shape[1])
        a = np.ascontiguousarray(a.ravel())
        f = L.mock_int_vector if integer else L.mock_real_vector
        return f(a.ctypes.data, a.shape[0])

    def value(self, s):
        L = self.lib
        t = L.mock_type(s)
        n = L.mock_length(s)
        if t == NILSXP:
            return None
        if t in (REALSXP, INTSXP):
            ct = C.c_double if t == REALSXP else C.c_int
            buf = (ct * n).from_address(L.mock_data(s)) if n else []
            a = np.array(buf, dtype=np.float64 if t == REALSXP else np.int32)
            nr, nc = L.mock_nrow(s), L.mock_ncol(s)
            return a.reshape((nr, nc), order="F") if nr >= 0 else a
        if t == VECSXP:
            names = [L.mock_name(s, i).decode() for i in range(n)]
            vals = [self.value(L.mock_list_get(s, i)) for i in range(n)]
            return dict(zip(names, vals)) if all(names) else vals
        raise RError("unsupported result type %d" % t)

    def call(self, name, *args):
        arr = (C.c_void_p * len(args))(*[self.sexp(a) for a in args])
        out = self.lib.mock_call(name.encode(), len(args), arr)
        if not out:
            raise RError(self.lib.mock_last_error().decode())
        v = self.value(out)
        self.lib.mock_free_all()
        return v

    def registered(self, name):
        return self.lib.mock_registered(name.encode())
