"""ctypes front end of the C oracle (oracle/cdf_exact.c, oracle/rangecoder_ref.c).  Test infrastructure."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_SRCS = [os.path.join(_HERE, "cdf_exact.c"), os.path.join(_HERE, "rangecoder_ref.c")]
_lib = None


def build(force: bool = False) -> str:
    """gcc the two C files into oracle/_build/liboracle.so (skipped when up to date)."""
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in _SRCS):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-o", _SO] + _SRCS + ["-lm"]
    subprocess.run(cmd, check=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        d = ctypes.c_double
        for name in ("log", "exp", "lgamma", "erfc", "ncdf"):
            fn = getattr(L, "sic_oracle_" + name)
            fn.restype, fn.argtypes = d, [d]
        L.sic_oracle_tcdf.restype, L.sic_oracle_tcdf.argtypes = d, [d, d]
        L.sic_oracle_exp_f32.restype, L.sic_oracle_exp_f32.argtypes = ctypes.c_float, [ctypes.c_float]
        p = ctypes.c_void_p
        L.sic_oracle_tables.restype = None
        L.sic_oracle_tables.argtypes = [ctypes.c_int, p, p, p, p, p, ctypes.c_int, ctypes.c_int, p]
        L.sic_oracle_rans_encode.restype = ctypes.c_long
        L.sic_oracle_rans_encode.argtypes = [p, ctypes.c_long, p, ctypes.c_int, ctypes.c_int, ctypes.c_long, p, ctypes.c_long]
        L.sic_oracle_rans_decode.restype = ctypes.c_int
        L.sic_oracle_rans_decode.argtypes = [p, ctypes.c_long, ctypes.c_long, p, ctypes.c_int, ctypes.c_int, ctypes.c_long, p]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def tcdf(t, nu):
    L = lib()
    t, nu = np.broadcast_arrays(np.asarray(t, np.float64), np.asarray(nu, np.float64))
    return np.array([L.sic_oracle_tcdf(float(a), float(b)) for a, b in zip(t.ravel(), nu.ravel())]).reshape(t.shape)


def ncdf(t):
    L = lib()
    t = np.asarray(t, np.float64)
    return np.array([L.sic_oracle_ncdf(float(a)) for a in t.ravel()]).reshape(t.shape)


def exp_f32(x):
    L = lib()
    x = np.asarray(x, np.float32)
    return np.array([L.sic_oracle_exp_f32(float(a)) for a in x.ravel()], np.float32).reshape(x.shape)


def build_tables(kind: str, sigma, nu, patch_of_row, mins, maxs):
    """uint16 CDF tables [n_rows, Lmax+1] (symbol axis last), rows beyond a patch's own L+1 are zero.

    kind 'gaussian' (T3, eval_selfcontained_entropy.py:37-47) or 'studentt' (T2, :51-61)."""
    sigma = np.ascontiguousarray(sigma, np.float32).ravel()
    n = sigma.size
    nu = np.ascontiguousarray(nu if nu is not None else np.zeros(n), np.float32).ravel()
    patch_of_row = np.ascontiguousarray(patch_of_row, np.int32).ravel()
    mins = np.ascontiguousarray(mins, np.int32)
    maxs = np.ascontiguousarray(maxs, np.int32)
    stride = int((maxs - mins).max()) + 2
    out = np.zeros((n, stride), np.uint16)
    lib().sic_oracle_tables(1 if kind == "studentt" else 0, _ptr(sigma), _ptr(nu), _ptr(patch_of_row), _ptr(mins),
                            _ptr(maxs), n, stride, _ptr(out))
    return out


def rans_encode(sym, tables, L, sym_per_row) -> bytes:
    sym = np.ascontiguousarray(sym, np.int32).ravel()
    tables = np.ascontiguousarray(tables, np.uint16)
    cap = 128 + 2 * sym.size + 16
    out = np.zeros(cap, np.uint8)
    n = lib().sic_oracle_rans_encode(_ptr(sym), sym.size, _ptr(tables), tables.shape[-1], int(L), int(sym_per_row), _ptr(out), cap)
    if n < 0:
        raise ValueError("oracle rANS encode failed (symbol out of range or buffer too small)")
    return out[:n].tobytes()


def rans_decode(data: bytes, n, tables, L, sym_per_row):
    tables = np.ascontiguousarray(tables, np.uint16)
    buf = np.frombuffer(data, np.uint8)
    sym = np.zeros(int(n), np.int32)
    rc = lib().sic_oracle_rans_decode(_ptr(buf), buf.size, int(n), _ptr(tables), tables.shape[-1], int(L), int(sym_per_row), _ptr(sym))
    if rc != 0:
        raise ValueError("oracle rANS decode failed (truncated stream)")
    return sym
