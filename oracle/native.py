"""ctypes front-end of the C oracle (numpy in / numpy out).  TEST INFRASTRUCTURE ONLY.

Function names and argument order are those of the reference's pybind module
``pointnet2._ext`` (lib/pointnet2/_ext_src/src/bindings.cpp:6-19); outputs are allocated
here exactly as the reference's C++ wrappers allocate them (zero-filled).
"""
import ctypes
import os

import numpy as np

from . import build as _build

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = _build.OUT
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(_build.SRC):
            path = _build.build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _out(shape, dtype):
    a = np.zeros(shape, dtype=dtype)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def num_threads():
    return int(lib().s2c_oracle_num_threads())


def opt_n_threads(n):
    return int(lib().s2c_oracle_opt_n_threads(int(n)))


def furthest_point_sampling(points, nsamples):
    points, pp = _f(points)
    B, N, _ = points.shape
    out, op = _out((B, nsamples), np.int32)
    lib().s2c_oracle_furthest_point_sampling(B, N, int(nsamples), pp, op)
    return out


def gather_points(points, idx):
    points, pp = _f(points)
    idx, ip = _i(idx)
    B, C, N = points.shape
    m = idx.shape[1]
    out, op = _out((B, C, m), np.float32)
    lib().s2c_oracle_gather_points(B, C, N, m, pp, ip, op)
    return out


def gather_points_grad(grad_out, idx, n):
    grad_out, gp = _f(grad_out)
    idx, ip = _i(idx)
    B, C, m = grad_out.shape
    out, op = _out((B, C, n), np.float32)
    lib().s2c_oracle_gather_points_grad(B, C, int(n), m, gp, ip, op)
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    new_xyz, cp = _f(new_xyz)
    xyz, pp = _f(xyz)
    B, M, _ = new_xyz.shape
    n = xyz.shape[1]
    out, op = _out((B, M, nsample), np.int32)
    lib().s2c_oracle_ball_query(B, n, M, ctypes.c_float(float(radius)), int(nsample), cp, pp, op)
    return out


def group_points(points, idx):
    points, pp = _f(points)
    idx, ip = _i(idx)
    B, C, N = points.shape
    _, npoints, nsample = idx.shape
    out, op = _out((B, C, npoints, nsample), np.float32)
    lib().s2c_oracle_group_points(B, C, N, npoints, nsample, pp, ip, op)
    return out


def group_points_grad(grad_out, idx, n):
    grad_out, gp = _f(grad_out)
    idx, ip = _i(idx)
    B, C, npoints, nsample = grad_out.shape
    out, op = _out((B, C, n), np.float32)
    lib().s2c_oracle_group_points_grad(B, C, int(n), npoints, nsample, gp, ip, op)
    return out


def three_nn(unknown, known):
    unknown, up = _f(unknown)
    known, kp = _f(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2, dp = _out((B, n, 3), np.float32)
    idx, ip = _out((B, n, 3), np.int32)
    lib().s2c_oracle_three_nn(B, n, m, up, kp, dp, ip)
    return dist2, idx


def three_interpolate(points, idx, weight):
    points, pp = _f(points)
    idx, ip = _i(idx)
    weight, wp = _f(weight)
    B, C, m = points.shape
    n = idx.shape[1]
    out, op = _out((B, C, n), np.float32)
    lib().s2c_oracle_three_interpolate(B, C, m, n, pp, ip, wp, op)
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, gp = _f(grad_out)
    idx, ip = _i(idx)
    weight, wp = _f(weight)
    B, C, n = grad_out.shape
    out, op = _out((B, C, m), np.float32)
    lib().s2c_oracle_three_interpolate_grad(B, C, n, int(m), gp, ip, wp, op)
    return out
