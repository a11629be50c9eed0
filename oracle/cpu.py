"""ctypes front end of oracle/ppt_oracle.c (numpy in, numpy out).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Function names follow the
reference functions they restate; each C function cites the reference lines.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libppt_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    src = os.path.join(_HERE, "ppt_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_max_threads.restype = ctypes.c_int
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_i64p)


def max_threads():
    return int(lib().orc_max_threads())


def set_threads(n):
    lib().orc_set_threads(ctypes.c_int(int(n)))


def farthest_point_sample(xyz, npoint, start):
    xyz, px = _f(xyz)
    B, N, _ = xyz.shape
    start, ps = _i(np.broadcast_to(np.asarray(start, dtype=np.int64), (B,)))
    out = np.empty((B, npoint), dtype=np.int64)
    lib().orc_fps(px, ps, out.ctypes.data_as(_i64p), B, N, npoint)
    return out


def square_distance(src, dst):
    src, ps = _f(src)
    dst, pd = _f(dst)
    B, S, _ = src.shape
    N = dst.shape[1]
    out = np.empty((B, S, N), dtype=np.float32)
    lib().orc_square_distance(ps, pd, out.ctypes.data_as(_f32p), B, S, N)
    return out


def knn_point(nsample, xyz, new_xyz, return_dist=False):
    xyz, px = _f(xyz)
    new_xyz, pq = _f(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    idx = np.empty((B, S, nsample), dtype=np.int64)
    dist = np.empty((B, S, nsample), dtype=np.float32)
    lib().orc_knn(px, pq, idx.ctypes.data_as(_i64p), dist.ctypes.data_as(_f32p), B, N, S, nsample)
    return (idx, dist) if return_dist else idx


def query_ball_point(radius, nsample, xyz, new_xyz):
    xyz, px = _f(xyz)
    new_xyz, pq = _f(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    thr = np.float32(float(radius) ** 2)  # torch casts the python double to fp32 (F7)
    idx = np.empty((B, S, nsample), dtype=np.int64)
    lib().orc_ball_query(px, pq, idx.ctypes.data_as(_i64p), ctypes.c_float(thr), B, N, S, nsample)
    return idx


def index_points(points, idx):
    points, pp = _f(points)
    B, N, C = points.shape
    idx, pi = _i(idx)
    M = int(np.prod(idx.shape[1:]))
    out = np.empty((B, M, C), dtype=np.float32)
    lib().orc_gather(pp, pi, out.ctypes.data_as(_f32p), B, N, C, M)
    return out.reshape(idx.shape + (C,))


def group_center(xyz, idx, center):
    xyz, px = _f(xyz)
    idx, pi = _i(idx)
    center, pc = _f(center)
    B, N, _ = xyz.shape
    _, G, K = idx.shape
    out = np.empty((B, G, K, 3), dtype=np.float32)
    lib().orc_group_center(px, pi, pc, out.ctypes.data_as(_f32p), B, N, G, K)
    return out


def three_nn(unknown, known):
    unknown, pu = _f(unknown)
    known, pk = _f(known)
    B, N, _ = unknown.shape
    S = known.shape[1]
    dist = np.empty((B, N, 3), dtype=np.float32)
    idx = np.empty((B, N, 3), dtype=np.int64)
    lib().orc_three_nn(pu, pk, dist.ctypes.data_as(_f32p), idx.ctypes.data_as(_i64p), B, N, S)
    return dist, idx


def three_interpolate(feats, idx, dist):
    feats, pf = _f(feats)
    idx, pi = _i(idx)
    dist, pd = _f(dist)
    B, S, D = feats.shape
    N = idx.shape[1]
    out = np.empty((B, N, D), dtype=np.float32)
    lib().orc_three_interpolate(pf, pi, pd, out.ctypes.data_as(_f32p), B, N, S, D)
    return out


def group_forward(xyz, num_group, group_size, start=0):
    """Group.forward (models/pointbert/dvae.py:159-181) -> (neighborhood, center, fps_idx, knn_idx)."""
    xyz, px = _f(xyz)
    B, N, _ = xyz.shape
    start, ps = _i(np.broadcast_to(np.asarray(start, dtype=np.int64), (B,)))
    nb = np.empty((B, num_group, group_size, 3), dtype=np.float32)
    ctr = np.empty((B, num_group, 3), dtype=np.float32)
    fidx = np.empty((B, num_group), dtype=np.int64)
    kidx = np.empty((B, num_group, group_size), dtype=np.int64)
    lib().orc_group_forward(px, ps, nb.ctypes.data_as(_f32p), ctr.ctypes.data_as(_f32p),
                            fidx.ctypes.data_as(_i64p), kidx.ctypes.data_as(_i64p),
                            B, N, num_group, group_size)
    return nb, ctr, fidx, kidx


def loader_fps_indices(point, npoint, start):
    """data/dataset_3d.py:40-61 restated (numpy, same statements) with the start index as an argument instead of the
    np.random.randint draw; returns the index array (the reference returns point[indices])."""
    N = point.shape[0]
    xyz = point[:, :3]
    centroids = np.zeros((npoint,))
    distance = np.ones((N,)) * 1e10
    farthest = int(start)
    for i in range(npoint):
        centroids[i] = farthest
        centroid = xyz[farthest, :]
        dist = np.sum((xyz - centroid) ** 2, -1)
        mask = dist < distance
        distance[mask] = dist[mask]
        farthest = np.argmax(distance, -1)
    return centroids.astype(np.int64)
