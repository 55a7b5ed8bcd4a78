"""Candidate generation on the device: the host loops the reference runs right before its
information-gain operators (``GraceRIGV3.py:235-294`` ``evaluateTraj``, ``:373-394``
``edgePointsToTrajPoints``, ``:396-427`` ``pathToTrajPoints`` and the fidelity labelling of
``:508-512``) for MANY candidate paths in one kernel launch.

A path is a list of edges ``(start_xy, end_xy, prims)``; a primitive is ``(type, a, b, c)`` with type
0 spiral ``(dz, _, speed)``, 1 glide ``(pitch, dz, speed)``, 2 swim ``(dist, speed)``, 3 flat dive
``(dz, speed)`` -- ``edges_of_path`` extracts that from the reference's ``V`` / ``E`` / ``path`` objects.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from .core import GPCore

_core = {}


def _handle(device=0):
    if device not in _core:
        _core[device] = GPCore(L.KIND_SF_RBF, 1, device)
    return _core[device]


def edges_of_path(agent, V, E, path):
    """The ``(start_xy, end_xy, prims)`` list of one reference path (``GraceRIGV3.py:402-409``):
    ``path`` holds ``(idx1, idx2, edge_idx, ...)`` entries, ``E[(idx1, idx2)][edge_idx]`` ends with the
    primitive chain whose first entries are names from ``agent.legTypes``."""
    names = list(agent.legTypes)
    out = []
    for data in path:
        idx1, idx2, edge_idx = data[0:3]
        prims = E[(idx1, idx2)][edge_idx][-1]
        enc = []
        for pr in prims:
            k = names.index(pr[0])
            vals = [float(v) for v in pr[1:]] + [0.0, 0.0, 0.0]
            enc.append((float(k), vals[0], vals[1], vals[2]))
        ps, pf = np.asarray(V[idx1].state, float).ravel(), np.asarray(V[idx2].state, float).ravel()
        out.append((ps[:2], pf[:2], enc))
    return out


def paths_to_points(paths, variance_rate, meas_rate, dense=False, with_var=True, t_off=0.0, fid_levels=None,
                    max_pts=64, device=0):
    """Returns (points, fids): ``points[c]`` is the (k_c, 5) array x, y, z, t, var of path c (var = 0 when
    ``with_var`` is False -- the reference's 4-column form is ``points[c][:, :4]``), ``fids[c]`` the
    fidelity index per point (None without ``fid_levels``)."""
    Cn = len(paths)
    edge_off = np.zeros(Cn + 1, dtype=np.int64)
    xy, prim_off, prims = [], [0], []
    for c, path in enumerate(paths):
        edge_off[c + 1] = edge_off[c] + len(path)
        for ps, pf, pr in path:
            xy.append([float(ps[0]), float(ps[1]), float(pf[0]), float(pf[1])])
            for q in pr:
                q = list(q) + [0.0] * (4 - len(q))
                prims.append([float(v) for v in q[:4]])
            prim_off.append(len(prims))
    xy = np.ascontiguousarray(np.asarray(xy, dtype=np.float64).reshape(-1, 4))
    prim_off = np.asarray(prim_off, dtype=np.int64)
    prims = np.ascontiguousarray(np.asarray(prims, dtype=np.float64).reshape(-1, 4))
    if prims.shape[0] == 0:
        prims = np.zeros((1, 4))
    pts = np.zeros((max(Cn, 1), max_pts, 5))
    fid = np.zeros((max(Cn, 1), max_pts)) if fid_levels is not None else None
    counts = np.zeros(max(Cn, 1), dtype=np.int64)
    fl = None if fid_levels is None else np.ascontiguousarray(np.asarray(fid_levels, dtype=np.float64)[:2])
    h = _handle(device)
    rc = h.lib.gpc_traj_points(h.h, Cn, L.lptr(edge_off), L.dptr(xy), L.lptr(prim_off), L.dptr(prims),
                               float(variance_rate), float(meas_rate), int(bool(dense)), int(bool(with_var)), float(t_off),
                               L.dptr(fl), int(max_pts), L.dptr(pts), L.dptr(fid), L.lptr(counts))
    L.check(h.lib, h.h, rc)
    points = [pts[c, :counts[c]].copy() for c in range(Cn)]
    fids = None if fid is None else [fid[c, :counts[c]].copy() for c in range(Cn)]
    return points, fids


def candidate_rows(paths, variance_rate, meas_rate, fid_levels, dense=False, max_pts=64, device=0):
    """(k_c, 4) rows x, y, z, fidelity for every path -- what ``gpc_ig_*`` / ``infogain`` consume
    (``calculatePathInfoEmu*`` build exactly this from ``pathToTrajPoints(withVar=True)``)."""
    pts, fids = paths_to_points(paths, variance_rate, meas_rate, dense=dense, with_var=True, fid_levels=fid_levels,
                                max_pts=max_pts, device=device)
    return [np.hstack([p[:, :3], f[:, None]]) for p, f in zip(pts, fids)]
