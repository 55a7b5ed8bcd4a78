"""NumPy restatement of the reference's path -> trajectory-point expansion (TEST INFRASTRUCTURE ONLY).

Restated from ``GraceRIGV3.py``:
* ``evaluateTraj``            ``:235-294``  primitive chain -> way-points (distance, depth, time[, var])
* ``edgePointsToTrajPoints``  ``:373-394``  way-points resampled at ``measRate`` and placed in 3-D
* ``pathToTrajPoints``        ``:396-427``  concatenate the edges of a path, de-duplicate rounded rows
The agent / graph objects are replaced by plain data: a path is a list of edges
``(start_xy, end_xy, prims)``; a primitive is ``(type, a, b, c)`` with type 0 spiral ``(dz, _, speed)``,
1 glide ``(gp, dz, speed)``, 2 swim ``(dist, speed)``, 3 flat dive ``(dz, speed)``.
Quirks kept on purpose: ``t_off += wpnts[-1][-1]`` adds the last *variance* when ``withVar`` (``:423``);
``uw`` only becomes true once depth > 0 and the variance is reset whenever depth <= 0.
"""
import numpy as np


def evaluate_traj(prims, variance_rate, with_var=True):
    timeTaken, distanceTraveled, var, depth = 0.0, 0.0, 0.0, 0.0
    uw = restart = False
    pnts = [(distanceTraveled, depth, timeTaken, var) if with_var else (distanceTraveled, depth, timeTaken)]
    for prim in prims:
        kind = int(prim[0])
        if kind == 0:
            dz, speed = prim[1], prim[3]
            timeTaken += abs(dz / speed)
            var += variance_rate * abs(dz / speed)
            depth = depth + dz
        elif kind == 1:
            gp, dz, speed = prim[1], prim[2], prim[3]
            timeTaken += abs(dz / speed)
            var += variance_rate * abs(dz / speed)
            distanceTraveled += dz / np.tan(gp)
            depth = depth + dz
        elif kind == 2:
            dist, speed = prim[1], prim[2]
            timeTaken += dist / speed
            var += variance_rate * uw * (dist / speed)
            distanceTraveled += dist
        elif kind == 3:
            dz, speed = prim[1], prim[2]
            timeTaken += abs(dz / speed)
            var += variance_rate * abs(dz / speed)
            depth = depth + dz
        if depth > 0:
            uw = restart = True
        elif depth <= 0.1 and restart:
            uw = restart = False
        if depth <= 0:
            var = 0
        pnts.append((distanceTraveled, depth, timeTaken, var) if with_var else (distanceTraveled, depth, timeTaken))
    return pnts


def edge_points_dense(ps, pf, pnts, meas_rate, t_off=0.0, with_var=True):
    ps = np.asarray(ps, float).reshape(-1, 1)
    pf = np.asarray(pf, float).reshape(-1, 1)
    diff = pf - ps
    b = np.arctan2(diff[1, 0], diff[0, 0])
    ddt = np.array(pnts)
    timePoints = np.arange(0, pnts[-1][2], 1 / meas_rate) + t_off
    timePoints.shape = (timePoints.shape[0], 1)
    extdist = np.interp(timePoints, ddt[:, 2] + t_off, ddt[:, 0])
    extdepth = np.interp(timePoints, ddt[:, 2] + t_off, ddt[:, 1])
    if with_var:
        extVar = np.interp(timePoints, ddt[:, 2] + t_off, ddt[:, 3])
        return np.concatenate((ps.T + np.zeros((extdepth.shape[0], ps.shape[0])), extdepth, timePoints, extVar), axis=1) + \
            extdist * np.array([np.cos(b), np.sin(b), 0, 0, 0])
    return np.concatenate((ps.T + np.zeros((extdepth.shape[0], ps.shape[0])), extdepth, timePoints), axis=1) + \
        extdist * np.array([np.cos(b), np.sin(b), 0, 0])


def path_to_traj_points(path, variance_rate, meas_rate, dense=False, t_off=0.0, with_var=False):
    """path: list of (start_xy, end_xy, prims).  Returns (k, 4 | 5) rows x, y, z, t[, var]."""
    ncol = 5 if with_var else 4
    pnts3D = np.zeros((0, ncol))
    densePoints = None
    for ps, pf, prims in path:
        wpnts = evaluate_traj(prims, variance_rate, with_var)
        if dense:
            ep = edge_points_dense(ps, pf, wpnts, meas_rate, t_off=t_off, with_var=with_var)
            densePoints = ep if densePoints is None else np.concatenate((densePoints, ep))
        ps_ = np.asarray(ps, float).reshape(-1, 1)
        pf_ = np.asarray(pf, float).reshape(-1, 1)
        diff = pf_ - ps_
        b = np.arctan2(diff[1, 0], diff[0, 0])
        ddt = np.array(wpnts)
        ddt[:, 2] = ddt[:, 2] + t_off
        if with_var:
            temp = np.concatenate((ps_.T + np.zeros((ddt.shape[0], ps_.shape[0])), ddt[:, 1:4]), axis=1) + \
                ddt[:, 0:1] * np.array([np.cos(b), np.sin(b), 0, 0, 0])
        else:
            temp = np.concatenate((ps_.T + np.zeros((ddt.shape[0], ps_.shape[0])), ddt[:, 1:3]), axis=1) + \
                ddt[:, 0:1] * np.array([np.cos(b), np.sin(b), 0, 0])
        pnts3D = np.concatenate((pnts3D, temp))
        t_off += wpnts[-1][-1]
    src = densePoints if dense else pnts3D
    _, ind = np.unique(np.round(src, 4), axis=0, return_index=True)
    return src[np.sort(ind), :]
