"""Host-side path table for the continuous task: two circles and two straights, as
reference environment/path_tracking/path_tracking_env.py:28-56 builds them
(generate_circle_path :103-122 = exact-arc unicycle roll-out with theta wrapped to [-pi,pi);
generate_direct_path :176-188 = linspace over 3-vectors, hence int(dist/sep)+1 nodes).
Paths are static inputs of the hot path, generated once on the host in numpy."""
import numpy as np


def _wrap(th):
    return np.mod(th + np.pi, 2 * np.pi) - np.pi


def circle_path(start_pose, time_step=0.25, node_separation=0.2, ccw=False):
    x = np.array(start_pose, dtype=np.float64)
    radius = abs(x[1])
    v = node_separation / time_step
    w = -v / (radius * 2) * 2
    if ccw:
        w = -w
    num_steps = int(np.pi * 2 / abs(w) / time_step) - 1
    nodes = [x.copy()]
    for _ in range(num_steps):
        dx_l = (v / w) * np.sin(w * time_step)
        dy_l = (v / w) * (1 - np.cos(w * time_step))
        c, s = np.cos(x[2]), np.sin(x[2])
        x = x + np.array([c * dx_l - s * dy_l, s * dx_l + c * dy_l, w * time_step])
        x[2] = _wrap(x[2])
        nodes.append(x.copy())
    return np.stack(nodes, axis=1)          # 3 x L


def direct_path(start, end, node_separation=0.2):
    start, end = np.asarray(start, dtype=np.float64), np.asarray(end, dtype=np.float64)
    angle = np.arctan2(end[1] - start[1], end[0] - start[0])
    a = np.array([start[0], start[1], angle])
    b = np.array([end[0], end[1], angle])
    num_nodes = int(np.linalg.norm(end - start) / node_separation) + 1   # 3-vectors, like the reference
    return np.linspace(a, b, num_nodes, endpoint=True).T


def continuous_task_paths(radius=4.0, time_step=0.25, node_separation=0.2):
    """The four paths of reset_continuous_task for a robot starting at (0,-radius,pi)."""
    p0 = circle_path([0, -radius, np.pi], time_step, node_separation)
    p1 = direct_path([0, -radius, np.pi / 2], [0, radius, np.pi / 2], node_separation)
    p2 = circle_path([0, radius, np.pi], time_step, node_separation, ccw=True)
    p3 = direct_path([0, radius, -np.pi / 2], [0, -radius, -np.pi / 2], node_separation)
    return [p0, p1, p2, p3]


def pack_paths(paths):
    """-> (table float64 [P,3,stride], lens int32 [P])"""
    stride = max(p.shape[1] for p in paths)
    tab = np.zeros((len(paths), 3, stride), dtype=np.float64)
    lens = np.zeros(len(paths), dtype=np.int32)
    for i, p in enumerate(paths):
        tab[i, :, :p.shape[1]] = p
        lens[i] = p.shape[1]
    return tab, lens
