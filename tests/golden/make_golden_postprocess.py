"""Generates ``tests/golden/postprocess.npz`` by running the REFERENCE's own post-processing code in this container
(the step right behind the hot path).  The agent module cannot be imported (it needs carla, cv2,
leaderboard, ...), so the class / method definitions are cut out of the reference sources with ``ast`` and executed
unmodified against numpy + scipy:

  team_code/agent_simlingo.py      LingoAgent.control_pid, LingoAgent.interpolate_waypoints
  team_code/nav_planner.py         LateralPIDController
  team_code/transfuser_utils.py    PIDController                (``t_u.PIDController`` at agent_simlingo.py:173)
  team_code/config_simlingo.py     GlobalConfig
  simlingo_training/models/driving.py   DrivingModel.equal_spacing_route

Two closed-loop episodes (the PID windows carry state from tick to tick) over seeded synthetic predictions, including
the edge cases the code handles: an all-zero route, repeated points, routes shorter than the look-ahead, a stopped
speed prediction (brake), high speed (longer look-ahead).

    python tests/golden/make_golden_postprocess.py
"""
import ast
import os
import types
from collections import deque
from copy import deepcopy

import numpy as np
import torch
from scipy.interpolate import PchipInterpolator

REF = "/root/reference"


def cut(path, cls, names=None):
    """source of class ``cls`` (optionally only the methods ``names``, re-wrapped in an empty class) from a reference file"""
    src = open(os.path.join(REF, path)).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == cls)
    if names is None:
        return ast.get_source_segment(src, node)
    import textwrap
    body = [textwrap.dedent(" " * m.col_offset + ast.get_source_segment(src, m)) for m in node.body
            if isinstance(m, ast.FunctionDef) and m.name in names]
    return f"class {cls}:\n" + "\n".join(textwrap.indent(b, "    ") for b in body)


ns = {"np": np, "deque": deque, "deepcopy": deepcopy, "PchipInterpolator": PchipInterpolator, "torch": torch}
exec(cut("team_code/nav_planner.py", "LateralPIDController"), ns)
exec(cut("team_code/transfuser_utils.py", "PIDController"), ns)
exec(cut("team_code/config_simlingo.py", "GlobalConfig"), ns)
exec(cut("team_code/agent_simlingo.py", "LingoAgent", ["control_pid", "interpolate_waypoints"]), ns)
exec(cut("simlingo_training/models/driving.py", "DrivingModel", ["equal_spacing_route"]), ns)


def new_agent():
    a = ns["LingoAgent"]()
    cfg = a.config = ns["GlobalConfig"]()
    a.speed_controller = ns["PIDController"](k_p=cfg.speed_kp, k_i=cfg.speed_ki, k_d=cfg.speed_kd, n=cfg.speed_n)  # agent_simlingo.py:173
    a.turn_controller = ns["LateralPIDController"](inference_mode=False)                                        # agent_simlingo.py:178
    return a


def episode(seed, ticks):
    """synthetic predictions shaped like the model's outputs: route = cumsum of ~1 m steps with a slowly turning
    heading, speed waypoints = cumsum of forward steps proportional to the target speed"""
    g = np.random.default_rng(seed)
    routes, wps, speeds = [], [], []
    heading, speed = 0.0, 2.0
    for t in range(ticks):
        curv = g.normal(0, 0.012)
        ang = heading + np.cumsum(np.full(20, curv) + g.normal(0, 0.004, 20))
        step = np.abs(g.normal(1.0, 0.1, 20))
        route = np.cumsum(np.stack([step * np.cos(ang), step * np.sin(ang)], 1), 0)
        target = max(0.0, 4.0 + 3.0 * np.sin(t / 5.0) + g.normal(0, 0.3))
        w = np.cumsum(np.stack([np.full(10, target / 4.0) + g.normal(0, 0.02, 10), g.normal(0, 0.02, 10)], 1), 0)
        if t % 11 == 5:
            route[:] = 0.0                                  # all points at the origin: the fallback branch
        if t % 11 == 7:
            route[5:9] = route[5]                           # repeated points (arc length only grows by the 1e-4 offset)
        if t % 11 == 9:
            route *= 0.08                                   # 1.6 m long: shorter than the 2.5 m look-ahead
        if t % 13 == 3:
            w[:] = w[0]                                     # predicted stop -> brake
        speed = float(np.clip(target * g.uniform(0.6, 1.15), 0.0, 20.0))
        if t % 13 == 8:                                     # 43-50 km/h: look-ahead beyond the default 24
            speed = 12.0 + t / 20.0
            w = w * (speed / max(target, 0.5)) * g.uniform(0.95, 1.3)
        if t % 7 == 6:
            route[:, 1] *= -1.0
        heading = 0.08 * np.sin(t / 3.0) + (np.pi if t % 17 == 16 else 0.0)  # once: route pointing backwards
        routes.append(route.astype(np.float32)); wps.append(w.astype(np.float32)); speeds.append(np.float32(speed))
    return np.stack(routes), np.stack(wps), np.asarray(speeds, np.float32)


out = {}
for name, seed, ticks in (("ep0", 0, 45), ("ep1", 1, 30)):
    routes, wps, speeds = episode(seed, ticks)
    agent = new_agent()
    drv = ns["DrivingModel"]()
    controls, counts, aims, equal = [], [], [], []
    for r, w, s in zip(routes, wps, speeds):
        interp = agent.interpolate_waypoints(torch.from_numpy(r)[None][0].numpy().squeeze())
        counts.append(interp.shape[0])
        aims.append(interp[min(24, interp.shape[0] - 1)])
        steer, throttle, brake = agent.control_pid(torch.from_numpy(r)[None], torch.from_numpy(np.asarray([s])), torch.from_numpy(w)[None])
        controls.append([float(steer), float(throttle), float(brake)])
        equal.append(drv.equal_spacing_route(torch.from_numpy(r)))
    out.update({f"{name}_route": routes, f"{name}_speed_wps": wps, f"{name}_speed": speeds,
                f"{name}_controls": np.asarray(controls, np.float64), f"{name}_interp_count": np.asarray(counts, np.int32),
                f"{name}_interp_at24": np.asarray(aims, np.float64), f"{name}_equal_spacing": np.asarray(equal, np.float64),
                f"{name}_interp_full_t0": agent.interpolate_waypoints(routes[0])})
    print(name, "brakes", int(np.asarray(controls)[:, 2].sum()), "counts", min(counts), max(counts),
          "interior steer", int((np.abs(np.asarray(controls)[:, 0]) < 1).sum()), "interior throttle",
          int(((np.asarray(controls)[:, 1] > 0) & (np.asarray(controls)[:, 1] < 1)).sum()))
np.savez_compressed(os.path.join(os.path.dirname(__file__), "postprocess.npz"), **out)
