"""TEST INFRASTRUCTURE ONLY — CPU oracle of the step *behind* the hot path: turning the predicted
route / speed waypoints into the quantities the agent's PID controllers consume.

Restates, in numpy, the reference's
  * ``DrivingModel.equal_spacing_route``            simlingo_training/models/driving.py:330-342
  * ``LingoAgent.interpolate_waypoints``            team_code/agent_simlingo.py:960-1003
  * ``LingoAgent.control_pid``                      team_code/agent_simlingo.py:915-958
  * ``PIDController``                               team_code/transfuser_utils.py:334-356
  * ``LateralPIDController``                        team_code/nav_planner.py:72-140
and the third-party piece they call, ``scipy.interpolate.PchipInterpolator`` (scipy is not pinned by the reference's
``environment.yaml``; 1.18.1 in this image): Fritsch-Carlson slopes with the three-point end rule
(``_find_derivatives`` / ``_edge_case``) evaluated as a cubic Hermite polynomial in power form.

Pinned by ``tests/golden/postprocess.npz`` = outputs of the reference's own functions (source extracted from
``/root/reference`` and executed against numpy + scipy by ``tests/golden/make_golden_postprocess.py``).
Only ``tests/`` may import this module."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

# team_code/config_simlingo.py:12-25,45-48 and the LateralPIDController defaults (nav_planner.py:78), the
# configuration the released agent runs with (agent_simlingo.py:173-178: ``LateralPIDController(inference_mode=False)``)
CONFIG: Dict[str, float] = dict(
    carla_fps=20, wp_dilation=1, data_save_freq=5, brake_speed=0.4, brake_ratio=1.1, clip_delta=1.0, clip_throttle=1.0,
    speed_kp=1.75, speed_ki=1.0, speed_kd=2.0, speed_n=20,
    lat_kp=3.118357247806046, lat_kd=1.3782508892109167, lat_ki=0.6406067986034124, lat_speed_scale=0.9755321901954155,
    lat_speed_offset=1.9152884533402488, lat_n=6)


def arc_length(points: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """origin-prefixed polyline and its cumulative arc length, in the dtype numpy gives the reference: float32 points
    keep float32 segment lengths / running sums, the 1e-4·k de-duplication offset is added in double and rounded back"""
    pts = np.asarray(points)
    poly = np.concatenate((np.zeros_like(pts[:1]), pts))
    seg = np.zeros(len(poly), poly.dtype)
    d = poly[1:] - poly[:-1]
    seg[1:] = np.sqrt((d * d).sum(1, dtype=poly.dtype))
    arc = np.cumsum(seg, dtype=poly.dtype)
    arc = (arc.astype(np.float64) + np.arange(len(arc)) * 1e-4).astype(poly.dtype)
    return poly, arc


def equal_spacing_route(points: np.ndarray, n_out: int = 20) -> np.ndarray:
    """driving.py:330-342: piecewise-linear resampling at arc lengths 0, 1, …, n_out-1 (clamped at both ends)"""
    poly, arc = arc_length(points)
    grid = np.arange(n_out, dtype=np.float64)
    return np.stack([np.interp(grid, arc, poly[:, 0]), np.interp(grid, arc, poly[:, 1])], axis=1)


def pchip_slopes(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """PchipInterpolator._find_derivatives for x [n] (n >= 3), y [n, d]"""
    h = np.diff(x)[:, None]
    m = np.diff(y, axis=0) / h
    d = np.zeros_like(y)
    flat = (np.sign(m[1:]) != np.sign(m[:-1])) | (m[1:] == 0) | (m[:-1] == 0)
    w1, w2 = 2 * h[1:] + h[:-1], h[1:] + 2 * h[:-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        harmonic = (w1 / m[:-1] + w2 / m[1:]) / (w1 + w2)
        d[1:-1] = np.where(flat, 0.0, 1.0 / harmonic)

    def end(h0, h1, m0, m1):
        e = ((2 * h0 + h1) * m0 - h0 * m1) / (h0 + h1)
        wrong_sign = np.sign(e) != np.sign(m0)
        overshoot = (np.sign(m0) != np.sign(m1)) & (np.abs(e) > 3 * np.abs(m0))
        return np.where(wrong_sign, 0.0, np.where(overshoot, 3 * m0, e))

    d[0] = end(h[0], h[1], m[0], m[1])
    d[-1] = end(h[-1], h[-2], m[-1], m[-2])
    return d


def pchip_eval(x: np.ndarray, y: np.ndarray, q: np.ndarray) -> np.ndarray:
    """cubic Hermite in scipy's power form (CubicHermiteSpline.__init__ + PPoly evaluation, last piece extrapolates)"""
    d = pchip_slopes(x, y)
    h = np.diff(x)[:, None]
    slope = np.diff(y, axis=0) / h
    t = (d[:-1] + d[1:] - 2 * slope) / h
    c0, c1, c2, c3 = t / h, (slope - d[:-1]) / h - t, d[:-1], y[:-1]
    i = np.clip(np.searchsorted(x, q, side="right") - 1, 0, len(x) - 2)
    s = (q - x[i])[:, None]
    return c3[i] + c2[i] * s + c1[i] * (s * s) + c0[i] * (s * s * s)


def interpolate_waypoints(waypoints: np.ndarray) -> np.ndarray:
    """agent_simlingo.py:960-1003: PCHIP through the origin-prefixed route, sampled every 0.1 m of arc length"""
    poly, arc = arc_length(waypoints)
    q = np.arange(0.1, arc[-1], 0.1)
    if q.shape[0] == 0:
        return poly[None, -1]
    return pchip_eval(arc.astype(np.float64), poly.astype(np.float64), q)


def control_inputs(route: np.ndarray, speed_wps: np.ndarray, speed: np.float32, cfg=CONFIG) -> Tuple[np.float32, float, np.ndarray]:
    """the stateless half of ``control_pid``: (desired speed, scaled heading error, aim point)"""
    one_second = int(cfg["carla_fps"] // (cfg["wp_dilation"] * cfg["data_save_freq"]))
    half_second = one_second // 2
    desired = np.linalg.norm(speed_wps[half_second - 2] - speed_wps[one_second - 2]) * 2.0  # agent_simlingo.py:944-946
    interp = interpolate_waypoints(route)
    kmh = np.float32(speed) * 3.6  # nav_planner.py:113-119 (inference_mode False)
    n = int(min(np.clip(cfg["lat_speed_scale"] * kmh + cfg["lat_speed_offset"], 24, 105), interp.shape[0] - 1))
    aim = np.asarray(interp[min(n, len(interp) - 1)], np.float64)
    yaw = np.arctan2(aim[1], aim[0]) % (2 * np.pi)
    yaw = yaw if yaw < np.pi else yaw - 2 * np.pi
    return desired, float(yaw * 180.0 / np.pi / 90.0), aim


class PIDState:
    """the two error windows of ``t_u.PIDController(n=20)`` (team_code/transfuser_utils.py:334-356) and
    ``LateralPIDController(n=6)`` (team_code/nav_planner.py:72-140)"""

    def __init__(self, cfg=CONFIG):
        self.cfg = cfg
        self.speed_window = [0] * int(cfg["speed_n"])
        self.turn_window = []

    def step(self, desired, heading_error: float, speed) -> Tuple[float, float, bool]:
        cfg = self.cfg
        brake = bool((desired < cfg["brake_speed"]) or ((speed / desired) > cfg["brake_ratio"]))
        delta = np.clip(desired - speed, 0.0, cfg["clip_delta"])
        self.speed_window = (self.speed_window + [delta])[-int(cfg["speed_n"]):]
        integral = np.mean(self.speed_window)  # transfuser_utils.py:346-356 (the agent's speed controller is t_u.PIDController)
        derivative = self.speed_window[-1] - self.speed_window[-2]
        throttle = np.clip(cfg["speed_kp"] * delta + cfg["speed_ki"] * integral + cfg["speed_kd"] * derivative, 0.0, cfg["clip_throttle"])
        throttle = throttle if not brake else 0.0
        self.turn_window = (self.turn_window + [heading_error])[-int(cfg["lat_n"]):]
        d = 0.0 if len(self.turn_window) == 1 else self.turn_window[-1] - self.turn_window[-2]
        steer = np.clip(cfg["lat_kp"] * heading_error + cfg["lat_kd"] * d + cfg["lat_ki"] * np.mean(self.turn_window), -1.0, 1.0).item()
        return round(float(np.clip(steer, -1.0, 1.0)), 3), float(throttle), brake


def control_pid(state: PIDState, route: np.ndarray, speed_wps: np.ndarray, speed) -> Tuple[float, float, bool]:
    """agent_simlingo.py:915-958 for one tick (route [20,2], speed_wps [10,2], float32 as the model returns them)"""
    desired, heading, _ = control_inputs(route, speed_wps, speed, state.cfg)
    return state.step(desired, heading, np.float32(speed))
