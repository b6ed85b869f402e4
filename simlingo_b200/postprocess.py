"""Post-processing behind the hot path: from the model's ``(speed_wps, route)`` to vehicle controls.

Drop-in for ``LingoAgent.control_pid`` / ``interpolate_waypoints`` (reference ``team_code/agent_simlingo.py:915-1003``)
and for the per-item ``equal_spacing_route`` loop of ``DrivingModel.predict_step`` (``simlingo_training/models/
driving.py:290-295,330-342``).  The geometry — arc lengths, the PCHIP sample at the look-ahead index, the heading error,
the desired speed — is one kernel launch on the predictions where they already live (``slb_control_inputs``), followed
by ONE 64-byte read-back into pinned memory; the two PID controllers (``team_code/transfuser_utils.py:334-356``,
``team_code/nav_planner.py:72-140``) are a few scalar operations on windows of past errors and stay on the host,
written with the same numpy scalar types as the reference so that the control sequence is reproduced exactly.
There is no CPU fallback for the geometry."""
from __future__ import annotations

import ctypes as C
from collections import deque
from typing import Optional, Tuple

import numpy as np
import torch

from . import lib


class ControlParams(C.Structure):  # mirrors slb_control_params
    _fields_ = [("wp_a", C.c_int32), ("wp_b", C.c_int32), ("lookahead_scale", C.c_float), ("lookahead_offset", C.c_float),
                ("lookahead_min", C.c_float), ("lookahead_max", C.c_float), ("sample_step", C.c_double)]


class ControlConfig:
    """the fields of ``GlobalConfig`` (team_code/config_simlingo.py:12-25,45-48) and the ``LateralPIDController`` defaults
    (nav_planner.py:78) that ``control_pid`` reads; any object with these attributes (e.g. the agent's own config) works"""
    carla_fps, wp_dilation, data_save_freq = 20, 1, 5
    brake_speed, brake_ratio, clip_delta, clip_throttle = 0.4, 1.1, 1.0, 1.0
    speed_kp, speed_ki, speed_kd, speed_n = 1.75, 1.0, 2.0, 20
    lateral_kp, lateral_kd, lateral_ki = 3.118357247806046, 1.3782508892109167, 0.6406067986034124
    lateral_speed_scale, lateral_speed_offset, lateral_n = 0.9755321901954155, 1.9152884533402488, 6
    lookahead_min, lookahead_max, sample_step = 24.0, 105.0, 0.1


def _params(cfg, n_wps: int) -> ControlParams:
    one_second = int(cfg.carla_fps // (cfg.wp_dilation * cfg.data_save_freq))
    half_second = one_second // 2
    g = lambda name: getattr(cfg, name, getattr(ControlConfig, name))
    return ControlParams((half_second - 2) % n_wps, (one_second - 2) % n_wps, g("lateral_speed_scale"), g("lateral_speed_offset"),
                         g("lookahead_min"), g("lookahead_max"), g("sample_step"))


def _f32_cuda(t: torch.Tensor, what: str) -> torch.Tensor:
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"simlingo_b200.postprocess: {what} must be a CUDA tensor (no CPU fallback)")
    return t.detach().float().contiguous()


def control_inputs(route: torch.Tensor, speed_wps: torch.Tensor, speed: torch.Tensor, cfg=ControlConfig,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """route [B,N,2], speed_wps [B,K,2], speed [B] (m/s) -> float64 [B,8] on the GPU:
    (desired speed, heading error, aim x, aim y, number of 0.1 m samples, look-ahead index, speed, 0)"""
    route, speed_wps, speed = _f32_cuda(route, "route"), _f32_cuda(speed_wps, "speed_wps"), _f32_cuda(speed, "speed").reshape(-1)
    B, N, K = route.shape[0], route.shape[1], speed_wps.shape[1]
    if speed_wps.shape[0] != B or speed.shape[0] != B or route.shape[2] != 2 or speed_wps.shape[2] != 2:
        raise RuntimeError("simlingo_b200.postprocess: inconsistent shapes")
    out = torch.empty((B, 8), device=route.device, dtype=torch.float64) if out is None else out
    prm = _params(cfg, K)
    lib._check(lib.load().slb_control_inputs(lib._p(route), lib._p(speed_wps), lib._p(speed), B, N, K, C.byref(prm), lib._p(out),
                                             lib._stream()), "control_inputs")
    return out


def equal_spacing_route(route: torch.Tensor, n_out: int = 20) -> torch.Tensor:
    """route [B,N,2] -> float64 [B,n_out,2]: the reference's per-item numpy ``equal_spacing_route`` for the whole batch"""
    route = _f32_cuda(route, "route")
    out = torch.empty((route.shape[0], n_out, 2), device=route.device, dtype=torch.float64)
    lib._check(lib.load().slb_equal_spacing_route(lib._p(route), route.shape[0], route.shape[1], n_out, lib._p(out), lib._stream()),
               "equal_spacing_route")
    return out


class ControlPID:
    """``control_pid`` of the agent as an object: ``steer, throttle, brake = pid.control_pid(pred_route, gt_velocity,
    pred_speed_wps)`` (same argument order and return triple as agent_simlingo.py:878,915)."""

    def __init__(self, config=ControlConfig):
        self.config = config
        g = lambda name: getattr(config, name, getattr(ControlConfig, name))
        self.speed_gains = (g("speed_kp"), g("speed_ki"), g("speed_kd"))
        self.lateral_gains = (g("lateral_kp"), g("lateral_ki"), g("lateral_kd"))
        self.lateral_n = int(g("lateral_n"))
        self.speed_window = deque([0 for _ in range(int(g("speed_n")))], maxlen=int(g("speed_n")))
        self.turn_window = []
        self._host = None
        self._dev = None

    def _read_back(self, route, speed_wps, speed) -> np.ndarray:
        if self._host is None:
            self._host = torch.empty((1, 8), dtype=torch.float64).pin_memory()
            self._dev = torch.empty((1, 8), dtype=torch.float64, device=route.device)
        control_inputs(route, speed_wps, speed, self.config, out=self._dev)
        self._host.copy_(self._dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._host.numpy()[0]

    def control_pid(self, route_waypoints: torch.Tensor, velocity: torch.Tensor, speed_waypoints: torch.Tensor) -> Tuple[float, float, bool]:
        assert route_waypoints.size(0) == 1
        velocity = velocity if velocity.is_cuda else velocity.to(route_waypoints.device, non_blocking=True)
        row = self._read_back(route_waypoints, speed_waypoints, velocity)
        return self.step(np.float32(row[0]), float(row[1]), np.float32(row[6]))

    def step(self, desired_speed: np.float32, heading_error: float, speed: np.float32) -> Tuple[float, float, bool]:
        """the stateful tail of ``control_pid``: longitudinal PID on the clipped speed error, lateral PID on the heading
        error (agent_simlingo.py:948-958, transfuser_utils.py:346-356, nav_planner.py:132-138)"""
        cfg = self.config
        brake = bool((desired_speed < cfg.brake_speed) or ((speed / desired_speed) > cfg.brake_ratio))
        delta = np.clip(desired_speed - speed, 0.0, cfg.clip_delta)
        self.speed_window.append(delta)
        kp, ki, kd = self.speed_gains
        throttle = kp * delta + ki * np.mean(self.speed_window) + kd * (self.speed_window[-1] - self.speed_window[-2])
        throttle = np.clip(throttle, 0.0, cfg.clip_throttle)
        throttle = throttle if not brake else 0.0
        self.turn_window = (self.turn_window + [heading_error])[-self.lateral_n:]
        derivative = 0.0 if len(self.turn_window) == 1 else self.turn_window[-1] - self.turn_window[-2]
        kp, ki, kd = self.lateral_gains
        steer = np.clip(kp * heading_error + kd * derivative + ki * np.mean(self.turn_window), -1.0, 1.0).item()
        return round(float(np.clip(steer, -1.0, 1.0)), 3), float(throttle), brake
