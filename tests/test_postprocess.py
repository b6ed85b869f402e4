"""Post-processing behind the hot path: equal-spacing route, PCHIP aim point, PID inputs.
CPU part: the numpy oracle against the golden outputs of the reference's own code (tests/golden/postprocess.npz).
GPU part: ``slb_control_inputs`` / ``slb_equal_spacing_route`` through the C ABI against both."""
import os

import numpy as np
import pytest
import torch

from oracle import postprocess as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "postprocess.npz"))
EPISODES = ("ep0", "ep1")


def test_oracle_reproduces_reference_interpolation():
    for ep in EPISODES:
        routes = GOLD[f"{ep}_route"]
        for t, r in enumerate(routes):
            interp = O.interpolate_waypoints(r)
            assert interp.shape[0] == GOLD[f"{ep}_interp_count"][t], (ep, t)
            np.testing.assert_allclose(interp[min(24, len(interp) - 1)], GOLD[f"{ep}_interp_at24"][t], rtol=0, atol=1e-12)
            np.testing.assert_allclose(O.equal_spacing_route(r), GOLD[f"{ep}_equal_spacing"][t], rtol=0, atol=1e-12)
        np.testing.assert_allclose(O.interpolate_waypoints(routes[0]), GOLD[f"{ep}_interp_full_t0"], rtol=0, atol=1e-12)


def test_oracle_reproduces_reference_control_sequence():
    """closed loop: the PID windows carry state, so every tick has to match for the next one to"""
    for ep in EPISODES:
        state = O.PIDState()
        for t, (r, w, s) in enumerate(zip(GOLD[f"{ep}_route"], GOLD[f"{ep}_speed_wps"], GOLD[f"{ep}_speed"])):
            steer, throttle, brake = O.control_pid(state, r, w, s)
            ref = GOLD[f"{ep}_controls"][t]
            assert (steer, throttle, float(brake)) == (ref[0], ref[1], ref[2]), (ep, t, steer, throttle, brake, ref)


def test_host_pid_reproduces_reference_control_sequence():
    """host half of the product (``ControlPID.step``: the two PID windows) fed with the oracle's geometry"""
    from simlingo_b200.postprocess import ControlPID
    for ep in EPISODES:
        pid = ControlPID()
        for t, (r, w, s) in enumerate(zip(GOLD[f"{ep}_route"], GOLD[f"{ep}_speed_wps"], GOLD[f"{ep}_speed"])):
            desired, heading, _ = O.control_inputs(r, w, s)
            steer, throttle, brake = pid.step(desired, heading, np.float32(s))
            ref = GOLD[f"{ep}_controls"][t]
            assert (steer, throttle, float(brake)) == (ref[0], ref[1], ref[2]), (ep, t)


def test_postprocess_rejects_host_tensors():
    from simlingo_b200 import postprocess
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        postprocess.equal_spacing_route(torch.zeros(1, 20, 2))


def _random_routes(n, seed, n_pts=20):
    g = np.random.default_rng(seed)
    ang = np.cumsum(g.normal(0, 0.08, (n, n_pts)), 1) + g.uniform(-np.pi, np.pi, (n, 1))
    step = np.abs(g.normal(1.0, 0.5, (n, n_pts))) * g.choice([0.0, 0.02, 0.3, 1.0, 2.5], (n, 1))
    step[g.random((n, n_pts)) < 0.05] = 0.0   # repeated points
    return np.cumsum(np.stack([step * np.cos(ang), step * np.sin(ang)], -1), 1).astype(np.float32)


def test_oracle_pchip_matches_scipy_on_random_routes():
    """the restated Fritsch-Carlson / end-rule / power-form evaluation against scipy's PchipInterpolator itself (the
    third-party piece the reference calls, agent_simlingo.py:986), on 300 seeded routes incl. repeated points, flat
    segments, reversals and 2- / 5-point polylines"""
    from scipy.interpolate import PchipInterpolator
    worst = 0.0
    for n_pts, n, seed in ((20, 200, 10), (5, 50, 11), (2, 50, 12)):
        for r in _random_routes(n, seed, n_pts):
            poly, arc = O.arc_length(r)
            q = np.arange(0.1, arc[-1], 0.1)
            if q.shape[0] == 0:
                assert np.array_equal(O.interpolate_waypoints(r), poly[None, -1])
                continue
            want = PchipInterpolator(arc, poly, axis=0)(q)
            got = O.interpolate_waypoints(r)
            assert got.shape == want.shape
            worst = max(worst, float(np.abs(got - want).max()))
    assert worst <= 1e-12, worst


def test_oracle_geometry_properties():
    """size-independent properties: a straight, already equally spaced route is a fixed point of the resampling; the
    aim point of a straight route lies on it at the look-ahead arc length; rotating the route rotates the heading error"""
    straight = np.stack([np.arange(1, 21, dtype=np.float32), np.zeros(20, np.float32)], 1)
    eq = O.equal_spacing_route(straight)
    np.testing.assert_allclose(eq[1:], straight[:19] * (1 - 1e-4), atol=3e-3)      # the 1e-4 * k offset shifts samples by < 2 mm
    assert np.array_equal(eq[0], [0.0, 0.0])
    wps = np.cumsum(np.full((10, 2), [1.0, 0.0], np.float32), 0)
    desired, heading, aim = O.control_inputs(straight, wps, np.float32(3.0))
    assert float(desired) == 4.0 and abs(heading) < 1e-12 and abs(aim[0] - 2.5) < 1e-3 and abs(aim[1]) < 1e-12
    g = np.random.default_rng(3)
    ang = np.cumsum(np.full(20, 0.03))
    base = np.cumsum(np.stack([np.cos(ang), np.sin(ang)], 1), 0).astype(np.float32)   # gentle left curve, 1 m steps
    _, h0, aim0 = O.control_inputs(base, wps, np.float32(3.0))
    for theta in g.uniform(-0.5, 0.5, 8):
        c, s_ = np.cos(theta), np.sin(theta)
        rot = (base.astype(np.float64) @ np.array([[c, s_], [-s_, c]])).astype(np.float32)
        _, h, aim = O.control_inputs(rot, wps, np.float32(3.0))
        assert abs(np.hypot(*aim) - np.hypot(*aim0)) < 1e-4
        d = (h - h0) * np.pi / 2 - theta                                            # heading error is yaw / (pi / 2)
        assert abs((d + np.pi) % (2 * np.pi) - np.pi) < 1e-4, (theta, h, h0)


# ---------------------------------------------------------------------------------------------- GPU (through the C ABI)
@pytest.mark.gpu
def test_control_inputs_kernel_matches_reference_golden():
    from simlingo_b200 import postprocess
    for ep in EPISODES:
        r, w, s = (torch.from_numpy(GOLD[f"{ep}_{k}"]).cuda() for k in ("route", "speed_wps", "speed"))
        out = postprocess.control_inputs(r, w, s).cpu().numpy()
        assert np.array_equal(out[:, 4], GOLD[f"{ep}_interp_count"])
        assert np.array_equal(out[:, 6], GOLD[f"{ep}_speed"].astype(np.float64))
        for t in range(len(out)):
            desired, heading, aim = O.control_inputs(GOLD[f"{ep}_route"][t], GOLD[f"{ep}_speed_wps"][t], GOLD[f"{ep}_speed"][t])
            assert out[t, 0] == pytest.approx(float(desired), rel=2e-7, abs=0)       # float32: numpy's dot may or may not fuse
            np.testing.assert_allclose(out[t, 2:4], aim, rtol=0, atol=1e-12)          # float64 PCHIP sample
            assert abs(out[t, 1] - heading) <= 1e-13                                   # float64 atan2
            if out[t, 5] == 24 and out[t, 4] > 24:
                np.testing.assert_allclose(out[t, 2:4], GOLD[f"{ep}_interp_at24"][t], rtol=0, atol=1e-12)


@pytest.mark.gpu
def test_control_pid_closed_loop_matches_reference():
    """the agent-facing call, tick by tick: kernel + 64-byte read-back + host PID vs the reference's control sequence"""
    from simlingo_b200.postprocess import ControlPID
    for ep in EPISODES:
        pid = ControlPID()
        for t, (r, w, s) in enumerate(zip(GOLD[f"{ep}_route"], GOLD[f"{ep}_speed_wps"], GOLD[f"{ep}_speed"])):
            steer, throttle, brake = pid.control_pid(torch.from_numpy(r)[None].cuda(), torch.tensor([s]), torch.from_numpy(w)[None].cuda())
            ref = GOLD[f"{ep}_controls"][t]
            assert abs(steer - ref[0]) <= 1e-3 and abs(throttle - ref[1]) <= 1e-5 and float(brake) == ref[2], (ep, t, steer, throttle, brake, ref)


@pytest.mark.gpu
def test_equal_spacing_and_aim_point_on_random_routes():
    """1000 seeded routes incl. degenerate ones (all at the origin, centimetre-long, repeated points) and short
    polylines (2 and 5 points) against the numpy oracle; the predict_step golden against the reference's own output"""
    from simlingo_b200 import postprocess
    for ep in EPISODES:
        got = postprocess.equal_spacing_route(torch.from_numpy(GOLD[f"{ep}_route"]).cuda()).cpu().numpy()
        np.testing.assert_allclose(got, GOLD[f"{ep}_equal_spacing"], rtol=0, atol=1e-12)
    for n_pts, n, seed in ((20, 1000, 0), (5, 64, 1), (2, 64, 2)):
        routes = _random_routes(n, seed, n_pts)
        g = np.random.default_rng(seed + 100)
        wps = np.cumsum(np.abs(g.normal(0.5, 0.5, (n, 10, 2))), 1).astype(np.float32)
        speed = g.uniform(0, 30, n).astype(np.float32)
        eq = postprocess.equal_spacing_route(torch.from_numpy(routes).cuda()).cpu().numpy()
        out = postprocess.control_inputs(torch.from_numpy(routes).cuda(), torch.from_numpy(wps).cuda(), torch.from_numpy(speed).cuda()).cpu().numpy()
        for i in range(n):
            np.testing.assert_allclose(eq[i], O.equal_spacing_route(routes[i]), rtol=0, atol=1e-12)
            desired, heading, aim = O.control_inputs(routes[i], wps[i], speed[i])
            interp = O.interpolate_waypoints(routes[i])
            assert out[i, 4] == interp.shape[0] and out[i, 0] == pytest.approx(float(desired), rel=2e-7)
            np.testing.assert_allclose(out[i, 2:4], aim, rtol=0, atol=1e-11)
            if np.hypot(*aim) > 1e-9:   # atan2 of a point at the origin is ill-conditioned; everywhere else 1e-12
                assert abs(out[i, 1] - heading) <= 1e-12
