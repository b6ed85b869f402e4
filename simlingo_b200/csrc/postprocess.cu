// Post-processing behind the hot path: what the reference does on the host, per sample, with the
// predicted route / speed waypoints before its PID controllers run —
//   DrivingModel.equal_spacing_route      simlingo_training/models/driving.py:330-342   (np.interp at 1 m arc-length steps)
//   LingoAgent.interpolate_waypoints      team_code/agent_simlingo.py:960-1003          (scipy PchipInterpolator at 0.1 m steps)
//   LingoAgent.control_pid                team_code/agent_simlingo.py:944-946           (desired speed from waypoint spacing)
//   LateralPIDController.step             team_code/nav_planner.py:113-128              (look-ahead index, heading error)
// Only ONE of the 0.1 m samples is ever consumed (the look-ahead point), so the kernel evaluates the PCHIP cubic at that
// single arc length from the two local Fritsch-Carlson slopes instead of building the whole spline.
// Arithmetic follows numpy's types: float32 segment lengths / running sums (round-to-nearest, no FMA contraction),
// everything after the hand-over to scipy / np.interp in float64.  One thread per sample: the work is ~20 points.
#include "common.cuh"
#include "../../include/simlingo_b200.h"

namespace {

constexpr int kMaxPoints = 64;  // origin + predicted points

struct Polyline {
  double x[kMaxPoints], px[kMaxPoints], py[kMaxPoints];  // arc length (strictly increasing), coordinates
  int n;                                                  // points including the origin
};

// np.concatenate((0, pts)) ; np.linalg.norm(diff, axis=1) ; np.cumsum ; += arange * 1e-4   (float32 in, float32 arc length)
__device__ void build_polyline(const float* __restrict__ pts, int n_pts, Polyline& p) {
  p.n = n_pts + 1;
  p.x[0] = 0.0;
  p.px[0] = 0.0;
  p.py[0] = 0.0;
  float run = 0.0f, lx = 0.0f, ly = 0.0f;
  for (int k = 1; k <= n_pts; ++k) {
    const float cx = pts[2 * (k - 1)], cy = pts[2 * (k - 1) + 1];
    const float dx = __fsub_rn(cx, lx), dy = __fsub_rn(cy, ly);
    run = __fadd_rn(run, __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))));
    p.x[k] = (double)(float)__dadd_rn((double)run, __dmul_rn((double)k, 1e-4));
    p.px[k] = (double)cx;
    p.py[k] = (double)cy;
    lx = cx;
    ly = cy;
  }
}

// largest i with x[i] <= q, clamped to a valid interval start
__device__ int interval_of(const Polyline& p, double q) {
  int i = 0;
  while (i + 2 < p.n && p.x[i + 1] <= q) ++i;
  return i;
}

__device__ double sgn(double v) { return (v > 0.0) - (v < 0.0); }

// PchipInterpolator._edge_case: one-sided three-point estimate, shape preserving
__device__ double pchip_end(double h0, double h1, double m0, double m1) {
  const double d = __ddiv_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(2.0, h0), h1), m0), __dmul_rn(h0, m1)), __dadd_rn(h0, h1));
  if (sgn(d) != sgn(m0)) return 0.0;
  if (sgn(m0) != sgn(m1) && fabs(d) > __dmul_rn(3.0, fabs(m0))) return __dmul_rn(3.0, m0);
  return d;
}

// PchipInterpolator._find_derivatives at point k of one coordinate
__device__ double pchip_slope(const Polyline& p, const double* y, int k) {
  const int last = p.n - 1;
  auto h = [&](int i) { return __dsub_rn(p.x[i + 1], p.x[i]); };
  auto m = [&](int i) { return __ddiv_rn(__dsub_rn(y[i + 1], y[i]), h(i)); };
  if (k == 0) return pchip_end(h(0), h(1), m(0), m(1));
  if (k == last) return pchip_end(h(last - 1), h(last - 2), m(last - 1), m(last - 2));
  const double m0 = m(k - 1), m1 = m(k), h0 = h(k - 1), h1 = h(k);
  if (sgn(m0) != sgn(m1) || m0 == 0.0 || m1 == 0.0) return 0.0;
  const double w1 = __dadd_rn(__dmul_rn(2.0, h1), h0), w2 = __dadd_rn(h1, __dmul_rn(2.0, h0));
  const double whmean = __ddiv_rn(__dadd_rn(__ddiv_rn(w1, m0), __ddiv_rn(w2, m1)), __dadd_rn(w1, w2));
  return __ddiv_rn(1.0, whmean);
}

// CubicHermiteSpline coefficients of interval i, evaluated like PPoly (power basis in s = q - x[i])
__device__ double pchip_eval(const Polyline& p, const double* y, int i, double q) {
  const double d0 = pchip_slope(p, y, i), d1 = pchip_slope(p, y, i + 1);
  const double h = __dsub_rn(p.x[i + 1], p.x[i]);
  const double slope = __ddiv_rn(__dsub_rn(y[i + 1], y[i]), h);
  const double t = __ddiv_rn(__dsub_rn(__dadd_rn(d0, d1), __dmul_rn(2.0, slope)), h);
  const double c0 = __ddiv_rn(t, h), c1 = __dsub_rn(__ddiv_rn(__dsub_rn(slope, d0), h), t);
  const double s = __dsub_rn(q, p.x[i]);
  double res = y[i], z = s;
  res = __dadd_rn(res, __dmul_rn(d0, z));
  z = __dmul_rn(z, s);
  res = __dadd_rn(res, __dmul_rn(c1, z));
  z = __dmul_rn(z, s);
  return __dadd_rn(res, __dmul_rn(c0, z));
}

__global__ void control_inputs_kernel(const float* __restrict__ route, const float* __restrict__ speed_wps, const float* __restrict__ speed,
                                      int batch, int n_route, int n_wps, slb_control_params prm, double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  // desired speed: distance covered between two speed waypoints, doubled (agent_simlingo.py:944-946; float32)
  const float* w = speed_wps + (size_t)b * n_wps * 2;
  const float wx = __fsub_rn(w[2 * prm.wp_a], w[2 * prm.wp_b]), wy = __fsub_rn(w[2 * prm.wp_a + 1], w[2 * prm.wp_b + 1]);
  const float desired = __fmul_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(wx, wx), __fmul_rn(wy, wy))), 2.0f);

  Polyline p;
  build_polyline(route + (size_t)b * n_route * 2, n_route, p);
  // np.arange(0.1, arc[-1], 0.1): ceil((stop - start) / step) samples, sample j at start + j * step
  const double len = __ddiv_rn(__dsub_rn(p.x[p.n - 1], prm.sample_step), prm.sample_step);
  const int n_interp = len > 0.0 ? (int)ceil(len) : 0;
  // look-ahead index (nav_planner.py:113-121, inference_mode False): clip(scale * km/h + offset, lo, hi), float32
  const float kmh = __fmul_rn(speed[b], 3.6f);
  float look = __fadd_rn(__fmul_rn(prm.lookahead_scale, kmh), prm.lookahead_offset);
  look = fminf(fmaxf(look, prm.lookahead_min), prm.lookahead_max);
  const int rows = n_interp > 0 ? n_interp : 1;  // an empty sampling falls back to the last waypoint (agent_simlingo.py:999-1001)
  const int n_look = (float)(rows - 1) < look ? rows - 1 : (int)look;
  double ax, ay;
  if (n_interp == 0) {
    ax = p.px[p.n - 1];
    ay = p.py[p.n - 1];
  } else {
    const double q = __dadd_rn(prm.sample_step, __dmul_rn((double)n_look, prm.sample_step));
    const int i = interval_of(p, q);
    ax = pchip_eval(p, p.px, i, q);
    ay = pchip_eval(p, p.py, i, q);
  }
  // heading error: atan2 wrapped to (-pi, pi], scaled by 180 / pi / 90 (nav_planner.py:123-130)
  const double two_pi = 6.283185307179586, pi = 3.141592653589793;
  double yaw = atan2(ay, ax);
  if (yaw < 0.0) yaw = __dadd_rn(yaw, two_pi);
  if (yaw >= two_pi) yaw = 0.0;
  if (!(yaw < pi)) yaw = __dsub_rn(yaw, two_pi);
  const double heading = __ddiv_rn(__ddiv_rn(__dmul_rn(yaw, 180.0), pi), 90.0);
  double* o = out + (size_t)b * 8;
  o[0] = (double)desired;
  o[1] = heading;
  o[2] = ax;
  o[3] = ay;
  o[4] = (double)rows;
  o[5] = (double)n_look;
  o[6] = (double)speed[b];
  o[7] = 0.0;
}

__global__ void equal_spacing_kernel(const float* __restrict__ route, int batch, int n_route, int n_out, double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  Polyline p;
  build_polyline(route + (size_t)b * n_route * 2, n_route, p);
  const int last = p.n - 1;
  double* o = out + (size_t)b * n_out * 2;
  int j = 0;
  for (int g = 0; g < n_out; ++g) {  // np.interp(arange(n_out), arc, coord): clamped at both ends, linear inside
    const double q = (double)g;
    double rx, ry;
    if (q >= p.x[last]) {
      rx = p.px[last];
      ry = p.py[last];
    } else {
      while (j + 1 < last && p.x[j + 1] <= q) ++j;
      if (p.x[j] == q) {
        rx = p.px[j];
        ry = p.py[j];
      } else {
        const double h = __dsub_rn(p.x[j + 1], p.x[j]), s = __dsub_rn(q, p.x[j]);
        rx = __dadd_rn(__dmul_rn(__ddiv_rn(__dsub_rn(p.px[j + 1], p.px[j]), h), s), p.px[j]);
        ry = __dadd_rn(__dmul_rn(__ddiv_rn(__dsub_rn(p.py[j + 1], p.py[j]), h), s), p.py[j]);
      }
    }
    o[2 * g] = rx;
    o[2 * g + 1] = ry;
  }
}

}  // namespace

extern "C" int slb_control_inputs(const float* route, const float* speed_wps, const float* speed, int batch, int n_route, int n_wps,
                                  const slb_control_params* params, double* out, void* stream) {
  SLB_CHECK_ARG(route && speed_wps && speed && params && out && batch > 0, "control_inputs: bad args");
  SLB_CHECK_ARG(n_route >= 2 && n_route < kMaxPoints, "control_inputs: route needs 2..63 points");
  SLB_CHECK_ARG(n_wps > 0 && params->wp_a >= 0 && params->wp_a < n_wps && params->wp_b >= 0 && params->wp_b < n_wps,
                "control_inputs: speed-waypoint indices out of range");
  SLB_CHECK_ARG(params->sample_step > 0.0 && params->lookahead_min >= 0.0f && params->lookahead_max >= params->lookahead_min,
                "control_inputs: bad sampling step / look-ahead range");
  control_inputs_kernel<<<(batch + 31) / 32, 32, 0, (cudaStream_t)stream>>>(route, speed_wps, speed, batch, n_route, n_wps, *params, out);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}

extern "C" int slb_equal_spacing_route(const float* route, int batch, int n_route, int n_out, double* out, void* stream) {
  SLB_CHECK_ARG(route && out && batch > 0 && n_out > 0, "equal_spacing_route: bad args");
  SLB_CHECK_ARG(n_route >= 1 && n_route < kMaxPoints, "equal_spacing_route: route needs 1..63 points");
  equal_spacing_kernel<<<(batch + 31) / 32, 32, 0, (cudaStream_t)stream>>>(route, batch, n_route, n_out, out);
  SLB_LAUNCH_CHECK();
  return SLB_OK;
}
