/*
 * oracle_tick.c - the float64 oracle of the batched tick in plain C (TEST INFRASTRUCTURE ONLY).
 *
 * Same algorithm, same order of float64 operations as oracle/physics.py, oracle/geometry.py and
 * oracle/scenario.py (which cite the reference lines); here so that full-size configurations
 * (4096 x 64 and larger) can be checked in well under a second per step and so that bench.py
 * can report a compiled multi-threaded CPU figure beside the reference-style Python loop.
 * Never linked or loaded by tactics2d_b200.  Held to the NumPy oracle (and through it to the
 * reference's golden vectors) by tests/test_oracle_c.py.
 *
 * Reference (paths relative to the reference root):
 *   kinematics   tactics2d/physics/single_track_kinematics.py:178-198,126-176
 *   dynamics     tactics2d/physics/single_track_dynamics.py:231-251,140-229
 *   point mass   tactics2d/physics/point_mass.py:209-232,83-175,177-207
 *   pose         tactics2d/participant/element/vehicle.py:263-281, pedestrian.py:138-149
 *   collisions   tactics2d/traffic/event_detection/collision.py:18-25,37-43
 *   out of bound tactics2d/traffic/event_detection/out_bound.py:37-48
 *
 * build: gcc -O2 -std=c11 -ffp-contract=off -fopenmp -shared -fPIC -o liboracle_tick.so oracle_tick.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

typedef struct {
  float half_len, half_wid, radius, lf, lr, steer_lo, steer_hi, speed_lo, speed_hi, accel_lo, accel_hi;
  float mass, mass_height, mu, I_z, cf, cr;
  int32_t model, shape;
} params_t; /* = t2d_type_params */

enum { KINEMATICS = 0, DYNAMICS = 1, PM_NEWTON = 2, PM_EULER = 3, STATIC_MODEL = 4 };
enum { OBB = 0, CIRCLE = 1, NOSHAPE = 2 };
#define INACTIVE 255
#define TWO_PI 6.283185307179586476925286766559
#define G_ACC 9.81

static double clipd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); } /* np.clip */
static double pymod(double a, double b) { double r = fmod(a, b); if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b; return r; } /* np.mod */

static void kinematics(const params_t* p, double* x, double* y, double* phi, double* v, double* vx, double* vy,
                       double accel, double delta, int interval, int delta_t) {
  accel = clipd(accel, p->accel_lo, p->accel_hi);
  delta = clipd(delta, p->steer_lo, p->steer_hi);
  const double lf = p->lf, lr = p->lr, L = lf + lr;
  const double beta = atan(lr / L * tan(delta));
  const int n_steps = interval / delta_t, rem = interval % delta_t;
  for (int i = 0; i < n_steps + (rem > 0); ++i) {
    const double h = i < n_steps ? (double)delta_t / 1000 : (double)rem / 1000;
    const double dx = *v * cos(*phi + beta), dy = *v * sin(*phi + beta);
    const double dphi = *v / L * tan(delta) * cos(beta);
    *x += dx * h; *y += dy * h; *phi += dphi * h; *v += accel * h;
    *v = clipd(*v, p->speed_lo, p->speed_hi);
  }
  *vx = *v * cos(*phi); *vy = *v * sin(*phi);
  *phi = pymod(*phi, TWO_PI);
}

static void dynamics(const params_t* p, double* x, double* y, double* phi, double* v, double* vx, double* vy,
                     double accel, double delta, int interval, int delta_t) {
  accel = clipd(accel, p->accel_lo, p->accel_hi);
  delta = clipd(delta, p->steer_lo, p->steer_hi);
  const double lf = p->lf, lr = p->lr, L = lf + lr, dt = (double)delta_t / 1000;
  const double mass = p->mass, mh = p->mass_height, mu = p->mu, Iz = p->I_z, cf = p->cf, cr = p->cr;
  const double ff = (G_ACC * lr - accel * mh) / L, fr = (G_ACC * lf + accel * mh) / L;
  const double a1 = lf * cf * ff, a2 = lr * cr * fr;
  const double b1 = lf * lf * cf * ff, b2 = lr * lr * cr * fr;
  const double c1 = cf * ff, c2 = cr * fr;
  double d_phi = *v / L * tan(delta);
  double beta = atan(lr / lf * tan(delta));
  const int n_steps = interval / delta_t;
  for (int i = 0; i < n_steps; ++i) {
    const double dx = *v * cos(*phi + beta), dy = *v * sin(*phi + beta);
    const double vs = fabs(*v) > 1e-6 ? *v : (*v >= 0 ? 1e-6 : -1e-6);
    double d_beta;
    if (fabs(*v) >= 0.1) {
      const double dd_phi = mu * mass / Iz * (a1 * delta + (a2 - a1) * beta - (b1 + b2) * d_phi / vs);
      d_beta = mu / vs * (c1 * delta - (c2 + c1) * beta + (a2 - a1) * d_phi / vs) - d_phi;
      d_phi += dd_phi * dt;
    } else {
      const double q = 1 + tan(delta) * lr / L;
      d_beta = lr / (q * q) / L / (cos(delta) * cos(delta)) * delta;
      d_phi += *v * cos(beta) / L * tan(delta) * dt;
    }
    *x += dx * dt; *y += dy * dt; *v += accel * dt; *phi += d_phi * dt; beta += d_beta * dt;
    *v = clipd(*v, p->speed_lo, p->speed_hi);
  }
  *phi = pymod(*phi, TWO_PI);
  *vx = *v * cos(*phi); *vy = *v * sin(*phi);
}

static double newton_t1(double ax, double ay, double vx, double vy, double lim, double sign, double dt) {
  const double a_ = ax * ax + ay * ay, b_ = 2 * (ax * vx + ay * vy), c_ = vx * vx + vy * vy - lim * lim;
  double t1;
  if (fabs(a_) < 1e-12) t1 = fabs(b_) < 1e-12 ? 0.0 : -c_ / b_;
  else { const double disc = fmax(0.0, b_ * b_ - 4 * a_ * c_); t1 = (-b_ + sign * sqrt(disc)) / (2 * a_); }
  return clipd(t1, 0.0, dt);
}

static void pm_newton(const params_t* p, double* x, double* y, double* phi, double* v, double* vx, double* vy,
                      double ax, double ay, int interval) {
  const double dt = (double)interval / 1000;
  const double nvx = *vx + ax * dt, nvy = *vy + ay * dt, nsp = sqrt(nvx * nvx + nvy * nvy);
  if (p->speed_lo <= nsp && nsp <= p->speed_hi) {
    *x = *x + *vx * dt + 0.5 * ax * (dt * dt); *y = *y + *vy * dt + 0.5 * ay * (dt * dt);
    *vx = nvx; *vy = nvy;
  } else {
    const int low = nsp < p->speed_lo;
    const double t1 = newton_t1(ax, ay, *vx, *vy, low ? p->speed_lo : p->speed_hi, low ? -1.0 : 1.0, dt), t2 = dt - t1;
    const double lx = *vx + ax * t1, ly = *vy + ay * t1;
    *x = *x + *vx * t1 + 0.5 * ax * (t1 * t1) + lx * t2; *y = *y + *vy * t1 + 0.5 * ay * (t1 * t1) + ly * t2;
    *vx = lx; *vy = ly;
  }
  *phi = atan2(*vy, *vx);
  *v = sqrt(*vx * *vx + *vy * *vy);
}

static void pm_euler(const params_t* p, double* x, double* y, double* phi, double* v, double* vx, double* vy,
                     double ax, double ay, int interval, int delta_t) {
  const int n_steps = interval / delta_t, rem = interval % delta_t;
  for (int i = 0; i < n_steps + (rem > 0); ++i) {
    const double h = i < n_steps ? (double)delta_t / 1000 : (double)rem / 1000;
    *vx += ax * h; *vy += ay * h;
    const double sp = sqrt(*vx * *vx + *vy * *vy), cl = clipd(sp, p->speed_lo, p->speed_hi);
    if (fabs(sp - cl) > 1e-12) { *vx = cl * cos(*phi); *vy = cl * sin(*phi); }
    *x += *vx * h; *y += *vy * h;
    *phi = atan2(*vy, *vx);
  }
  *v = sqrt(*vx * *vx + *vy * *vy);
}

/* physics of all N*M participants: fp32 in, float64 out */
void oracle_physics(int n_total, const params_t* table, int n_types, const float* x, const float* y, const float* h,
                    const float* v, const float* vx, const float* vy, const uint8_t* type_id, const float* action,
                    int interval, int delta_t, int steer_first, double* ox, double* oy, double* oh, double* ov,
                    double* ovx, double* ovy) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n_total; ++i) {
    double X = x[i], Y = y[i], H = h[i], V = v[i], VX = vx[i], VY = vy[i];
    const int t = type_id[i];
    if (t < n_types) {
      const params_t* p = &table[t];
      double a0 = action[2 * i], a1 = action[2 * i + 1];
      if (steer_first && p->model <= DYNAMICS) { const double s = a0; a0 = a1; a1 = s; }
      if (p->model == KINEMATICS) kinematics(p, &X, &Y, &H, &V, &VX, &VY, a0, a1, interval, delta_t);
      else if (p->model == DYNAMICS) dynamics(p, &X, &Y, &H, &V, &VX, &VY, a0, a1, interval, delta_t);
      else if (p->model == PM_NEWTON) pm_newton(p, &X, &Y, &H, &V, &VX, &VY, a0, a1, interval);
      else if (p->model == PM_EULER) pm_euler(p, &X, &Y, &H, &V, &VX, &VY, a0, a1, interval, delta_t);
    }
    ox[i] = X; oy[i] = Y; oh[i] = H; ov[i] = V; ovx[i] = VX; ovy[i] = VY;
  }
}

/* ---- closed-set predicates = oracle/geometry.py ---- */
typedef struct { double x, y, c, s, l, w; int circle; int solid; } pose_t;

static int obb_obb(const pose_t* a, const pose_t* b) {
  const double tx = b->x - a->x, ty = b->y - a->y;
  const double c = a->c * b->c + a->s * b->s, s = a->c * b->s - a->s * b->c;
  const double ac = fabs(c), as = fabs(s);
  return fabs(tx * a->c + ty * a->s) <= a->l + (b->l * ac + b->w * as) && fabs(ty * a->c - tx * a->s) <= a->w + (b->l * as + b->w * ac) &&
         fabs(tx * b->c + ty * b->s) <= b->l + (a->l * ac + a->w * as) && fabs(ty * b->c - tx * b->s) <= b->w + (a->l * as + a->w * ac);
}
static int obb_circle(const pose_t* a, double xc, double yc, double r) {
  const double tx = xc - a->x, ty = yc - a->y;
  const double qx = fabs(tx * a->c + ty * a->s) - a->l, qy = fabs(ty * a->c - tx * a->s) - a->w;
  const double dx = fmax(qx, 0.0), dy = fmax(qy, 0.0);
  return dx * dx + dy * dy <= r * r;
}
static int pair(const pose_t* a, const pose_t* b) {
  if (!a->circle && !b->circle) return obb_obb(a, b);
  if (!a->circle) return obb_circle(a, b->x, b->y, b->l);
  if (!b->circle) return obb_circle(b, a->x, a->y, a->l);
  const double tx = b->x - a->x, ty = b->y - a->y;
  return tx * tx + ty * ty <= (a->l + b->l) * (a->l + b->l);
}
static int seg_hit(const pose_t* a, const float* sg) {
  const double x1 = sg[0], y1 = sg[1], x2 = sg[2], y2 = sg[3];
  if (a->circle) {
    const double dx = x2 - x1, dy = y2 - y1, ux = a->x - x1, uy = a->y - y1, dd = dx * dx + dy * dy;
    double t = dd > 0 ? (ux * dx + uy * dy) / dd : 0.0;
    t = clipd(t, 0.0, 1.0);
    const double ex = ux - t * dx, ey = uy - t * dy;
    return ex * ex + ey * ey <= a->l * a->l;
  }
  const double ux = x1 - a->x, uy = y1 - a->y, vx = x2 - a->x, vy = y2 - a->y;
  const double p1x = ux * a->c + uy * a->s, p1y = uy * a->c - ux * a->s, p2x = vx * a->c + vy * a->s, p2y = vy * a->c - vx * a->s;
  const double dx = p2x - p1x, dy = p2y - p1y;
  return fmax(p1x, p2x) >= -a->l && fmin(p1x, p2x) <= a->l && fmax(p1y, p2y) >= -a->w && fmin(p1y, p2y) <= a->w &&
         fabs(p1x * dy - p1y * dx) <= a->l * fabs(dy) + a->w * fabs(dx);
}

/* events on given float64 poses; first hit in list order (collision.py:18-25,37-43) */
void oracle_events(int N, int M, const params_t* table, int n_types, const double* x, const double* y, const double* h,
                   const uint8_t* type_id, const float* seg, int S, const float* bounds, uint8_t* flags,
                   int16_t* hit_index, int16_t* hit_segment) {
#pragma omp parallel for schedule(dynamic, 16)
  for (int n = 0; n < N; ++n) {
    pose_t ps[256];
    for (int m = 0; m < M; ++m) {
      const size_t i = (size_t)n * M + m;
      const int t = type_id[i];
      pose_t* p = &ps[m];
      p->solid = t < n_types && table[t].shape != NOSHAPE;
      if (!p->solid) continue;
      p->x = x[i]; p->y = y[i]; p->c = cos(h[i]); p->s = sin(h[i]);
      p->circle = table[t].shape == CIRCLE;
      p->l = p->circle ? table[t].radius : table[t].half_len;
      p->w = p->circle ? 0.0 : table[t].half_wid;
    }
    for (int m = 0; m < M; ++m) {
      const size_t i = (size_t)n * M + m;
      uint8_t f = 0; int hi = -1, hs = -1;
      const pose_t* a = &ps[m];
      if (a->solid) {
        for (int j = 0; j < M; ++j) { if (j == m || !ps[j].solid) continue; if (pair(a, &ps[j])) { hi = j; f |= 1; break; } }
        for (int k = 0; k < S; ++k) if (seg_hit(a, seg + 4 * k)) { hs = k; f |= 2; break; }
        if (bounds) {
          const double ex = a->circle ? a->l : a->l * fabs(a->c) + a->w * fabs(a->s);
          const double ey = a->circle ? a->l : a->l * fabs(a->s) + a->w * fabs(a->c);
          if (a->x - ex < bounds[0] || a->x + ex > bounds[1] || a->y - ey < bounds[2] || a->y + ey > bounds[3]) f |= 4;
        }
      }
      flags[i] = f; hit_index[i] = (int16_t)hi; hit_segment[i] = (int16_t)hs;
    }
  }
}

int oracle_abi(void) { return 1; }
