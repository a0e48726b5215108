/*
 * swarm_oracle.c - CPU restatement of the SwarmACB-isaaclab swarm step.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker for the CUDA path: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product package never
 * links, imports or falls back to it.
 *
 * Parity status: PINNED.  The restatement is checked against single-step fixtures generated
 * from the unmodified reference (tests/golden/gen_golden.py -> the tests/golden npz files, see
 * tests/test_oracle_golden.py).  The un-vendored isaaclab DirectMARLEnv.step hook order is
 * restated from SURVEY.md 3.2 (reference relies on it at ENV:66-68, ENV:1202, HOM:88) and is
 * pinned by no reference-owned test.
 *
 * Plain C99, float32 scalar arithmetic in the reference's operation order; build with
 * -ffp-contract=off so no FMA is formed outside the explicit fmaf() of include/swarm_detmath.h.
 * sin/cos/atan2 come from that header (deterministic, <= ~1.5 ulp), the reference's from SLEEF
 * (<= 1 ulp): last-ulp differences are expected and covered by the stated tolerances (1e-5 m /
 * 1e-5 rad poses, 1e-4 sensors); integer state is exact.
 *
 * Citations: ENV = missions/directional_gate/directional_gate_env.py, SENS = epuck/epuck_sensors.py,
 * BEH = epuck/behavior_modules.py, XOR/HOM/FOR/SHL = the mission env files of the reference.
 */
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/swarm_abi.h"
#include "../include/swarm_detmath.h"

#define N SWARM_N
#define PI_F 3.14159265358979323846f

typedef struct {
  float x[N], y[N], yaw[N];
} Pose;

/* sin/cos/atan2: the deterministic float32 routines shared with the CUDA kernel
 * (include/swarm_detmath.h, <= ~1.5 ulp, bit-identical on CPU and GPU); the reference's come from
 * SLEEF (<= 1 ulp).  tests/test_detmath.py pins them against double-precision libm. */
static inline float cr_sinf(float a) { float sn, cs; swarm_sincosf(a, &sn, &cs); return sn; }
static inline float cr_cosf(float a) { float sn, cs; swarm_sincosf(a, &sn, &cs); return cs; }
static inline float cr_atan2f(float y, float x) { return swarm_atan2f(y, x); }

static inline float signf(float v) { return (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f); }
static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ---- fsm word helpers (include/swarm_abi.h) ---- */
static inline float dec_dir(int c) { return c == 1 ? 1.0f : (c == 2 ? -1.0f : 0.0f); }
static inline int enc_dir(float d) { return d > 0.0f ? 1 : (d < 0.0f ? 2 : 0); }

/* ---- ENV:1048-1078 arena faces, Jacobi sum of all penetrating faces ---- */
static void resolve_walls(const SwarmParams* p, Pose* s) {
  for (int i = 0; i < N; ++i) {
    float tx = 0.0f, ty = 0.0f;
    for (int f = 0; f < 12; ++f) {
      float dx = s->x[i] - p->face_px[f];
      float dy = s->y[i] - p->face_py[f];
      float sd = dx * p->face_nx[f] + dy * p->face_ny[f];
      float pen = p->wall_r_eff - sd;
      pen = pen * ((pen > 0.0f) ? 1.0f : 0.0f);
      tx += pen * p->face_nx[f];
      ty += pen * p->face_ny[f];
    }
    s->x[i] = s->x[i] + tx;
    s->y[i] = s->y[i] + ty;
  }
}

/* ---- ENV:1080-1112 one Jacobi pass over pairs i<j ---- */
static void resolve_robots(const SwarmParams* p, Pose* s) {
  float ax[N], ay[N], bx[N], by[N];
  for (int i = 0; i < N; ++i) ax[i] = ay[i] = bx[i] = by[i] = 0.0f;
  for (int i = 0; i < N; ++i) {
    for (int j = i + 1; j < N; ++j) {
      float dx = s->x[i] - s->x[j];
      float dy = s->y[i] - s->y[j];
      float dist = sqrtf(dx * dx + dy * dy + 1e-8f);
      float ov = p->two_radius - dist;
      if (ov < 0.0f) ov = 0.0f;
      float nx = dx / (dist + 1e-8f);
      float ny = dy / (dist + 1e-8f);
      float px = ov * nx * 0.5f, py = ov * ny * 0.5f;
      ax[i] += px; ay[i] += py;   /* .sum(dim=2): effect on i */
      bx[j] += px; by[j] += py;   /* .sum(dim=1): reverse push on j */
    }
  }
  for (int i = 0; i < N; ++i) {
    s->x[i] = (s->x[i] + ax[i]) - bx[i];
    s->y[i] = (s->y[i] + ay[i]) - by[i];
  }
}

/* ---- ENV:658-705 (DGT, inherited by XOR) and SHL:124-155 ---- */
static void resolve_gate(const SwarmParams* p, Pose* s) {
  if (p->gate_mode == SWARM_GATE_DGT) {
    const float hw = p->gate[0], gs = p->gate[1], top = p->gate[2], r = p->robot_radius;
    for (int i = 0; i < N; ++i) {
      float px = s->x[i], py = s->y[i];
      int in_y = (py > gs) && (py < top);
      float dxl = px - (-hw);
      float pen = r - fabsf(dxl);
      if (pen > 0.0f && in_y && px < 0.0f) {
        float sg = signf(dxl);
        if (sg == 0.0f) sg = -1.0f;
        s->x[i] = -hw + sg * r;
      }
      px = s->x[i];
      float dxr = px - hw;
      pen = r - fabsf(dxr);
      if (pen > 0.0f && in_y && px > 0.0f) {
        float sg = signf(dxr);
        if (sg == 0.0f) sg = 1.0f;
        s->x[i] = hw + sg * r;
      }
    }
  } else if (p->gate_mode == SWARM_GATE_SHL) {
    const float left = p->gate[0], right = p->gate[1], top = p->gate[3], c = p->gate[4];
    const float y_lo = p->gate[5], y_hi = p->gate[6], x_lo = p->gate[7], x_hi = p->gate[8];
    for (int i = 0; i < N; ++i) {
      float py = s->y[i];
      int vertical_y = (py > y_lo) && (py < y_hi);
      for (int w = 0; w < 2; ++w) {
        float x0 = w ? right : left;
        float dx = s->x[i] - x0;   /* px is a live view: sees the first wall's write */
        if (fabsf(dx) < c && vertical_y) {
          float sg = signf(dx);
          if (sg == 0.0f) sg = 1.0f;
          s->x[i] = x0 + sg * c;
        }
      }
      float px = s->x[i];
      int horizontal_x = (px > x_lo) && (px < x_hi);
      float dy = py - top;
      if (fabsf(dy) < c && horizontal_x) {
        float sg = signf(dy);
        if (sg == 0.0f) sg = 1.0f;
        s->y[i] = top + sg * c;
      }
    }
  }
}

/* ---- ENV:898-974 swept side test against each internal wall (sequential over walls) ---- */
static void prevent_crossing(const SwarmParams* p, Pose* s, const float* prx, const float* pry) {
  const float eps = 1e-8f;
  for (int w = 0; w < p->n_internal; ++w) {
    const float ax = p->iw_ax[w], ay = p->iw_ay[w], nx = p->iw_nx[w], ny = p->iw_ny[w];
    const float tx = p->iw_tx[w], ty = p->iw_ty[w], lsq = p->iw_len_sq[w];
    for (int i = 0; i < N; ++i) {
      float prev_signed = (prx[i] - ax) * nx + (pry[i] - ay) * ny;
      float curr_signed = (s->x[i] - ax) * nx + (s->y[i] - ay) * ny;
      float denom = prev_signed - curr_signed;
      int ok = fabsf(denom) > eps;
      float sweep_t = ok ? (prev_signed / denom) : 0.0f;
      float ix = prx[i] + (s->x[i] - prx[i]) * sweep_t;
      float iy = pry[i] + (s->y[i] - pry[i]) * sweep_t;
      float wall_u = ((ix - ax) * tx + (iy - ay) * ty) / lsq;
      int crossed = (prev_signed * curr_signed < 0.0f) && (sweep_t >= 0.0f) && (sweep_t <= 1.0f) &&
                    (wall_u >= 0.0f) && (wall_u <= 1.0f);
      if (crossed) {
        float side = signf(prev_signed);
        if (side == 0.0f) side = -signf(curr_signed);
        if (side == 0.0f) side = 1.0f;
        float desired = side * p->crossing_clearance;
        float corr = desired - curr_signed;
        s->x[i] = s->x[i] + corr * nx;
        s->y[i] = s->y[i] + corr * ny;
      }
    }
  }
}

/* ---- ENV:976-1046 capsule push-out; prx == NULL means prev_pos is None ---- */
static void resolve_capsules(const SwarmParams* p, Pose* s, const float* prx, const float* pry) {
  const float eps = 1e-8f;
  for (int w = 0; w < p->n_internal; ++w) {
    const float ax = p->iw_ax[w], ay = p->iw_ay[w], nx = p->iw_nx[w], ny = p->iw_ny[w];
    const float tx = p->iw_tx[w], ty = p->iw_ty[w], lsq = p->iw_len_sq[w];
    for (int i = 0; i < N; ++i) {
      float relx = s->x[i] - ax, rely = s->y[i] - ay;
      float u = (relx * tx + rely * ty) / lsq;
      float uc = clampf(u, 0.0f, 1.0f);
      float cx = ax + uc * tx, cy = ay + uc * ty;
      float dx = s->x[i] - cx, dy = s->y[i] - cy;
      float raw = sqrtf(dx * dx + dy * dy);
      float dist = raw < eps ? eps : raw;
      float curr_signed = relx * nx + rely * ny;
      float side;
      if (prx) {
        side = signf((prx[i] - ax) * nx + (pry[i] - ay) * ny);
        if (side == 0.0f) side = signf(curr_signed);
      } else {
        side = signf(curr_signed);
      }
      if (side == 0.0f) side = 1.0f;
      float sdx = side * nx, sdy = side * ny;
      float rdx = (raw > eps) ? dx / dist : sdx;
      float rdy = (raw > eps) ? dy / dist : sdy;
      int on_span = (u >= 0.0f) && (u <= 1.0f);
      float pdx = on_span ? sdx : rdx, pdy = on_span ? sdy : rdy;
      float pen = p->capsule_clearance - dist;
      if (pen > 0.0f) {
        s->x[i] = s->x[i] + pen * pdx;
        s->y[i] = s->y[i] + pen * pdy;
      }
    }
  }
}

/* ---- ENV:874-896 solver schedule ---- */
static void resolve_collisions(const SwarmParams* p, Pose* s, const float* prx, const float* pry) {
  float bx[N], by[N];
  resolve_walls(p, s);
  if (prx) prevent_crossing(p, s, prx, pry);
  resolve_capsules(p, s, prx, pry);
  resolve_gate(p, s);
  for (int it = 0; it < p->solver_iterations; ++it) {
    memcpy(bx, s->x, sizeof bx);
    memcpy(by, s->y, sizeof by);
    resolve_robots(p, s);
    resolve_walls(p, s);
    prevent_crossing(p, s, bx, by);
    resolve_capsules(p, s, bx, by);
    resolve_gate(p, s);
  }
  resolve_walls(p, s);
  if (prx) prevent_crossing(p, s, prx, pry);
  resolve_capsules(p, s, prx, pry);
  resolve_gate(p, s);
}

/* ---- ground colour: ENV:707-750, XOR:119-124, HOM:81-85, FOR:119-125, SHL:116-122 ---- */
static int in_circle(float x, float y, float cx, float cy, float rsq) {
  float dx = x - cx, dy = y - cy;
  return (dx * dx + dy * dy) <= rsq;
}

static float ground_color(const SwarmParams* p, float x, float y) {
  const float* z = p->zone;
  float c = 0.5f;
  switch (p->mission) {
    case SWARM_DGT: {
      if (fabsf(x) < z[0] && y > z[1] && y < z[2]) c = 1.0f;
      if (fabsf(x) < z[3] && y >= z[2] && y < z[4]) c = 0.0f;
      break;
    }
    case SWARM_XOR:
      if (in_circle(x, y, z[0], z[1], z[4]) || in_circle(x, y, z[2], z[3], z[4])) c = 0.0f;
      break;
    case SWARM_HOM:
      if (in_circle(x, y, z[0], z[1], z[4])) c = 0.0f;
      break;
    case SWARM_FOR:
      if (in_circle(x, y, z[0], z[1], z[4]) || in_circle(x, y, z[2], z[3], z[4])) c = 0.0f;
      if (y <= z[6]) c = 1.0f;
      break;
    case SWARM_SHL:
      if (in_circle(x, y, z[0], z[1], z[4]) || in_circle(x, y, z[2], z[3], z[4])) c = 0.0f;
      if (x >= z[7] && x <= z[8] && y >= z[9] && y <= z[10]) c = 1.0f;
      break;
  }
  return c;
}

/* ---- SENS:545-586 ---- */
static void critic_state(const SwarmParams* p, const Pose* s, float* out /* (N,5) */) {
  for (int i = 0; i < N; ++i) {
    float rx = s->x[i], ry = s->y[i];
    float norm = sqrtf(rx * rx + ry * ry);
    if (norm < 1e-6f) norm = 1e-6f;
    float rho = clampf(norm / p->critic_radius, 0.0f, 1.0f);
    float hx = rx / norm, hy = ry / norm;
    float ca = hx * 0.0f + hy * 1.0f;
    float sa = hx * 1.0f - hy * 0.0f;
    float cy = cr_cosf(s->yaw[i]), sy = cr_sinf(s->yaw[i]);
    float cb = cy * hx + sy * hy;
    float sb = hx * sy - hy * cy;
    float* o = out + i * 5;
    o[0] = rho; o[1] = ca; o[2] = sa; o[3] = cb; o[4] = sb;
  }
}

/* ---- BEH:50-90 ---- */
static void wheels_from_vector(float dx, float dy, float ms, float* l, float* r) {
  int near_zero = (fabsf(dx) < 1e-5f) && (fabsf(dy) < 1e-5f);
  float angle = cr_atan2f(dy, dx);
  if (angle < 0.0f) angle = angle + 2.0f * PI_F;
  float ca = cr_cosf(angle);
  int front = angle < PI_F;
  float left = front ? ca : 1.0f, right = front ? 1.0f : ca;
  float mv = fmaxf(fabsf(left), fabsf(right));
  if (mv < 1e-5f) mv = 1e-5f;
  float scale = ms / mv;
  left = left * scale; right = right * scale;
  if (near_zero) { left = 0.0f; right = 0.0f; }
  *l = left; *r = right;
}

static void steer(float rx, float ry, float ms, float* l, float* r) {
  float mag = sqrtf(rx * rx + ry * ry);
  if (mag < 0.1f) { rx = 1.0f; ry = 0.0f; }   /* forward fallback */
  wheels_from_vector(rx, ry, ms, l, r);
}

/* BEH:245-251 */
static int obstacle_in_front(const SwarmParams* p, float pv, float pa) {
  return (pv >= p->prox_threshold) && (fabsf(pa) <= (float)(3.14159265358979323846 * 0.5));
}

/* ---- BEH:177-241 dispatch for one robot; fsm word updated in place ---- */
static void dispatch_robot(const SwarmParams* p, int64_t id, const float c[6], float prev_l, float prev_r,
                           const int32_t dur[3], int32_t* fsm, float* out_l, float* out_r) {
  const float ms = p->max_wheel_speed;
  const float pv = c[0], pa = c[1], lv = c[2], la = c[3], rabx = c[4], raby = c[5];
  float l = 0.0f, r = 0.0f;
  int32_t w = *fsm;
  if (id == 1) { /* BEH:266-341 */
    int state = w & 1, steps = (w >> 1) & 7;
    float dir = dec_dir((w >> 4) & 3);
    int walking = (state == 0), was_avoiding = (state == 1);
    if (walking && obstacle_in_front(p, pv, pa)) {
      dir = (pa < 0.0f) ? -1.0f : 1.0f;
      steps = dur[0];
      state = 1;
    }
    if (was_avoiding) {
      steps = steps - 1;
      if (steps <= 0) state = 0;
    }
    if (was_avoiding) { l = dir * ms; r = -dir * ms; } else { l = ms; r = ms; }
    w = (w & ~63) | (state & 1) | ((steps & 7) << 1) | (enc_dir(dir) << 4);
  } else if (id == 4 || id == 5) { /* BEH:343-393, 395-516 */
    int sh = (id == 4) ? 6 : 12;
    int g = (w >> sh) & 63;
    int avoiding = g & 1, steps = (g >> 1) & 7;
    float dir = dec_dir((g >> 4) & 3);
    int was_avoiding = avoiding;
    if (was_avoiding) {
      steps = steps - 1;
      if (steps <= 0) avoiding = 0;
    }
    int not_avoiding = !was_avoiding && !avoiding;
    int trigger = not_avoiding && obstacle_in_front(p, pv, pa);
    if (trigger) {
      dir = (pa < 0.0f) ? -1.0f : 1.0f;
      steps = dur[id == 4 ? 1 : 2];
      avoiding = 1;
    }
    float lx = lv * cr_cosf(la), ly = lv * cr_sinf(la);
    float px = pv * cr_cosf(pa), py = pv * cr_sinf(pa);
    float rx, ry;
    if (id == 4) { rx = lx - 0.5f * px; ry = ly - 0.5f * py; }
    else { rx = -lx - 0.5f * px; ry = -ly - 0.5f * py; }
    steer(rx, ry, ms, &l, &r);
    if (was_avoiding) { l = dir * ms; r = -dir * ms; }
    if (trigger) { l = prev_l; r = prev_r; }
    g = (avoiding & 1) | ((steps & 7) << 1) | (enc_dir(dir) << 4);
    w = (w & ~(63 << sh)) | (g << sh);
  } else if (id == 2) { /* BEH:518-545 */
    float px = pv * cr_cosf(pa), py = pv * cr_sinf(pa);
    steer(rabx - 0.6f * px, raby - 0.6f * py, ms, &l, &r);
  } else if (id == 3) { /* BEH:547-574 */
    float px = pv * cr_cosf(pa), py = pv * cr_sinf(pa);
    steer(-p->alpha * rabx - 0.5f * px, -p->alpha * raby - 0.5f * py, ms, &l, &r);
  }
  *fsm = w;
  *out_l = l; *out_r = r;
}

/* ---- sensors ---- */
typedef struct {
  float prox_vals[N][8], light_vals[N][8];
  float cache[6][N]; /* prox_value, prox_angle, light_value, light_angle, rab_attr_x, rab_attr_y */
  float ztilde[N], rab_proj[N][4], ground[N];
} Sensors;

/* SENS:85-293 */
static void sense_proximity(const SwarmParams* p, const Pose* s, Sensors* o) {
  for (int i = 0; i < N; ++i) {
    float cy = cr_cosf(s->yaw[i]), sy = cr_sinf(s->yaw[i]);
    float sum_x = 0.0f, sum_y = 0.0f;
    for (int k = 0; k < 8; ++k) {
      float wdx = p->cos_a[k] * cy - p->sin_a[k] * sy;
      float wdy = p->cos_a[k] * sy + p->sin_a[k] * cy;
      float best = 0.0f;
      for (int g = 0; g < p->n_segments; ++g) { /* SENS:184-242 */
        float sx = p->seg_sx[g], sY = p->seg_sy[g];
        float denom = wdx * sY - wdy * sx;
        int valid = fabsf(denom) > 1e-8f;
        float ex = p->seg_ax[g] - s->x[i], ey = p->seg_ay[g] - s->y[i];
        float t = (ex * sY - ey * sx) / (denom + 1e-12f);
        float u = (ex * wdy - ey * wdx) / (denom + 1e-12f);
        int hit = valid && (t >= 0.0f) && (t <= p->prox_range) && (u >= 0.0f) && (u <= 1.0f);
        float rd = hit ? (1.0f - t / p->prox_range) : 0.0f;
        if (rd > best) best = rd;
      }
      for (int j = 0; j < N; ++j) { /* SENS:244-293 */
        if (j == i) continue;
        float dx = s->x[j] - s->x[i], dy = s->y[j] - s->y[i];
        float dist_sq = dx * dx + dy * dy;
        float proj = wdx * dx + wdy * dy;
        float closest_sq = dist_sq - proj * proj;
        float hc_arg = p->robot_radius_sq - closest_sq;
        if (hc_arg < 0.0f) hc_arg = 0.0f;
        float hit_dist = proj - sqrtf(hc_arg);
        if (hit_dist < 0.0f) hit_dist = 0.0f;
        int hit = (proj > 0.0f) && (closest_sq <= p->robot_radius_sq) && (hit_dist <= p->prox_range);
        float rd = hit ? clampf(1.0f - hit_dist / p->prox_range, 0.0f, 1.0f) : 0.0f;
        if (rd > best) best = rd;
      }
      o->prox_vals[i][k] = best;
      sum_x += best * p->cos_a[k];
      sum_y += best * p->sin_a[k];
    }
    float mag = sqrtf(sum_x * sum_x + sum_y * sum_y);
    o->cache[0][i] = mag > 1.0f ? 1.0f : mag;
    o->cache[1][i] = cr_atan2f(sum_y, sum_x);
  }
}

/* SENS:299-356, ENV:351-362 */
static void sense_light(const SwarmParams* p, const Pose* s, Sensors* o) {
  for (int i = 0; i < N; ++i) {
    if (!p->has_light) {
      for (int k = 0; k < 8; ++k) o->light_vals[i][k] = 0.0f;
      o->cache[2][i] = 0.0f; o->cache[3][i] = 0.0f;
      continue;
    }
    float lx = p->light_x - s->x[i], ly = p->light_y - s->y[i];
    float dist = sqrtf(lx * lx + ly * ly + 1e-6f);
    float base = p->light_intensity / (dist / p->unit_scale);
    float cy = cr_cosf(s->yaw[i]), sy = cr_sinf(s->yaw[i]);
    float nlx = lx / (dist + 1e-8f), nly = ly / (dist + 1e-8f);
    float mx = -INFINITY, sum_x = 0.0f, sum_y = 0.0f;
    for (int k = 0; k < 8; ++k) {
      float wdx = p->cos_a[k] * cy - p->sin_a[k] * sy;
      float wdy = p->cos_a[k] * sy + p->sin_a[k] * cy;
      float dot = wdx * nlx + wdy * nly;
      if (dot < 0.0f) dot = 0.0f;
      float raw = base * dot;
      o->light_vals[i][k] = clampf(raw, 0.0f, 1.0f);
      if (raw > mx) mx = raw;
      sum_x += raw * p->cos_a[k];
      sum_y += raw * p->sin_a[k];
    }
    int above = mx > p->light_threshold;
    o->cache[2][i] = above ? mx : 0.0f;
    o->cache[3][i] = above ? cr_atan2f(sum_y, sum_x) : 0.0f;
  }
}

/* SENS:382-501; rab_u is (N,N) for this env */
static void sense_rab(const SwarmParams* p, const Pose* s, const float* rab_u, Sensors* o) {
  for (int i = 0; i < N; ++i) {
    float cy = cr_cosf(s->yaw[i]), sy = cr_sinf(s->yaw[i]);
    float n = 0.0f, wx = 0.0f, wy = 0.0f, axs = 0.0f, ays = 0.0f;
    for (int j = 0; j < N; ++j) {
      float dx = s->x[j] - s->x[i], dy = s->y[j] - s->y[i];
      float dist = sqrtf(dx * dx + dy * dy + 1e-8f);
      int in_range = (dist < p->rab_range) && (j != i);
      if (in_range && p->n_segments > 0) { /* SENS:462-501 line of sight */
        float rdx = dx / (dist + 1e-8f), rdy = dy / (dist + 1e-8f);
        int blocked = 0;
        for (int g = 0; g < p->n_segments; ++g) {
          float sx = p->seg_sx[g], sY = p->seg_sy[g];
          float denom = rdx * sY - rdy * sx;
          int valid = fabsf(denom) > 1e-8f;
          float ex = p->seg_ax[g] - s->x[i], ey = p->seg_ay[g] - s->y[i];
          float t = (ex * sY - ey * sx) / (denom + 1e-12f);
          float u = (ex * rdy - ey * rdx) / (denom + 1e-12f);
          if (valid && (t > 1e-5f) && (t < dist - 1e-5f) && (u >= 0.0f) && (u <= 1.0f)) blocked = 1;
        }
        in_range = in_range && !blocked;
      }
      if (p->rab_loss_probability > 0.0f) in_range = in_range && (rab_u[i * N + j] >= p->rab_loss_probability);
      float inf = in_range ? 1.0f : 0.0f;
      n += inf;
      float dist_units = dist / p->unit_scale;
      float inv_dist = 1.0f / (dist_units + 1e-8f);
      float bx = dx * cy + dy * sy;
      float by = -dx * sy + dy * cy;
      /* SENS:438-441 cos/sin(atan2(by, bx)) evaluated as the normalised vector (bx, by)/|(bx, by)| (equal
       * within 2 ulp, and free of library-dependent transcendentals); coincident robots keep the reference's
       * atan2-of-signed-zeros bearing (0 or +-float32(pi)). */
      float nrm2 = bx * bx + by * by, cb, sb;
      if (nrm2 > 0.0f) {
        float nrm = sqrtf(nrm2);
        cb = bx / nrm;
        sb = by / nrm;
      } else {
        float bearing = cr_atan2f(by, bx);
        cb = cr_cosf(bearing);
        sb = cr_sinf(bearing);
      }
      wx += inv_dist * cb * inf;
      wy += inv_dist * sb * inf;
      float aw = p->alpha / (1.0f + dist_units);
      axs += aw * cb * inf;
      ays += aw * sb * inf;
    }
    o->ztilde[i] = p->ztilde_lut[(int)n]; /* SENS:425 tabulated on the host with the reference's torch ops */
    for (int k = 0; k < 4; ++k) o->rab_proj[i][k] = wx * p->rab_cos[k] + wy * p->rab_sin[k];
    o->cache[4][i] = axs;
    o->cache[5][i] = ays;
  }
}

/* ---- spawn: ENV:1215-1240, 1259-1260 (injected uniform draws) ---- */
static void spawn_env(const SwarmParams* p, const SwarmNoise* nz, int E, int e, Pose* s) {
  for (int i = 0; i < N; ++i) {
    float x = 0.0f, y = 0.0f;
    for (int r = 0; r < nz->spawn_rounds; ++r) {
      const float* u = nz->spawn_u + (((size_t)r * E + e) * N + i) * 2;
      if (r > 0) {
        if (!(p->spawn_circle_radius > 0.0f)) break;
        float rx = x - p->spawn_cx, ry = y - p->spawn_cy;
        if (!(sqrtf(rx * rx + ry * ry) > p->spawn_circle_radius)) break;
      }
      x = p->spawn_cx + (u[0] - 0.5f) * p->spawn_sx;
      y = p->spawn_cy + (u[1] - 0.5f) * p->spawn_sy;
    }
    s->x[i] = x; s->y[i] = y;
    s->yaw[i] = nz->yaw_u[e * N + i] * 2.0f * PI_F - PI_F;
  }
}

static void load_pose(const SwarmState* st, int e, Pose* s) {
  for (int i = 0; i < N; ++i) {
    s->x[i] = st->pos[(e * N + i) * 2];
    s->y[i] = st->pos[(e * N + i) * 2 + 1];
    s->yaw[i] = st->yaw[e * N + i];
  }
}

static void store_pose(const SwarmState* st, int e, const Pose* s) {
  for (int i = 0; i < N; ++i) {
    st->pos[(e * N + i) * 2] = s->x[i];
    st->pos[(e * N + i) * 2 + 1] = s->y[i];
    st->yaw[e * N + i] = s->yaw[i];
  }
}

/* reset bookkeeping of one env after the all-env re-solve: ENV:1264-1273, FOR:140-151 */
static void finish_reset(const SwarmParams* p, const SwarmState* st, int e, const Pose* s) {
  for (int i = 0; i < N; ++i) {
    st->prev_ground[e * N + i] = ground_color(p, s->x[i], s->y[i]);
    st->fsm[e * N + i] = 0;
    if (p->mission == SWARM_FOR)
      st->mission_flags[e * N + i] = (uint8_t)((s->y[i] <= p->zone[6]) ? 2 : 0);
  }
}

/* sensors + observation packing: ENV:364-391, 1118-1148, SENS:507-539 */
static void observe_env(const SwarmParams* p, const SwarmState* st, const SwarmNoise* nz, const SwarmOut* out,
                        int e, const Pose* s) {
  Sensors o;
  sense_proximity(p, s, &o);
  sense_light(p, s, &o);
  sense_rab(p, s, nz->rab_u + (size_t)e * N * N, &o);
  for (int i = 0; i < N; ++i) {
    float g = ground_color(p, s->x[i], s->y[i]);
    float* ob = out->obs + ((size_t)e * N + i) * p->obs_dim;
    if (p->obs_dim == 24) {
      for (int k = 0; k < 8; ++k) ob[k] = o.prox_vals[i][k];
      for (int k = 0; k < 8; ++k) ob[8 + k] = o.light_vals[i][k];
      ob[16] = ob[17] = ob[18] = g;
      ob[19] = o.ztilde[i];
      for (int k = 0; k < 4; ++k) ob[20 + k] = o.rab_proj[i][k];
    } else {
      ob[0] = ob[1] = ob[2] = g;
      ob[3] = o.ztilde[i];
    }
    if (st->beh_cache)
      for (int c = 0; c < 6; ++c) st->beh_cache[((size_t)e * 6 + c) * N + i] = o.cache[c][i];
  }
}

/* mission reward: ENV:1154-1194, XOR:126-131, HOM:87-92, FOR:127-138, SHL:157-160 */
static float mission_reward(const SwarmParams* p, const SwarmState* st, int e, const Pose* s, int is_final) {
  const float* z = p->zone;
  float reward = 0.0f;
  switch (p->mission) {
    case SWARM_DGT: {
      float kp = 0.0f, km = 0.0f;
      for (int i = 0; i < N; ++i) {
        float cur = ground_color(p, s->x[i], s->y[i]);
        float prev = st->prev_ground[e * N + i];
        if (prev < 0.25f && cur > 0.75f) kp += 1.0f;
        if (prev > 0.75f && cur < 0.25f) km += 1.0f;
        st->prev_ground[e * N + i] = cur;
      }
      reward = kp - km;
      break;
    }
    case SWARM_XOR: {
      float c0 = 0.0f, c1 = 0.0f;
      for (int i = 0; i < N; ++i) {
        if (in_circle(s->x[i], s->y[i], z[0], z[1], z[4])) c0 += 1.0f;
        if (in_circle(s->x[i], s->y[i], z[2], z[3], z[4])) c1 += 1.0f;
      }
      reward = c0 > c1 ? c0 : c1;
      break;
    }
    case SWARM_HOM: {
      float c = 0.0f;
      for (int i = 0; i < N; ++i)
        if (in_circle(s->x[i], s->y[i], z[0], z[1], z[4])) c += 1.0f;
      reward = is_final ? c : 0.0f;
      break;
    }
    case SWARM_FOR: {
      for (int i = 0; i < N; ++i) {
        float x = s->x[i], y = s->y[i];
        int in_food = (fabsf(x - z[0]) <= z[5] && fabsf(y - z[1]) <= z[5]) ||
                      (fabsf(x - z[2]) <= z[5] && fabsf(y - z[3]) <= z[5]);
        int in_nest = y <= z[6];
        if (p->mc_mode) /* MC:386: the standalone env picks food up on the disc the ground sensor sees */
          in_food = in_circle(x, y, z[0], z[1], z[4]) || in_circle(x, y, z[2], z[3], z[4]);
        int has_food = (st->mission_flags[e * N + i] & 1) | in_food;
        int arrived = in_nest && has_food;
        if (p->mc_mode && (st->mission_flags[e * N + i] & 2)) arrived = 0; /* MC:389 ... & ~prev_in_nest */
        if (arrived) { reward += 1.0f; has_food = 0; }
        st->mission_flags[e * N + i] = (uint8_t)(has_food | (in_nest ? 2 : 0));
      }
      break;
    }
    case SWARM_SHL: {
      for (int i = 0; i < N; ++i) {
        float x = s->x[i], y = s->y[i];
        if (x >= z[7] && x <= z[8] && y >= z[9] && y <= z[10]) reward += 1.0f;
      }
      break;
    }
  }
  return reward;
}

static int check_args(const SwarmParams* p, const SwarmState* st, const SwarmNoise* nz, const SwarmOut* out, int E) {
  if (!p || !st || !nz || !out) return SWARM_E_NULL;
  if (p->abi_version != SWARM_ABI_VERSION) return SWARM_E_VERSION;
  if (E <= 0) return SWARM_E_SIZE;
  if (!st->pos || !st->yaw || !st->prev_ground || !st->cached_left || !st->cached_right || !st->fsm ||
      !st->mission_flags || !st->episode_length_buf || !st->episode_group_reward ||
      !st->completed_group_reward || !st->completed_terminal_critic_state || !out->obs)
    return SWARM_E_NULL;
  if (p->discrete_actions && !st->beh_cache) return SWARM_E_NULL;
  if (!nz->rab_u) return SWARM_E_NULL; /* the oracle has no RNG: noise is always injected */
  if (p->mission < 0 || p->mission > 4 || (p->obs_dim != 24 && p->obs_dim != 4) || p->decimation < 1 ||
      p->n_segments > SWARM_MAX_SEG || p->n_internal > SWARM_MAX_INTERNAL)
    return SWARM_E_PARAM;
  return 0;
}

/* One env.step (SURVEY.md 3.2 order). Host pointers everywhere. */
int swarm_oracle_step(const SwarmParams* p, const SwarmState* st, const void* actions, const SwarmNoise* nz,
                      const SwarmOut* out, int E) {
  int rc = check_args(p, st, nz, out, E);
  if (rc) return rc;
  if (!actions || !out->reward || !out->time_out) return SWARM_E_NULL;
  if (p->discrete_actions && !nz->turn_dur) return SWARM_E_NULL;
  const float ms = p->max_wheel_speed;
  int any_reset = 0;

  /* phase 1: action -> wheels -> integrate -> collide -> dones -> rewards -> respawn */
#pragma omp parallel for schedule(static) reduction(| : any_reset)
  for (int e = 0; e < E; ++e) {
    Pose s;
    load_pose(st, e, &s);
    float lw[N], rw[N];
    for (int i = 0; i < N; ++i) {
      int idx = e * N + i;
      if (p->discrete_actions) { /* ENV:774-795 */
        float c[6];
        for (int k = 0; k < 6; ++k) c[k] = st->beh_cache[((size_t)e * 6 + k) * N + i];
        dispatch_robot(p, ((const int64_t*)actions)[idx], c, st->cached_left[idx], st->cached_right[idx],
                       nz->turn_dur + (size_t)idx * 3, &st->fsm[idx], &lw[i], &rw[i]);
      } else { /* ENV:802-809 */
        const float* a = (const float*)actions + (size_t)idx * 2;
        lw[i] = clampf(a[0], -1.0f, 1.0f) * ms;
        rw[i] = clampf(a[1], -1.0f, 1.0f) * ms;
      }
    }
    for (int i = 0; i < N; ++i) {
      st->cached_left[e * N + i] = lw[i];
      st->cached_right[e * N + i] = rw[i];
    }
    for (int d = 0; d < p->decimation; ++d) { /* ENV:816-836 */
      float prx[N], pry[N];
      memcpy(prx, s.x, sizeof prx);
      memcpy(pry, s.y, sizeof pry);
      for (int i = 0; i < N; ++i) { /* SENS:592-617 */
        float v = 0.5f * (lw[i] + rw[i]);
        float omega = (rw[i] - lw[i]) / p->wheelbase;
        float cy = cr_cosf(s.yaw[i]), sy = cr_sinf(s.yaw[i]);
        s.x[i] += v * cy * p->dt;
        s.y[i] += v * sy * p->dt;
        float yw = s.yaw[i] + omega * p->dt;
        s.yaw[i] = cr_atan2f(cr_sinf(yw), cr_cosf(yw));
      }
      resolve_walls(p, &s);
      resolve_gate(p, &s);
      resolve_robots(p, &s);
      resolve_collisions(p, &s, prx, pry);
    }
    int64_t len = st->episode_length_buf[e] + 1;
    int time_out = len >= p->max_episode_length; /* ENV:1202 */
    if (time_out) critic_state(p, &s, st->completed_terminal_critic_state + (size_t)e * N * 5);
    float reward = mission_reward(p, st, e, &s, time_out);
    float acc = st->episode_group_reward[e] + reward;
    out->reward[e] = reward;
    out->time_out[e] = (uint8_t)time_out;
    if (time_out) { /* ENV:1253-1260 */
      st->completed_group_reward[e] = acc;
      acc = 0.0f;
      len = 0;
      spawn_env(p, nz, E, e, &s);
      any_reset = 1;
    }
    st->episode_group_reward[e] = acc;
    st->episode_length_buf[e] = len;
    store_pose(st, e, &s);
  }

  /* phase 2: ENV:1262 re-solves ALL envs when any env reset; then observe.  A sharded job may hand in the job-wide
   * flag (ABI v2, SwarmNoise.any_reset_mode). */
  if (nz->any_reset_mode == 1) any_reset = (int)(nz->any_reset_bits & 1u);
#pragma omp parallel for schedule(static)
  for (int e = 0; e < E; ++e) {
    Pose s;
    load_pose(st, e, &s);
    if (any_reset) {
      resolve_collisions(p, &s, NULL, NULL);
      store_pose(st, e, &s);
      if (out->time_out[e]) finish_reset(p, st, e, &s);
    }
    observe_env(p, st, nz, out, e, &s);
    if (out->critic) critic_state(p, &s, out->critic + (size_t)e * N * 5); /* ABI v2: ENV:1279-1290 of the new state */
  }
  return 0;
}

/* env.reset(): ENV:1242-1273 over all envs, then ENV:1118-1148 */
int swarm_oracle_reset(const SwarmParams* p, const SwarmState* st, const SwarmNoise* nz, const SwarmOut* out, int E) {
  int rc = check_args(p, st, nz, out, E);
  if (rc) return rc;
  if (!nz->spawn_u || !nz->yaw_u || nz->spawn_rounds < 1) return SWARM_E_NULL;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < E; ++e) {
    Pose s;
    st->episode_length_buf[e] = 0;
    st->completed_group_reward[e] = st->episode_group_reward[e];
    st->episode_group_reward[e] = 0.0f;
    spawn_env(p, nz, E, e, &s);
    resolve_collisions(p, &s, NULL, NULL);
    store_pose(st, e, &s);
    finish_reset(p, st, e, &s);
    observe_env(p, st, nz, out, e, &s);
    if (out->critic) critic_state(p, &s, out->critic + (size_t)e * N * 5);
  }
  return 0;
}

int swarm_oracle_critic_state(const SwarmParams* p, const SwarmState* st, float* critic_out, int E) {
  if (!p || !st || !critic_out || !st->pos || !st->yaw) return SWARM_E_NULL;
  if (E <= 0) return SWARM_E_SIZE;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < E; ++e) {
    Pose s;
    load_pose(st, e, &s);
    critic_state(p, &s, critic_out + (size_t)e * N * 5);
  }
  return 0;
}

int swarm_oracle_abi_version(void) { return SWARM_ABI_VERSION; }

/* ======================= scripts/manual_control.py compatibility (BASELINE config 1) ======================= */

/* MC:531-553: faces one after another (Gauss-Seidel), r = robot_radius, face data derived from angles. */
static void mc_resolve_walls(const SwarmParams* p, Pose* s) {
  for (int f = 0; f < 12; ++f) {
    const float nx = p->mc_face_nx[f], ny = p->mc_face_ny[f];
    for (int i = 0; i < N; ++i) {
      float dx = s->x[i] - p->mc_face_px[f];
      float dy = s->y[i] - p->mc_face_py[f];
      float sd = dx * nx + dy * ny;
      float pen = p->robot_radius - sd;
      if (pen > 0.0f) {
        s->x[i] += pen * nx;
        s->y[i] += pen * ny;
      }
    }
  }
}

/* MC:245-269 */
static void mc_reset_env(const SwarmParams* p, const SwarmState* st, const SwarmNoise* nz, int e, Pose* s) {
  for (int i = 0; i < N; ++i) {
    const float* u = nz->mc_spawn_u + ((size_t)e * N + i) * 3;
    float r = sqrtf(u[0]) * p->mc_spawn_safe;
    float th = u[1] * p->mc_spawn_theta_max;
    s->x[i] = r * cr_cosf(th);
    s->y[i] = r * cr_sinf(th);
    if (p->mission == SWARM_HOM) s->y[i] = fabsf(s->y[i]);
    s->yaw[i] = u[2] * 2.0f * PI_F - PI_F;
    st->prev_ground[e * N + i] = ground_color(p, s->x[i], s->y[i]);
    st->fsm[e * N + i] = 0;
    st->mission_flags[e * N + i] = (uint8_t)((p->mission == SWARM_FOR && s->y[i] <= p->zone[6]) ? 2 : 0);
  }
  st->episode_length_buf[e] = 0;
  st->episode_group_reward[e] = 0.0f;
}

int swarm_oracle_mc_reset(const SwarmParams* p, const SwarmState* st, const SwarmNoise* nz, int E) {
  if (!p || !st || !nz || !nz->mc_spawn_u) return SWARM_E_NULL;
  if (!p->mc_mode) return SWARM_E_PARAM;
  for (int e = 0; e < E; ++e) {
    Pose s;
    mc_reset_env(p, st, nz, e, &s);
    store_pose(st, e, &s);
  }
  return 0;
}

/* One tick of the manual-control loop, MC:721-757 (see swarm_mc_tick in include/swarm_abi.h). */
int swarm_oracle_mc_tick(const SwarmParams* p, const SwarmState* st, const int64_t* module_ids, const float* wheels,
                         const SwarmNoise* nz, const SwarmOut* out, int flags, int E) {
  if (!p || !st || !nz || !out) return SWARM_E_NULL;
  if (!p->mc_mode || p->obs_dim != 24) return SWARM_E_PARAM;
  if ((flags & SWARM_MC_PRE) && (!module_ids || !nz->rab_u || !nz->turn_dur)) return SWARM_E_NULL;
  if ((flags & SWARM_MC_PHYSICS) && (!wheels || !nz->mc_spawn_u || !out->reward || !out->time_out)) return SWARM_E_NULL;
  if ((flags & SWARM_MC_POST) && (!nz->rab_u2 || !out->obs)) return SWARM_E_NULL;
  const float ms = p->max_wheel_speed;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < E; ++e) {
    Pose s;
    load_pose(st, e, &s);
    float lw[N], rw[N];
    for (int i = 0; i < N; ++i) {
      lw[i] = wheels ? wheels[((size_t)e * N + i) * 2] : 0.0f;
      rw[i] = wheels ? wheels[((size_t)e * N + i) * 2 + 1] : 0.0f;
    }
    if (flags & SWARM_MC_PRE) { /* MC:729-749 */
      Sensors o;
      sense_proximity(p, &s, &o);
      sense_light(p, &s, &o);
      sense_rab(p, &s, nz->rab_u + (size_t)e * N * N, &o);
      for (int i = 0; i < N; ++i) {
        float c[6], l, r;
        for (int k = 0; k < 6; ++k) c[k] = o.cache[k][i];
        dispatch_robot(p, module_ids[e * N + i], c, 0.0f, 0.0f, nz->turn_dur + ((size_t)e * N + i) * 3,
                       &st->fsm[e * N + i], &l, &r);
        if (i > 0) { lw[i] = l; rw[i] = r; } /* robot 0 keeps the keyboard command (MC:725-726, 748-749) */
      }
    }
    if (flags & SWARM_MC_PHYSICS) { /* MC:355-423 */
      for (int i = 0; i < N; ++i) {
        float l = clampf(lw[i], -ms, ms), r = clampf(rw[i], -ms, ms);
        float v = 0.5f * (l + r);
        float omega = (r - l) / p->wheelbase;
        float cy = cr_cosf(s.yaw[i]), sy = cr_sinf(s.yaw[i]);
        s.x[i] += v * cy * p->dt;
        s.y[i] += v * sy * p->dt;
        float yw = s.yaw[i] + omega * p->dt;
        s.yaw[i] = cr_atan2f(cr_sinf(yw), cr_cosf(yw));
      }
      mc_resolve_walls(p, &s);
      resolve_gate(p, &s);
      resolve_robots(p, &s);
      int64_t len = st->episode_length_buf[e] + 1;
      int final_step = len >= p->max_episode_length; /* MC:380 */
      float reward = mission_reward(p, st, e, &s, final_step);
      float acc = st->episode_group_reward[e] + reward;
      out->reward[e] = reward;
      out->time_out[e] = (uint8_t)final_step;
      st->episode_group_reward[e] = acc;
      st->episode_length_buf[e] = len;
      if (final_step) { /* MC:753-754 reset(advance_episode=True) */
        st->completed_group_reward[e] = acc;
        mc_reset_env(p, st, nz, e, &s);
      }
    }
    store_pose(st, e, &s);
    if (flags & SWARM_MC_POST) { /* MC:425-440 */
      SwarmNoise nz2 = *nz;
      nz2.rab_u = nz->rab_u2;
      SwarmState st2 = *st;
      st2.beh_cache = NULL; /* compute_obs_robot0 does not feed the behaviour modules */
      observe_env(p, &st2, &nz2, out, e, &s);
    }
  }
  return 0;
}

/* number of OpenMP threads the env loops use (torchrun exports OMP_NUM_THREADS=1 by default) */
int swarm_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

/* test hooks for include/swarm_detmath.h */
void swarm_oracle_sincos(int n, const float* a, float* sn, float* cs) {
  for (int i = 0; i < n; ++i) swarm_sincosf(a[i], &sn[i], &cs[i]);
}
void swarm_oracle_atan2(int n, const float* y, const float* x, float* out) {
  for (int i = 0; i < n; ++i) out[i] = swarm_atan2f(y[i], x[i]);
}
