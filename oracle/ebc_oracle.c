/*
 * ebc_oracle.c — CPU restatement of the EB-CADRL hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under eb-cadrl_b200/ may import, link or execute
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it, and only as the checker / reported CPU baseline.
 *
 * Parity status
 *   - In-tree reference code (env step, collisions, reward, lookahead, rotate, value
 *     net): PINNED.  tests/golden/ holds vectors produced by importing the unmodified
 *     reference (/root/reference) with oracle/shims on PYTHONPATH; tests/test_oracle_golden.py
 *     replays them through this file.
 *   - rvo2 (Python-RVO2 over RVO2 Library v2.0.x, un-vendored and un-pinned by the
 *     reference, absent from this image): restated below from the published RVO2
 *     algorithm (Agent::computeNeighbors / insertAgentNeighbor / computeNewVelocity,
 *     linearProgram1/2/3, RVO_EPSILON = 1e-5f).  Numerically "parity unpinned" (the
 *     reference holds no float golden for it); pinned at event level by the reference's
 *     own episode-outcome tests (tests/test_collisions_simulation.py,
 *     tests/test_basic_simulation.py), which pass with this restatement plugged in as
 *     `rvo2` (oracle/run_reference_tests.py).
 *
 * Arithmetic: rvo2 part in IEEE fp32, everything the reference does in Python floats in
 * fp64 (promoted from the fp32 state), the rotate / value-net part in fp32 (the reference
 * casts to torch.float32 at rl/policy/multi_human_rl.py:52-60).  Build with
 * -ffp-contract=off so that no multiply-add is fused (RVO2 on baseline x86-64 has none).
 *
 * Every function cites the reference file:line it follows.
 */
#include "../include/ebcadrl.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define RVO_EPSILON 0.00001f
#define MAX_LINES 160 /* >= obstacle lines (<= 64) + agent neighbours (orca_max_neighbors <= 64) */

typedef struct { float px, py, dx, dy; } line_t; /* RVO2 Line: point, direction */

static inline float det2(float ax, float ay, float bx, float by) { return ax * by - ay * bx; }
static inline float dot2(float ax, float ay, float bx, float by) { return ax * bx + ay * by; }

/* RVO2 Agent.cpp linearProgram1. */
static int lp1(const line_t *L, int i, float radius, float ox, float oy, int dir_opt, float *rx,
               float *ry) {
  const float dp = dot2(L[i].px, L[i].py, L[i].dx, L[i].dy);
  const float disc = dp * dp + radius * radius - dot2(L[i].px, L[i].py, L[i].px, L[i].py);
  if (disc < 0.0f) return 0;
  const float s = sqrtf(disc);
  float tl = -dp - s;
  float tr = -dp + s;
  for (int j = 0; j < i; ++j) {
    const float den = det2(L[i].dx, L[i].dy, L[j].dx, L[j].dy);
    const float num = det2(L[j].dx, L[j].dy, L[i].px - L[j].px, L[i].py - L[j].py);
    if (fabsf(den) <= RVO_EPSILON) {
      if (num < 0.0f) return 0;
      continue;
    }
    const float t = num / den;
    if (den >= 0.0f) {
      tr = fminf(tr, t);
    } else {
      tl = fmaxf(tl, t);
    }
    if (tl > tr) return 0;
  }
  float t;
  if (dir_opt) {
    t = (dot2(ox, oy, L[i].dx, L[i].dy) > 0.0f) ? tr : tl;
  } else {
    t = dot2(L[i].dx, L[i].dy, ox - L[i].px, oy - L[i].py);
    if (t < tl) t = tl;
    else if (t > tr) t = tr;
  }
  *rx = L[i].px + t * L[i].dx;
  *ry = L[i].py + t * L[i].dy;
  return 1;
}

/* RVO2 Agent.cpp linearProgram2. */
static int lp2(const line_t *L, int n, float radius, float ox, float oy, int dir_opt, float *rx,
               float *ry) {
  if (dir_opt) {
    *rx = ox * radius;
    *ry = oy * radius;
  } else if (dot2(ox, oy, ox, oy) > radius * radius) {
    const float inv = 1.0f / sqrtf(dot2(ox, oy, ox, oy)); /* normalize(v) = v / abs(v) */
    *rx = (ox * inv) * radius;
    *ry = (oy * inv) * radius;
  } else {
    *rx = ox;
    *ry = oy;
  }
  for (int i = 0; i < n; ++i) {
    if (det2(L[i].dx, L[i].dy, L[i].px - *rx, L[i].py - *ry) > 0.0f) {
      const float tx = *rx, ty = *ry;
      if (!lp1(L, i, radius, ox, oy, dir_opt, rx, ry)) {
        *rx = tx;
        *ry = ty;
        return i;
      }
    }
  }
  return n;
}

/* RVO2 Agent.cpp linearProgram3 (numObstLines = 0 on the reference's live path). */
static void lp3(const line_t *L, int n, int num_obst, int begin, float radius, float *rx,
                float *ry) {
  float distance = 0.0f;
  line_t proj[MAX_LINES];
  for (int i = begin; i < n; ++i) {
    if (det2(L[i].dx, L[i].dy, L[i].px - *rx, L[i].py - *ry) > distance) {
      int m = 0;
      for (int j = 0; j < num_obst; ++j) proj[m++] = L[j];
      for (int j = num_obst; j < i; ++j) {
        line_t ln;
        const float d = det2(L[i].dx, L[i].dy, L[j].dx, L[j].dy);
        if (fabsf(d) <= RVO_EPSILON) {
          if (dot2(L[i].dx, L[i].dy, L[j].dx, L[j].dy) > 0.0f) continue;
          ln.px = 0.5f * (L[i].px + L[j].px);
          ln.py = 0.5f * (L[i].py + L[j].py);
        } else {
          const float t = det2(L[j].dx, L[j].dy, L[i].px - L[j].px, L[i].py - L[j].py) / d;
          ln.px = L[i].px + t * L[i].dx;
          ln.py = L[i].py + t * L[i].dy;
        }
        const float ddx = L[j].dx - L[i].dx, ddy = L[j].dy - L[i].dy;
        const float inv = 1.0f / sqrtf(dot2(ddx, ddy, ddx, ddy));
        ln.dx = ddx * inv;
        ln.dy = ddy * inv;
        proj[m++] = ln;
      }
      const float tx = *rx, ty = *ry;
      if (lp2(proj, m, radius, -L[i].dy, L[i].dx, 1, rx, ry) < m) {
        *rx = tx;
        *ry = ty;
      }
      distance = det2(L[i].dx, L[i].dy, L[i].px - *rx, L[i].py - *ry);
    }
  }
}

/* ------------------------------------------------------------------------------------
 * RVO2 static obstacles: RVOSimulator::addObstacle, KdTree::buildObstacleTree /
 * queryObstacleTreeRecursive, Agent::insertObstacleNeighbor and the obstacle section of
 * Agent::computeNewVelocity, restated from the published RVO2 v2.0 sources [RVO2-from-memory].
 * Reference caller: simulator/policy/orca_obstacles.py:102-107 (addObstacle per wall,
 * processObstacles) -- dead code on the reference's live path, so this part is "parity
 * unpinned" like the rest of rvo2; the CUDA kernel is held bit-exact against THIS.
 * ---------------------------------------------------------------------------------- */
#define OBST_MAX 64
typedef struct {
  float px, py, ux, uy;
  int next, prev, convex;
} obst_t;

static inline float left_of(float ax, float ay, float bx, float by, float cx, float cy) {
  return det2(ax - cx, ay - cy, bx - ax, by - ay); /* RVO2 leftOf(a, b, c) = det(a - c, b - a) */
}

typedef struct {
  obst_t v[OBST_MAX];
  int n, cap, overflow;
  int left[OBST_MAX], right[OBST_MAX]; /* kd-tree children (vertex ids, -1 = none) */
  int root;
  uint64_t anc[OBST_MAX], anc_left[OBST_MAX];
} obst_set;

/* KdTree::buildObstacleTreeRecursive; returns the node (= vertex id) or -1 */
static int build_obstacle_tree(obst_set *S, const int *list, int count, uint64_t anc, uint64_t anc_left) {
  if (count == 0 || S->overflow) return -1;
  int optimal = 0, min_left = count, min_right = count;
  for (int i = 0; i < count; ++i) {
    int ls = 0, rs = 0;
    const obst_t *i1 = &S->v[list[i]], *i2 = &S->v[i1->next];
    for (int j = 0; j < count; ++j) {
      if (i == j) continue;
      const obst_t *j1 = &S->v[list[j]], *j2 = &S->v[j1->next];
      const float l1 = left_of(i1->px, i1->py, i2->px, i2->py, j1->px, j1->py);
      const float l2 = left_of(i1->px, i1->py, i2->px, i2->py, j2->px, j2->py);
      if (l1 >= -RVO_EPSILON && l2 >= -RVO_EPSILON) ++ls;
      else if (l1 <= RVO_EPSILON && l2 <= RVO_EPSILON) ++rs;
      else { ++ls; ++rs; }
      /* std::make_pair(max, min) >= std::make_pair(max(minLeft, minRight), min(minLeft, minRight)) */
      const int a = ls > rs ? ls : rs, b = ls > rs ? rs : ls;
      const int c = min_left > min_right ? min_left : min_right, d = min_left > min_right ? min_right : min_left;
      if (a > c || (a == c && b >= d)) break;
    }
    const int a = ls > rs ? ls : rs, b = ls > rs ? rs : ls;
    const int c = min_left > min_right ? min_left : min_right, d = min_left > min_right ? min_right : min_left;
    if (a < c || (a == c && b < d)) { min_left = ls; min_right = rs; optimal = i; }
  }
  int lefts[2 * OBST_MAX], rights[2 * OBST_MAX], nl = 0, nr = 0;
  const int node = list[optimal];
  for (int j = 0; j < count; ++j) {
    if (j == optimal) continue;
    const obst_t *i1 = &S->v[node], *i2 = &S->v[i1->next];
    obst_t *j1 = &S->v[list[j]], *j2 = &S->v[j1->next];
    const float l1 = left_of(i1->px, i1->py, i2->px, i2->py, j1->px, j1->py);
    const float l2 = left_of(i1->px, i1->py, i2->px, i2->py, j2->px, j2->py);
    if (l1 >= -RVO_EPSILON && l2 >= -RVO_EPSILON) lefts[nl++] = list[j];
    else if (l1 <= RVO_EPSILON && l2 <= RVO_EPSILON) rights[nr++] = list[j];
    else { /* split obstacle j at the line through i */
      if (S->n >= S->cap) { S->overflow = 1; return -1; }
      const float t = det2(i2->px - i1->px, i2->py - i1->py, j1->px - i1->px, j1->py - i1->py) /
                      det2(i2->px - i1->px, i2->py - i1->py, j1->px - j2->px, j1->py - j2->py);
      const int nid = S->n++;
      obst_t *nw = &S->v[nid];
      nw->px = j1->px + t * (j2->px - j1->px);
      nw->py = j1->py + t * (j2->py - j1->py);
      nw->prev = list[j];
      nw->next = j1->next;
      nw->convex = 1;
      nw->ux = j1->ux;
      nw->uy = j1->uy;
      j2->prev = nid;
      j1->next = nid;
      if (l1 > 0.0f) { lefts[nl++] = list[j]; rights[nr++] = nid; }
      else { rights[nr++] = list[j]; lefts[nl++] = nid; }
    }
  }
  S->anc[node] = anc;
  S->anc_left[node] = anc_left;
  const uint64_t bit = (uint64_t)1 << node;
  const int l = build_obstacle_tree(S, lefts, nl, anc | bit, anc_left | bit);
  const int r = build_obstacle_tree(S, rights, nr, anc | bit, anc_left);
  S->left[node] = l;
  S->right[node] = r;
  return node;
}

/* RVOSimulator::addObstacle for every polygon, then processObstacles (= buildObstacleTree). */
static int pack_obstacles(const float *xy, const int32_t *poly_size, int32_t n_poly, int cap, obst_set *S) {
  memset(S, 0, sizeof(*S));
  S->cap = cap < OBST_MAX ? cap : OBST_MAX;
  int off = 0;
  for (int p = 0; p < n_poly; ++p) {
    const int m = poly_size[p];
    if (m < 2 || S->n + m > S->cap) return -1;
    const float *q = xy + (size_t)off * 2;
    const int base = S->n;
    for (int i = 0; i < m; ++i) {
      obst_t *o = &S->v[base + i];
      const int nx = i == m - 1 ? 0 : i + 1, pr = i == 0 ? m - 1 : i - 1;
      o->px = q[i * 2]; o->py = q[i * 2 + 1];
      o->next = base + nx; o->prev = base + pr;
      const float dx = q[nx * 2] - q[i * 2], dy = q[nx * 2 + 1] - q[i * 2 + 1];
      const float inv = 1.0f / sqrtf(dot2(dx, dy, dx, dy)); /* normalize(v) = v / abs(v), v / s = v * (1 / s) */
      o->ux = dx * inv; o->uy = dy * inv;
      o->convex = m == 2 ? 1 : (left_of(q[pr * 2], q[pr * 2 + 1], q[i * 2], q[i * 2 + 1], q[nx * 2], q[nx * 2 + 1]) >= 0.0f);
    }
    S->n += m;
    off += m;
  }
  int list[OBST_MAX];
  for (int i = 0; i < S->n; ++i) list[i] = i;
  const int n0 = S->n;
  S->root = build_obstacle_tree(S, list, n0, 0, 0);
  return S->overflow ? -1 : 0;
}

int ebc_ref_pack_obstacles(const float *xy, const int32_t *poly_size, int32_t n_poly, ebc_obst_vertex *out,
                           int32_t cap, int32_t *n_out) {
  if (!poly_size || !out || !n_out || n_poly < 0 || cap < 0 || (n_poly > 0 && !xy)) return EBC_ERR_INVALID;
  obst_set S;
  if (pack_obstacles(xy, poly_size, n_poly, cap, &S)) return EBC_ERR_INVALID;
  for (int i = 0; i < S.n; ++i) {
    memset(&out[i], 0, sizeof(out[i]));
    out[i].px = S.v[i].px; out[i].py = S.v[i].py; out[i].ux = S.v[i].ux; out[i].uy = S.v[i].uy;
    out[i].next = (int16_t)S.v[i].next; out[i].prev = (int16_t)S.v[i].prev;
    out[i].convex = (int16_t)S.v[i].convex;
    out[i].anc = S.anc[i]; out[i].anc_left = S.anc_left[i];
  }
  *n_out = S.n;
  return EBC_OK;
}

/* Rebuild the tree's child links from the packed records (anc / anc_left). */
static void unpack_obstacles(const ebc_obst_vertex *rec, int n, obst_set *S) {
  memset(S, 0, sizeof(*S));
  S->n = n;
  S->root = -1;
  for (int i = 0; i < n; ++i) {
    S->v[i].px = rec[i].px; S->v[i].py = rec[i].py; S->v[i].ux = rec[i].ux; S->v[i].uy = rec[i].uy;
    S->v[i].next = rec[i].next; S->v[i].prev = rec[i].prev; S->v[i].convex = rec[i].convex;
    S->anc[i] = rec[i].anc; S->anc_left[i] = rec[i].anc_left;
    S->left[i] = S->right[i] = -1;
  }
  for (int i = 0; i < n; ++i) {
    if (S->anc[i] == 0) { S->root = i; continue; }
    /* parent = the ancestor whose own ancestor set is anc[i] minus itself */
    for (int z = 0; z < n; ++z) {
      const uint64_t bit = (uint64_t)1 << z;
      if ((S->anc[i] & bit) && S->anc[z] == (S->anc[i] & ~bit)) {
        if (S->anc_left[i] & bit) S->left[z] = i; else S->right[z] = i;
        break;
      }
    }
  }
}

/* distSqPointLineSegment(a, b, c) */
static inline float dist_sq_point_segment(float ax, float ay, float bx, float by, float cx, float cy) {
  const float r = dot2(cx - ax, cy - ay, bx - ax, by - ay) / dot2(bx - ax, by - ay, bx - ax, by - ay);
  if (r < 0.0f) return dot2(cx - ax, cy - ay, cx - ax, cy - ay);
  if (r > 1.0f) return dot2(cx - bx, cy - by, cx - bx, cy - by);
  const float qx = cx - (ax + r * (bx - ax)), qy = cy - (ay + r * (by - ay));
  return dot2(qx, qy, qx, qy);
}

typedef struct { int id[OBST_MAX]; float d[OBST_MAX]; int n; } obst_nb;

/* KdTree::queryObstacleTreeRecursive + Agent::insertObstacleNeighbor */
static void query_obstacle_tree(const obst_set *S, int node, float px, float py, float range_sq, obst_nb *nb) {
  if (node < 0) return;
  const obst_t *o1 = &S->v[node], *o2 = &S->v[o1->next];
  const float agent_left = left_of(o1->px, o1->py, o2->px, o2->py, px, py);
  query_obstacle_tree(S, agent_left >= 0.0f ? S->left[node] : S->right[node], px, py, range_sq, nb);
  const float ex = o2->px - o1->px, ey = o2->py - o1->py;
  const float dist_sq_line = agent_left * agent_left / dot2(ex, ey, ex, ey);
  if (dist_sq_line < range_sq) {
    if (agent_left < 0.0f) { /* agent on the right side: it can see this edge */
      const float d = dist_sq_point_segment(o1->px, o1->py, o2->px, o2->py, px, py);
      if (d < range_sq) {
        int i = nb->n++;
        while (i != 0 && d < nb->d[i - 1]) { nb->d[i] = nb->d[i - 1]; nb->id[i] = nb->id[i - 1]; --i; }
        nb->d[i] = d; nb->id[i] = node;
      }
    }
    query_obstacle_tree(S, agent_left >= 0.0f ? S->right[node] : S->left[node], px, py, range_sq, nb);
  }
}

/* Agent::computeNewVelocity, "Create obstacle ORCA lines".  Returns numObstLines. */
static int obstacle_lines(const obst_set *S, float px, float py, float vx, float vy, float radius, float max_speed,
                          float time_horizon_obst, line_t *L) {
  obst_nb nb;
  nb.n = 0;
  const float rr = time_horizon_obst * max_speed + radius;
  query_obstacle_tree(S, S->root, px, py, rr * rr, &nb); /* Agent::computeNeighbors */
  const float inv_th = 1.0f / time_horizon_obst;
  int n = 0;
  for (int i = 0; i < nb.n; ++i) {
    const obst_t *o1 = &S->v[nb.id[i]], *o2 = &S->v[o1->next];
    const float r1x = o1->px - px, r1y = o1->py - py, r2x = o2->px - px, r2y = o2->py - py;
    int covered = 0;
    for (int j = 0; j < n; ++j) {
      if (det2(inv_th * r1x - L[j].px, inv_th * r1y - L[j].py, L[j].dx, L[j].dy) - inv_th * radius >= -RVO_EPSILON &&
          det2(inv_th * r2x - L[j].px, inv_th * r2y - L[j].py, L[j].dx, L[j].dy) - inv_th * radius >= -RVO_EPSILON) {
        covered = 1;
        break;
      }
    }
    if (covered) continue;
    const float d1 = dot2(r1x, r1y, r1x, r1y), d2 = dot2(r2x, r2y, r2x, r2y);
    const float rsq = radius * radius;
    const float ovx = o2->px - o1->px, ovy = o2->py - o1->py;
    const float s = dot2(-r1x, -r1y, ovx, ovy) / dot2(ovx, ovy, ovx, ovy);
    const float qx = -r1x - s * ovx, qy = -r1y - s * ovy;
    const float dline = dot2(qx, qy, qx, qy);
    line_t ln;
    if (s < 0.0f && d1 <= rsq) { /* collision with the left vertex; ignore if non-convex */
      if (o1->convex) {
        ln.px = 0.0f; ln.py = 0.0f;
        const float inv = 1.0f / sqrtf(dot2(-r1y, r1x, -r1y, r1x));
        ln.dx = -r1y * inv; ln.dy = r1x * inv;
        L[n++] = ln;
      }
      continue;
    } else if (s > 1.0f && d2 <= rsq) { /* right vertex; ignore if non-convex or handled by the neighbouring edge */
      if (o2->convex && det2(r2x, r2y, o2->ux, o2->uy) >= 0.0f) {
        ln.px = 0.0f; ln.py = 0.0f;
        const float inv = 1.0f / sqrtf(dot2(-r2y, r2x, -r2y, r2x));
        ln.dx = -r2y * inv; ln.dy = r2x * inv;
        L[n++] = ln;
      }
      continue;
    } else if (s >= 0.0f && s < 1.0f && dline <= rsq) { /* collision with the segment */
      ln.px = 0.0f; ln.py = 0.0f;
      ln.dx = -o1->ux; ln.dy = -o1->uy;
      L[n++] = ln;
      continue;
    }
    /* no collision: legs */
    float llx, lly, rlx, rly;
    const obst_t *a1 = o1, *a2 = o2; /* obstacle1 / obstacle2 after the oblique-view reassignment */
    if (s < 0.0f && dline <= rsq) {
      if (!o1->convex) continue;
      a2 = o1;
      const float leg1 = sqrtf(d1 - rsq);
      const float inv = 1.0f / d1;
      llx = (r1x * leg1 - r1y * radius) * inv; lly = (r1x * radius + r1y * leg1) * inv;
      rlx = (r1x * leg1 + r1y * radius) * inv; rly = (-r1x * radius + r1y * leg1) * inv;
    } else if (s > 1.0f && dline <= rsq) {
      if (!o2->convex) continue;
      a1 = o2;
      const float leg2 = sqrtf(d2 - rsq);
      const float inv = 1.0f / d2;
      llx = (r2x * leg2 - r2y * radius) * inv; lly = (r2x * radius + r2y * leg2) * inv;
      rlx = (r2x * leg2 + r2y * radius) * inv; rly = (-r2x * radius + r2y * leg2) * inv;
    } else {
      if (o1->convex) {
        const float leg1 = sqrtf(d1 - rsq);
        const float inv = 1.0f / d1;
        llx = (r1x * leg1 - r1y * radius) * inv; lly = (r1x * radius + r1y * leg1) * inv;
      } else { llx = -o1->ux; lly = -o1->uy; }
      if (o2->convex) {
        const float leg2 = sqrtf(d2 - rsq);
        const float inv = 1.0f / d2;
        rlx = (r2x * leg2 + r2y * radius) * inv; rly = (-r2x * radius + r2y * leg2) * inv;
      } else { rlx = o1->ux; rly = o1->uy; }
    }
    /* legs never point into the neighbouring edge of a convex vertex */
    const obst_t *ln_left = &S->v[a1->prev];
    int left_foreign = 0, right_foreign = 0;
    if (a1->convex && det2(llx, lly, -ln_left->ux, -ln_left->uy) >= 0.0f) {
      llx = -ln_left->ux; lly = -ln_left->uy;
      left_foreign = 1;
    }
    if (a2->convex && det2(rlx, rly, a2->ux, a2->uy) <= 0.0f) {
      rlx = a2->ux; rly = a2->uy;
      right_foreign = 1;
    }
    const float lcx = inv_th * (a1->px - px), lcy = inv_th * (a1->py - py);
    const float rcx = inv_th * (a2->px - px), rcy = inv_th * (a2->py - py);
    const float cvx = rcx - lcx, cvy = rcy - lcy;
    const int same = a1 == a2;
    const float t = same ? 0.5f : dot2(vx - lcx, vy - lcy, cvx, cvy) / dot2(cvx, cvy, cvx, cvy);
    const float tl = dot2(vx - lcx, vy - lcy, llx, lly);
    const float tr = dot2(vx - rcx, vy - rcy, rlx, rly);
    if ((t < 0.0f && tl < 0.0f) || (same && tl < 0.0f && tr < 0.0f)) { /* left cut-off circle */
      const float wx = vx - lcx, wy = vy - lcy;
      const float inv = 1.0f / sqrtf(dot2(wx, wy, wx, wy));
      const float ux = wx * inv, uy = wy * inv;
      ln.dx = uy; ln.dy = -ux;
      ln.px = lcx + radius * inv_th * ux; ln.py = lcy + radius * inv_th * uy;
      L[n++] = ln;
      continue;
    } else if (t > 1.0f && tr < 0.0f) { /* right cut-off circle */
      const float wx = vx - rcx, wy = vy - rcy;
      const float inv = 1.0f / sqrtf(dot2(wx, wy, wx, wy));
      const float ux = wx * inv, uy = wy * inv;
      ln.dx = uy; ln.dy = -ux;
      ln.px = rcx + radius * inv_th * ux; ln.py = rcy + radius * inv_th * uy;
      L[n++] = ln;
      continue;
    }
    float dcut, dleft, dright;
    if (t < 0.0f || t > 1.0f || same) dcut = INFINITY;
    else { const float ax = vx - (lcx + t * cvx), ay = vy - (lcy + t * cvy); dcut = dot2(ax, ay, ax, ay); }
    if (tl < 0.0f) dleft = INFINITY;
    else { const float ax = vx - (lcx + tl * llx), ay = vy - (lcy + tl * lly); dleft = dot2(ax, ay, ax, ay); }
    if (tr < 0.0f) dright = INFINITY;
    else { const float ax = vx - (rcx + tr * rlx), ay = vy - (rcy + tr * rly); dright = dot2(ax, ay, ax, ay); }
    if (dcut <= dleft && dcut <= dright) { /* cut-off line */
      ln.dx = -a1->ux; ln.dy = -a1->uy;
      ln.px = lcx + radius * inv_th * -ln.dy; ln.py = lcy + radius * inv_th * ln.dx;
      L[n++] = ln;
      continue;
    } else if (dleft <= dright) { /* left leg */
      if (left_foreign) continue;
      ln.dx = llx; ln.dy = lly;
      ln.px = lcx + radius * inv_th * -ln.dy; ln.py = lcy + radius * inv_th * ln.dx;
      L[n++] = ln;
      continue;
    } else { /* right leg */
      if (right_foreign) continue;
      ln.dx = -rlx; ln.dy = -rly;
      ln.px = rcx + radius * inv_th * -ln.dy; ln.py = rcy + radius * inv_th * ln.dx;
      L[n++] = ln;
      continue;
    }
  }
  return n;
}

/* One agent's RVO2 step: computeNeighbors + computeNewVelocity (agent-agent part).
 * SURVEY Appendix A.1-A.4.  `others` are visited in index order (RVO2's kd-tree leaf
 * order for <= 10 agents; for more agents the order only matters for exact distance
 * ties). */
void ebc_ref_orca_agent(float px, float py, float vx, float vy, float radius, float max_speed,
                        float pref_x, float pref_y, int n_other, const float *opx,
                        const float *opy, const float *ovx, const float *ovy, const float *orad,
                        float neighbor_dist, int max_neighbors, float time_horizon,
                        float time_step, const void *obstacles, float time_horizon_obst, float *out_vx,
                        float *out_vy) {
  line_t L[MAX_LINES];
  /* obstacle ORCA lines come first (Agent::computeNewVelocity); `obstacles` = an obst_set or NULL */
  const int num_obst = obstacles ? obstacle_lines((const obst_set *)obstacles, px, py, vx, vy, radius, max_speed,
                                                  time_horizon_obst, L)
                                 : 0;
  int nb[MAX_LINES];
  float nbd[MAX_LINES];
  int cnt = 0;
  if (max_neighbors > MAX_LINES) max_neighbors = MAX_LINES;
  float range_sq = neighbor_dist * neighbor_dist;
  if (max_neighbors > 0) {
    for (int j = 0; j < n_other; ++j) { /* Agent::insertAgentNeighbor */
      const float ddx = px - opx[j], ddy = py - opy[j];
      const float d = dot2(ddx, ddy, ddx, ddy);
      if (d < range_sq) {
        if (cnt < max_neighbors) ++cnt;
        int i = cnt - 1;
        while (i != 0 && d < nbd[i - 1]) {
          nbd[i] = nbd[i - 1];
          nb[i] = nb[i - 1];
          --i;
        }
        nbd[i] = d;
        nb[i] = j;
        if (cnt == max_neighbors) range_sq = nbd[cnt - 1];
      }
    }
  }
  const float inv_th = 1.0f / time_horizon;
  if (cnt > MAX_LINES - num_obst) cnt = MAX_LINES - num_obst;
  for (int k = 0; k < cnt; ++k) {
    const int j = nb[k];
    const float rpx = opx[j] - px, rpy = opy[j] - py;
    const float rvx = vx - ovx[j], rvy = vy - ovy[j];
    const float dist_sq = dot2(rpx, rpy, rpx, rpy);
    const float R = radius + orad[j];
    const float R_sq = R * R;
    float ux, uy;
    line_t ln;
    if (dist_sq > R_sq) {
      const float wx = rvx - inv_th * rpx, wy = rvy - inv_th * rpy;
      const float wlen_sq = dot2(wx, wy, wx, wy);
      const float dp1 = dot2(wx, wy, rpx, rpy);
      if (dp1 < 0.0f && dp1 * dp1 > R_sq * wlen_sq) {
        const float wlen = sqrtf(wlen_sq);
        const float inv = 1.0f / wlen;
        const float uwx = wx * inv, uwy = wy * inv;
        ln.dx = uwy;
        ln.dy = -uwx;
        const float s = R * inv_th - wlen;
        ux = s * uwx;
        uy = s * uwy;
      } else {
        const float leg = sqrtf(dist_sq - R_sq);
        const float inv = 1.0f / dist_sq;
        if (det2(rpx, rpy, wx, wy) > 0.0f) {
          ln.dx = (rpx * leg - rpy * R) * inv;
          ln.dy = (rpx * R + rpy * leg) * inv;
        } else {
          ln.dx = (-(rpx * leg + rpy * R)) * inv;
          ln.dy = (-(-rpx * R + rpy * leg)) * inv;
        }
        const float dp2 = dot2(rvx, rvy, ln.dx, ln.dy);
        ux = dp2 * ln.dx - rvx;
        uy = dp2 * ln.dy - rvy;
      }
    } else {
      const float inv_ts = 1.0f / time_step;
      const float wx = rvx - inv_ts * rpx, wy = rvy - inv_ts * rpy;
      const float wlen = sqrtf(dot2(wx, wy, wx, wy));
      const float inv = 1.0f / wlen;
      const float uwx = wx * inv, uwy = wy * inv;
      ln.dx = uwy;
      ln.dy = -uwx;
      const float s = R * inv_ts - wlen;
      ux = s * uwx;
      uy = s * uwy;
    }
    ln.px = vx + 0.5f * ux;
    ln.py = vy + 0.5f * uy;
    L[num_obst + k] = ln;
  }
  float rx, ry;
  const int total = num_obst + cnt;
  const int fail = lp2(L, total, max_speed, pref_x, pref_y, 0, &rx, &ry);
  if (fail < total) lp3(L, total, num_obst, fail, max_speed, &rx, &ry);
  *out_vx = rx;
  *out_vy = ry;
}

/* ------------------------------------------------------------------------------------
 * Minimal stand-in for rvo2.PyRVOSimulator (Python-RVO2 src/rvo2.pyx), used through
 * oracle/shims/rvo2.py so that the unmodified reference can run here.
 * Call sites: simulator/policy/orca.py:105-154.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  float ts;
  int n, cap;
  float *px, *py, *vx, *vy, *prx, *pry, *rad, *maxs, *nd, *th, *tho;
  int *maxn;
  float oxy[2 * OBST_MAX]; /* addObstacle vertices, processed by processObstacles */
  int32_t opoly[OBST_MAX];
  int on, opolys, obst_ready;
  obst_set obst;
} rvo_sim;

void *ebc_rvo_create(float time_step) {
  rvo_sim *s = (rvo_sim *)calloc(1, sizeof(rvo_sim));
  s->ts = time_step;
  return s;
}
void ebc_rvo_destroy(void *h) {
  rvo_sim *s = (rvo_sim *)h;
  if (!s) return;
  free(s->px); free(s->py); free(s->vx); free(s->vy); free(s->prx); free(s->pry);
  free(s->rad); free(s->maxs); free(s->nd); free(s->th); free(s->maxn); free(s->tho);
  free(s);
}
int ebc_rvo_add_agent(void *h, float px, float py, float nd, int maxn, float th, float rad,
                      float maxs, float vx, float vy) {
  rvo_sim *s = (rvo_sim *)h;
  if (s->n == s->cap) {
    s->cap = s->cap ? 2 * s->cap : 16;
#define GROW(p, T) s->p = (T *)realloc(s->p, sizeof(T) * (size_t)s->cap)
    GROW(px, float); GROW(py, float); GROW(vx, float); GROW(vy, float); GROW(prx, float);
    GROW(pry, float); GROW(rad, float); GROW(maxs, float); GROW(nd, float); GROW(th, float);
    GROW(tho, float); GROW(maxn, int);
#undef GROW
  }
  const int i = s->n++;
  s->px[i] = px; s->py[i] = py; s->vx[i] = vx; s->vy[i] = vy; s->prx[i] = 0; s->pry[i] = 0;
  s->rad[i] = rad; s->maxs[i] = maxs; s->nd[i] = nd; s->th[i] = th; s->maxn[i] = maxn;
  s->tho[i] = th; /* timeHorizonObst: set separately by ebc_rvo_set_time_horizon_obst (orca*.py pass 5 for both) */
  return i;
}
void ebc_rvo_set_time_horizon_obst(void *h, int i, float tho) { ((rvo_sim *)h)->tho[i] = tho; }
/* RVOSimulator::addObstacle / processObstacles (orca_obstacles.py:105-107) */
int ebc_rvo_add_obstacle(void *h, const float *xy, int n) {
  rvo_sim *s = (rvo_sim *)h;
  if (n < 2 || s->on + n > OBST_MAX || s->opolys >= OBST_MAX) return -1;
  memcpy(s->oxy + 2 * s->on, xy, sizeof(float) * 2 * (size_t)n);
  s->on += n;
  s->opoly[s->opolys] = n;
  return s->opolys++;
}
int ebc_rvo_process_obstacles(void *h) {
  rvo_sim *s = (rvo_sim *)h;
  if (pack_obstacles(s->oxy, s->opoly, s->opolys, OBST_MAX, &s->obst)) return -1;
  s->obst_ready = 1;
  return 0;
}
int ebc_rvo_num_agents(void *h) { return ((rvo_sim *)h)->n; }
void ebc_rvo_set_pos(void *h, int i, float x, float y) { rvo_sim *s = h; s->px[i] = x; s->py[i] = y; }
void ebc_rvo_set_vel(void *h, int i, float x, float y) { rvo_sim *s = h; s->vx[i] = x; s->vy[i] = y; }
void ebc_rvo_set_pref(void *h, int i, float x, float y) { rvo_sim *s = h; s->prx[i] = x; s->pry[i] = y; }
void ebc_rvo_get_vel(void *h, int i, float *x, float *y) { rvo_sim *s = h; *x = s->vx[i]; *y = s->vy[i]; }
void ebc_rvo_get_pos(void *h, int i, float *x, float *y) { rvo_sim *s = h; *x = s->px[i]; *y = s->py[i]; }
/* RVOSimulator::doStep: new velocity for every agent, then update(). */
void ebc_rvo_do_step(void *h) {
  rvo_sim *s = (rvo_sim *)h;
  const int n = s->n;
  float *nvx = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1) * 7);
  float *nvy = nvx + n, *ox = nvy + n, *oy = ox + n, *ovx = oy + n, *ovy = ovx + n,
        *orad = ovy + n;
  for (int i = 0; i < n; ++i) {
    int m = 0;
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      ox[m] = s->px[j]; oy[m] = s->py[j]; ovx[m] = s->vx[j]; ovy[m] = s->vy[j];
      orad[m] = s->rad[j];
      ++m;
    }
    ebc_ref_orca_agent(s->px[i], s->py[i], s->vx[i], s->vy[i], s->rad[i], s->maxs[i], s->prx[i],
                       s->pry[i], m, ox, oy, ovx, ovy, orad, s->nd[i], s->maxn[i], s->th[i],
                       s->ts, s->obst_ready ? &s->obst : NULL, s->tho[i], &nvx[i], &nvy[i]);
  }
  for (int i = 0; i < n; ++i) {
    s->vx[i] = nvx[i];
    s->vy[i] = nvy[i];
    s->px[i] += s->vx[i] * s->ts;
    s->py[i] += s->vy[i] * s->ts;
  }
  free(nvx);
}

/* ------------------------------------------------------------------------------------
 * Batched CPU twin of the C ABI.
 * ---------------------------------------------------------------------------------- */
struct ebc_sim {
  ebc_config cfg;
  ebc_state st;
  ebc_stats stats; /* all-null when unbound */
  int bound;
  double *actions; /* A*2 */
  int have_actions;
  /* value net (copied) */
  int have_weights;
  float *attn_out; /* ebc_set_attention_output */
  int D, self_dim, with_global;
  struct { float *w, *b; int in, out; } lin[11]; /* mlp1[2] mlp2[2] att[3] mlp3[4] */
  char err[256];
};

static char g_create_err[256];
static int g_threads = 1;

void ebc_ref_set_threads(int n) { g_threads = n > 0 ? n : 1; }

static int fail(ebc_sim *s, int code, const char *msg) {
  snprintf(s ? s->err : g_create_err, 256, "%s", msg);
  return code;
}

int ebc_ref_create(const ebc_config *cfg, ebc_sim **out) {
  if (!cfg || !out) return fail(NULL, EBC_ERR_INVALID, "null argument");
  if (cfg->abi_version != EBC_ABI_VERSION) return fail(NULL, EBC_ERR_INVALID, "abi_version mismatch");
  if (cfg->n_episodes < 1 || cfg->max_humans < 1 || cfg->max_humans > 64 || cfg->max_statics < 0 ||
      cfg->max_humans + cfg->max_statics > 64 || cfg->max_rects < 0 || cfg->n_actions < 1 ||
      cfg->n_actions > 256 || cfg->orca_max_neighbors < 0 || cfg->orca_max_neighbors > 64 ||
      cfg->max_obst < 0 || cfg->max_obst > OBST_MAX || (cfg->orca_obstacles && !(cfg->orca_time_horizon_obst > 0.0f)))
    return fail(NULL, EBC_ERR_INVALID, "config out of range");
  ebc_sim *s = (ebc_sim *)calloc(1, sizeof(ebc_sim));
  if (!s) return fail(NULL, EBC_ERR_NOMEM, "calloc failed");
  s->cfg = *cfg;
  *out = s;
  return EBC_OK;
}

void ebc_ref_destroy(ebc_sim *s) {
  if (!s) return;
  free(s->actions);
  for (int i = 0; i < 11; ++i) { free(s->lin[i].w); free(s->lin[i].b); }
  free(s);
}

const char *ebc_ref_last_error(const ebc_sim *s) { return s ? s->err : g_create_err; }

int ebc_ref_bind(ebc_sim *s, const ebc_state *st) {
  if (!s || !st) return EBC_ERR_INVALID;
  if (!st->hum_pv || !st->hum_gr || !st->hum_type || !st->hum_count || !st->hum_nv ||
      !st->stat_count || !st->rect_count || !st->rob_pv || !st->rob_gr || !st->rob_theta ||
      !st->time)
    return fail(s, EBC_ERR_INVALID, "ebc_bind: null state array");
  s->st = *st;
  s->bound = 1;
  return EBC_OK;
}

/* Running statistics of the explorer's inner loop: rl/utils/explorer.py:33-94 (too_close, min_dist, outcome of the
 * last step), :189-193 (cumulative discounted reward), rl/test_parallel.py:52-130. */
int ebc_ref_bind_stats(ebc_sim *s, const ebc_stats *x) {
  if (!s) return EBC_ERR_INVALID;
  if (!x) { memset(&s->stats, 0, sizeof(s->stats)); return EBC_OK; }
  if (!x->alive || !x->final_event || !x->steps || !x->too_close || !x->cum_reward || !x->discount ||
      !x->min_dist_sum || !x->alive_count)
    return fail(s, EBC_ERR_INVALID, "ebc_bind_stats: null statistics array");
  s->stats = *x;
  return EBC_OK;
}

int ebc_ref_set_actions(ebc_sim *s, const double *a, int32_t n) {
  if (!s || !a || n != s->cfg.n_actions) return fail(s, EBC_ERR_INVALID, "ebc_set_actions: bad table");
  free(s->actions);
  s->actions = (double *)malloc(sizeof(double) * 2 * (size_t)n);
  memcpy(s->actions, a, sizeof(double) * 2 * (size_t)n);
  s->have_actions = 1;
  return EBC_OK;
}

int ebc_ref_set_weights(ebc_sim *s, const ebc_weights *w) {
  if (!s || !w) return EBC_ERR_INVALID;
  const int D = s->cfg.with_agent_type ? 17 : 13;
  if (w->input_dim != D) return fail(s, EBC_ERR_INVALID, "ebc_set_weights: input_dim != D");
  const ebc_linear *src[11] = {&w->mlp1[0], &w->mlp1[1], &w->mlp2[0], &w->mlp2[1],
                               &w->attention[0], &w->attention[1], &w->attention[2],
                               &w->mlp3[0], &w->mlp3[1], &w->mlp3[2], &w->mlp3[3]};
  for (int i = 0; i < 11; ++i) {
    free(s->lin[i].w); free(s->lin[i].b);
    const size_t nw = (size_t)src[i]->in_dim * (size_t)src[i]->out_dim;
    s->lin[i].w = (float *)malloc(sizeof(float) * nw);
    s->lin[i].b = (float *)malloc(sizeof(float) * (size_t)src[i]->out_dim);
    memcpy(s->lin[i].w, src[i]->weight, sizeof(float) * nw);
    memcpy(s->lin[i].b, src[i]->bias, sizeof(float) * (size_t)src[i]->out_dim);
    s->lin[i].in = src[i]->in_dim;
    s->lin[i].out = src[i]->out_dim;
  }
  s->D = D;
  s->self_dim = w->self_state_dim;
  s->with_global = w->with_global_state;
  /* shape chain (sarl.py:23-36) */
  const int h1 = s->lin[1].out;
  if (s->lin[0].in != D || s->lin[1].in != s->lin[0].out || s->lin[2].in != h1 ||
      s->lin[3].in != s->lin[2].out || s->lin[4].in != (s->with_global ? 2 * h1 : h1) ||
      s->lin[5].in != s->lin[4].out || s->lin[6].in != s->lin[5].out || s->lin[6].out != 1 ||
      s->lin[7].in != s->lin[3].out + s->self_dim || s->lin[8].in != s->lin[7].out ||
      s->lin[9].in != s->lin[8].out || s->lin[10].in != s->lin[9].out || s->lin[10].out != 1)
    return fail(s, EBC_ERR_INVALID, "ebc_set_weights: inconsistent layer shapes");
  s->have_weights = 1;
  return EBC_OK;
}

/* simulator/policy/orca.py:113-119,136-140: self radius / max speed / preferred velocity.
 * Python computes these in float64 and rvo2 narrows to float at the Cython boundary. */
static inline void orca_self_params(float px, float py, float gx, float gy, float radius,
                                    double safety, float *r_out, float *prefx, float *prefy) {
  *r_out = (float)((double)radius + 0.01 + safety);
  const double vx = (double)gx - (double)px, vy = (double)gy - (double)py;
  const double speed = sqrt(vx * vx + vy * vy);
  if (speed > 1.0) {
    *prefx = (float)(vx / speed);
    *prefy = (float)(vy / speed);
  } else {
    *prefx = (float)vx;
    *prefy = (float)vy;
  }
}

/* K1: simulator/env.py:392-405 + simulator/policy/orca.py:85-157 (+ linear.py:17-23). */
int ebc_ref_orca(ebc_sim *s) {
  if (!s || !s->bound) return fail(s, EBC_ERR_UNBOUND, "ebc_orca: state not bound");
  const ebc_config *c = &s->cfg;
  const int Hm = c->max_humans;
#pragma omp parallel for schedule(static) num_threads(g_threads)
  for (int e = 0; e < c->n_episodes; ++e) {
    const int H = s->st.hum_count[e];
    const float *pv = s->st.hum_pv + (size_t)e * Hm * 4;
    const float *gr = s->st.hum_gr + (size_t)e * Hm * 4;
    const uint8_t *ty = s->st.hum_type + (size_t)e * Hm;
    float *nv = s->st.hum_nv + (size_t)e * Hm * 2;
    float ox[65], oy[65], ovx[65], ovy[65], orad[65];
    obst_set obst;
    const int use_obst = c->orca_obstacles && s->st.obst && s->st.obst_count;
    if (use_obst) unpack_obstacles(s->st.obst + (size_t)e * c->max_obst, s->st.obst_count[e], &obst);
    for (int h = 0; h < H; ++h) {
      const float px = pv[h * 4], py = pv[h * 4 + 1], vx = pv[h * 4 + 2], vy = pv[h * 4 + 3];
      const float gx = gr[h * 4], gy = gr[h * 4 + 1], vpref = gr[h * 4 + 2], rad = gr[h * 4 + 3];
      if (c->human_policy[ty[h] < 3 ? ty[h] : 0] == EBC_POLICY_LINEAR) {
        /* linear.py:17-23, float64 then stored as fp32 state */
        const double th = atan2((double)gy - (double)py, (double)gx - (double)px);
        nv[h * 2] = (float)(cos(th) * (double)vpref);
        nv[h * 2 + 1] = (float)(sin(th) * (double)vpref);
        continue;
      }
      int m = 0;
      for (int j = 0; j < H; ++j) {
        if (j == h) continue;
        ox[m] = pv[j * 4]; oy[m] = pv[j * 4 + 1]; ovx[m] = pv[j * 4 + 2]; ovy[m] = pv[j * 4 + 3];
        orad[m] = (float)((double)gr[j * 4 + 3] + 0.01 + c->orca_safety_space);
        ++m;
      }
      if (c->robot_visible) { /* env.py:401-402: robot appended last */
        const float *rp = s->st.rob_pv + (size_t)e * 4;
        ox[m] = rp[0]; oy[m] = rp[1]; ovx[m] = rp[2]; ovy[m] = rp[3];
        orad[m] = (float)((double)s->st.rob_gr[(size_t)e * 4 + 3] + 0.01 + c->orca_safety_space);
        ++m;
      }
      float r_self, prefx, prefy;
      orca_self_params(px, py, gx, gy, rad, c->orca_safety_space, &r_self, &prefx, &prefy);
      ebc_ref_orca_agent(px, py, vx, vy, r_self, vpref, prefx, prefy, m, ox, oy, ovx, ovy, orad,
                         c->orca_neighbor_dist, c->orca_max_neighbors, c->orca_time_horizon,
                         (float)c->time_step, use_obst ? &obst : NULL, c->orca_time_horizon_obst,
                         &nv[h * 2], &nv[h * 2 + 1]);
    }
  }
  return EBC_OK;
}

/* Robot as ORCA agent 0 (imitation learning): rl/train.py:99-143 + orca.py:85-157 with
 * ob = humans + static discs (env.py:188-193). */
int ebc_ref_robot_orca(ebc_sim *s, double safety, double *out) {
  if (!s || !s->bound) return fail(s, EBC_ERR_UNBOUND, "ebc_robot_orca: state not bound");
  const ebc_config *c = &s->cfg;
  const int Hm = c->max_humans, Sm = c->max_statics;
#pragma omp parallel for schedule(static) num_threads(g_threads)
  for (int e = 0; e < c->n_episodes; ++e) {
    const int H = s->st.hum_count[e], S = s->st.stat_count[e];
    const float *pv = s->st.hum_pv + (size_t)e * Hm * 4;
    const float *gr = s->st.hum_gr + (size_t)e * Hm * 4;
    float ox[65], oy[65], ovx[65], ovy[65], orad[65];
    obst_set obst;
    const int use_obst = c->orca_obstacles && s->st.obst && s->st.obst_count;
    if (use_obst) unpack_obstacles(s->st.obst + (size_t)e * c->max_obst, s->st.obst_count[e], &obst);
    int m = 0;
    for (int j = 0; j < H; ++j) {
      ox[m] = pv[j * 4]; oy[m] = pv[j * 4 + 1]; ovx[m] = pv[j * 4 + 2]; ovy[m] = pv[j * 4 + 3];
      orad[m] = (float)((double)gr[j * 4 + 3] + 0.01 + safety);
      ++m;
    }
    for (int k = 0; k < S; ++k) {
      const float *sd = s->st.stat + ((size_t)e * Sm + k) * 4;
      ox[m] = sd[0]; oy[m] = sd[1]; ovx[m] = 0.0f; ovy[m] = 0.0f;
      orad[m] = (float)((double)sd[2] + 0.01 + safety);
      ++m;
    }
    const float *rp = s->st.rob_pv + (size_t)e * 4;
    const float *rg = s->st.rob_gr + (size_t)e * 4;
    float r_self, prefx, prefy, vx, vy;
    orca_self_params(rp[0], rp[1], rg[0], rg[1], rg[3], safety, &r_self, &prefx, &prefy);
    ebc_ref_orca_agent(rp[0], rp[1], rp[2], rp[3], r_self, rg[2], prefx, prefy, m, ox, oy, ovx, ovy,
                       orad, c->orca_neighbor_dist, c->orca_max_neighbors, c->orca_time_horizon,
                       (float)c->time_step, use_obst ? &obst : NULL, c->orca_time_horizon_obst, &vx, &vy);
    out[(size_t)e * 2] = (double)vx;
    out[(size_t)e * 2 + 1] = (double)vy;
  }
  return EBC_OK;
}

/* simulator/utils/collisions.py:4-26 with (x3, y3) = (0, 0). */
static inline double point_to_segment_dist0(double x1, double y1, double x2, double y2) {
  const double px = x2 - x1, py = y2 - y1;
  if (px == 0.0 && py == 0.0) return sqrt((0.0 - x1) * (0.0 - x1) + (0.0 - y1) * (0.0 - y1));
  double u = ((0.0 - x1) * px + (0.0 - y1) * py) / (px * px + py * py);
  if (u > 1.0) u = 1.0;
  else if (u < 0.0) u = 0.0;
  const double x = x1 + u * px, y = y1 + u * py;
  return sqrt((x - 0.0) * (x - 0.0) + (y - 0.0) * (y - 0.0));
}

typedef struct {
  double dmin[3];
  int coll[3];
  int coll_obst;
  double end_x, end_y; /* robot next position (agent.py:164-188) */
  double dist_to_goal;
  double reward;
  double min_dist; /* Danger.min_dist (reward.py:138-166) */
  int done, event;
} outcome_t;

/* env.py:424-444: compute_collisions + Reward.compute for one (episode, action).
 * a0, a1 = (vx, vy) holonomic or (v, r) otherwise. */
static void evaluate_action(const ebc_sim *s, int e, double a0, double a1, outcome_t *o) {
  const ebc_config *c = &s->cfg;
  const int Hm = c->max_humans;
  const int H = s->st.hum_count[e];
  const float *pv = s->st.hum_pv + (size_t)e * Hm * 4;
  const float *gr = s->st.hum_gr + (size_t)e * Hm * 4;
  const uint8_t *ty = s->st.hum_type + (size_t)e * Hm;
  const float *rp = s->st.rob_pv + (size_t)e * 4;
  const float *rg = s->st.rob_gr + (size_t)e * 4;
  const double rpx = rp[0], rpy = rp[1], rrad = rg[3], theta = s->st.rob_theta[e];
  const double dt = c->time_step;
  /* robot velocity implied by the action (collisions.py:37-42) and next position
   * (agent.py:164-176) */
  double avx, avy;
  if (c->robot_kinematics == EBC_KIN_HOLONOMIC) {
    avx = a0;
    avy = a1;
    o->end_x = rpx + a0 * dt;
    o->end_y = rpy + a1 * dt;
  } else {
    const double th = theta + a1;
    avx = a0 * cos(a1 + theta);
    avy = a0 * sin(a1 + theta);
    o->end_x = rpx + cos(th) * a0 * dt;
    o->end_y = rpy + sin(th) * a0 * dt;
  }
  /* env.py:303-338: per type, list order, stop at the first collision */
  for (int t = 0; t < 3; ++t) { o->dmin[t] = INFINITY; o->coll[t] = 0; }
  for (int h = 0; h < H; ++h) {
    const int t = ty[h];
    if (t > 2 || o->coll[t]) continue;
    const double px = (double)pv[h * 4] - rpx, py = (double)pv[h * 4 + 1] - rpy;
    const double vx = (double)pv[h * 4 + 2] - avx, vy = (double)pv[h * 4 + 3] - avy;
    const double ex = px + vx * dt, ey = py + vy * dt;
    const double closest = point_to_segment_dist0(px, py, ex, ey) - (double)gr[h * 4 + 3] - rrad;
    if (closest < 0.0) o->coll[t] = 1;
    else if (closest < o->dmin[t]) o->dmin[t] = closest;
  }
  /* env.py:227-261: occupancy-grid window test == integer AABB overlap against the
   * zero-cell rectangles of scene.map */
  o->coll_obst = 0;
  {
    const int ix = (int)rint((o->end_x + c->map_size_m / 2.0) / c->map_resolution);
    const int iy = (int)rint((o->end_y + c->map_size_m / 2.0) / c->map_resolution);
    const int sz = (int)ceil(rrad / sqrt(2.0) / c->map_resolution);
    const int G = (int)rint(c->map_size_m / c->map_resolution);
    int sx = ix - sz, ex = sx + sz * 2, sy = iy - sz, ey = sy + sz * 2;
    if (sx < 0) sx = 0;
    if (ex > G) ex = G;
    if (sy < 0) sy = 0;
    if (ey > G) ey = G;
    if (ex > sx && ey > sy) {
      const int R = s->st.rect_count[e];
      const int16_t *rc = s->st.rect + (size_t)e * c->max_rects * 4;
      for (int j = 0; j < R; ++j)
        if (sx < rc[j * 4 + 2] && rc[j * 4] < ex && sy < rc[j * 4 + 3] && rc[j * 4 + 1] < ey) {
          o->coll_obst = 1;
          break;
        }
    }
  }
  /* reward.py:80-181 */
  const double gdx = o->end_x - (double)rg[0], gdy = o->end_y - (double)rg[1];
  const double dist = sqrt(gdx * gdx + gdy * gdy);
  o->dist_to_goal = dist;
  const int reaching = dist < rrad;
  const double goal_reward = c->has_max_goal_distance ? 1.0 - dist / c->max_goal_distance : 0.0;
  double reward = c->new_reward ? goal_reward : 0.0;
  const double gt = s->st.time[e];
  int done, ev;
  o->min_dist = 0.0;
  if (gt >= c->time_limit) { done = 1; ev = EBC_EV_TIMEOUT; }
  else if (o->coll[EBC_CHILD]) { reward += c->collision_penalty_child; done = 1; ev = EBC_EV_COLLISION_CHILD; }
  else if (o->coll[EBC_BICYCLE]) { reward += c->collision_penalty_bicycle; done = 1; ev = EBC_EV_COLLISION_BICYCLE; }
  else if (o->coll[EBC_ADULT]) { reward += c->collision_penalty_adult; done = 1; ev = EBC_EV_COLLISION_ADULT; }
  else if (o->coll_obst) { reward += c->collision_penalty_obstacle; done = 1; ev = EBC_EV_COLLISION_OBSTACLE; }
  else if (reaching) {
    if (c->new_reward) {
      double tr; /* reward.py:8-14 */
      if (gt < c->time_good) tr = 1.0;
      else if (gt <= c->time_max) tr = (c->time_max - gt) / (c->time_max - c->time_good);
      else tr = 0.0;
      reward += tr;
    } else {
      reward += c->success_reward;
    }
    done = 1; ev = EBC_EV_REACH_GOAL;
  } else if (o->dmin[EBC_CHILD] < c->discomfort_dist_child) {
    reward = (o->dmin[EBC_CHILD] - c->discomfort_dist_child) * c->discomfort_penalty_factor_child * dt;
    o->min_dist = o->dmin[EBC_CHILD];
    done = 0; ev = EBC_EV_DANGER;
  } else if (o->dmin[EBC_BICYCLE] < c->discomfort_dist_bicycle) {
    reward = (o->dmin[EBC_BICYCLE] - c->discomfort_dist_bicycle) * c->discomfort_penalty_factor_bicycle * dt;
    o->min_dist = o->dmin[EBC_BICYCLE];
    done = 0; ev = EBC_EV_DANGER;
  } else if (o->dmin[EBC_ADULT] < c->discomfort_dist_adult) {
    reward = (o->dmin[EBC_ADULT] - c->discomfort_dist_adult) * c->discomfort_penalty_factor_adult * dt;
    o->min_dist = o->dmin[EBC_ADULT];
    done = 0; ev = EBC_EV_DANGER;
  } else if (c->robot_kinematics != EBC_KIN_HOLONOMIC && fabs(a1) > 0.0 && c->rotation_penalty_factor != 0.0) {
    reward = fabs(a1) * c->rotation_penalty_factor;
    done = 0; ev = EBC_EV_NOTHING;
  } else { reward = 0.0; done = 0; ev = EBC_EV_NOTHING; }
  o->reward = reward; o->done = done; o->event = ev;
}

/* rl/policy/cadrl.py:236-337 in fp32, one row.  js = the 15-tuple of state.py:19-30,65-66
 * already narrowed to fp32 (torch.Tensor([...]) at multi_human_rl.py:52-60). */
static void rotate_row(const float *js, int rotate_theta, int with_type, float *out) {
  const float dx = js[5] - js[0], dy = js[6] - js[1];
  const float rot = atan2f(js[6] - js[1], js[5] - js[0]);
  const float cr = cosf(rot), sr = sinf(rot);
  out[0] = sqrtf(dx * dx + dy * dy);                 /* dg */
  out[1] = js[7];                                    /* v_pref */
  out[2] = rotate_theta ? js[8] - rot : 0.0f;        /* theta */
  out[3] = js[4];                                    /* radius */
  out[4] = js[2] * cr + js[3] * sr;                  /* vx */
  out[5] = js[3] * cr - js[2] * sr;                  /* vy */
  out[6] = (js[9] - js[0]) * cr + (js[10] - js[1]) * sr;  /* px1 */
  out[7] = (js[10] - js[1]) * cr - (js[9] - js[0]) * sr;  /* py1 */
  out[8] = js[11] * cr + js[12] * sr;                /* vx1 */
  out[9] = js[12] * cr - js[11] * sr;                /* vy1 */
  out[10] = js[13];                                  /* radius1 */
  const float ax = js[0] - js[9], ay = js[1] - js[10];
  out[11] = sqrtf(ax * ax + ay * ay);                /* da */
  out[12] = js[4] + js[13];                          /* radius sum */
  if (with_type) {
    const int t = (int)js[14];
    for (int k = 0; k < 4; ++k) out[13 + k] = (k == t) ? 1.0f : 0.0f;
  }
}

/* K3: rl/policy/multi_human_rl.py:38-61 + env.onestep_lookahead (env.py:207-209). */
int ebc_ref_lookahead(ebc_sim *s, float *vin, double *reward, uint8_t *done, uint8_t *event) {
  if (!s || !s->bound) return fail(s, EBC_ERR_UNBOUND, "ebc_lookahead: state not bound");
  if (!s->have_actions) return fail(s, EBC_ERR_UNBOUND, "ebc_lookahead: action table not set");
  const ebc_config *c = &s->cfg;
  const int Hm = c->max_humans, Sm = c->max_statics, A = c->n_actions;
  const int n = Hm + Sm, D = c->with_agent_type ? 17 : 13;
  const double dt = c->time_step;
#pragma omp parallel for schedule(static) num_threads(g_threads)
  for (int e = 0; e < c->n_episodes; ++e) {
    const int H = s->st.hum_count[e], S = s->st.stat_count[e];
    const float *pv = s->st.hum_pv + (size_t)e * Hm * 4;
    const float *gr = s->st.hum_gr + (size_t)e * Hm * 4;
    const float *nv = s->st.hum_nv + (size_t)e * Hm * 2;
    const uint8_t *ty = s->st.hum_type + (size_t)e * Hm;
    const float *rp = s->st.rob_pv + (size_t)e * 4;
    const float *rg = s->st.rob_gr + (size_t)e * 4;
    const double theta = s->st.rob_theta[e];
    for (int a = 0; a < A; ++a) {
      const double a0 = s->actions[a * 2], a1 = s->actions[a * 2 + 1];
      outcome_t o;
      evaluate_action(s, e, a0, a1, &o);
      const size_t ea = (size_t)e * A + a;
      if (reward) reward[ea] = o.reward;
      if (done) done[ea] = (uint8_t)o.done;
      if (event) event[ea] = (uint8_t)o.event;
      if (!vin) continue;
      /* cadrl.py:118-165 propagate(robot FullState) in float64, narrowed by torch.Tensor */
      float js[15];
      if (c->robot_kinematics == EBC_KIN_HOLONOMIC) {
        js[0] = (float)((double)rp[0] + a0 * dt);
        js[1] = (float)((double)rp[1] + a1 * dt);
        js[2] = (float)a0;
        js[3] = (float)a1;
        js[8] = (float)theta;
      } else {
        const double nth = theta + a1;
        const double nvx = a0 * cos(nth), nvy = a0 * sin(nth);
        js[0] = (float)((double)rp[0] + nvx * dt);
        js[1] = (float)((double)rp[1] + nvy * dt);
        js[2] = (float)nvx;
        js[3] = (float)nvy;
        js[8] = (float)nth;
      }
      js[4] = rg[3]; js[5] = rg[0]; js[6] = rg[1]; js[7] = rg[2];
      float *rows = vin + ea * (size_t)n * D;
      memset(rows, 0, sizeof(float) * (size_t)n * D);
      for (int h = 0; h < H; ++h) { /* agent.py:80-93 get_next_observable_state(orca action) */
        js[9] = (float)((double)pv[h * 4] + (double)nv[h * 2] * dt);
        js[10] = (float)((double)pv[h * 4 + 1] + (double)nv[h * 2 + 1] * dt);
        js[11] = nv[h * 2];
        js[12] = nv[h * 2 + 1];
        js[13] = gr[h * 4 + 3];
        js[14] = (float)ty[h];
        rotate_row(js, c->rotate_theta, c->with_agent_type, rows + (size_t)h * D);
      }
      for (int k = 0; k < S; ++k) { /* env.py:457-458 static discs, type 3 */
        const float *sd = s->st.stat + ((size_t)e * Sm + k) * 4;
        js[9] = sd[0]; js[10] = sd[1]; js[11] = 0.0f; js[12] = 0.0f; js[13] = sd[2];
        js[14] = (float)EBC_ADULT_STATIC;
        rotate_row(js, c->rotate_theta, c->with_agent_type, rows + (size_t)(H + k) * D);
      }
    }
  }
  return EBC_OK;
}

/* rl/policy/multi_human_rl.py:128-149: rotate(current joint state). */
int ebc_ref_transform(ebc_sim *s, float *out) {
  if (!s || !s->bound) return fail(s, EBC_ERR_UNBOUND, "ebc_transform: state not bound");
  const ebc_config *c = &s->cfg;
  const int Hm = c->max_humans, Sm = c->max_statics;
  const int n = Hm + Sm, D = c->with_agent_type ? 17 : 13;
#pragma omp parallel for schedule(static) num_threads(g_threads)
  for (int e = 0; e < c->n_episodes; ++e) {
    const int H = s->st.hum_count[e], S = s->st.stat_count[e];
    const float *pv = s->st.hum_pv + (size_t)e * Hm * 4;
    const float *gr = s->st.hum_gr + (size_t)e * Hm * 4;
    const uint8_t *ty = s->st.hum_type + (size_t)e * Hm;
    const float *rp = s->st.rob_pv + (size_t)e * 4;
    const float *rg = s->st.rob_gr + (size_t)e * 4;
    float js[15];
    js[0] = rp[0]; js[1] = rp[1]; js[2] = rp[2]; js[3] = rp[3];
    js[4] = rg[3]; js[5] = rg[0]; js[6] = rg[1]; js[7] = rg[2]; js[8] = s->st.rob_theta[e];
    float *rows = out + (size_t)e * n * D;
    memset(rows, 0, sizeof(float) * (size_t)n * D);
    for (int h = 0; h < H; ++h) {
      js[9] = pv[h * 4]; js[10] = pv[h * 4 + 1]; js[11] = pv[h * 4 + 2]; js[12] = pv[h * 4 + 3];
      js[13] = gr[h * 4 + 3]; js[14] = (float)ty[h];
      rotate_row(js, c->rotate_theta, c->with_agent_type, rows + (size_t)h * D);
    }
    for (int k = 0; k < S; ++k) {
      const float *sd = s->st.stat + ((size_t)e * Sm + k) * 4;
      js[9] = sd[0]; js[10] = sd[1]; js[11] = 0.0f; js[12] = 0.0f; js[13] = sd[2];
      js[14] = (float)EBC_ADULT_STATIC;
      rotate_row(js, c->rotate_theta, c->with_agent_type, rows + (size_t)(H + k) * D);
    }
  }
  return EBC_OK;
}

/* nn.Linear: y = W x + b, fp32 in / fp32 out, accumulated in fp64 and rounded once (the
 * correctly-rounded result the reference's MKL sgemm approximates). */
static void linear_f(const float *w, const float *b, int in, int out, const float *x, float *y,
                     int relu) {
  for (int o = 0; o < out; ++o) {
    double acc = (double)b[o];
    const float *wr = w + (size_t)o * in;
    for (int k = 0; k < in; ++k) acc += (double)wr[k] * (double)x[k];
    float v = (float)acc;
    y[o] = (relu && v < 0.0f) ? 0.0f : v;
  }
}

/* rl/policy/sarl.py:38-82, one state of `rows` real rows. */
static float value_forward(const ebc_sim *s, const float *x, int rows, int stride, float *attn) {
  const int D = s->D;
  const int d1a = s->lin[0].out, h1d = s->lin[1].out, d2a = s->lin[2].out, h2d = s->lin[3].out;
  const int a1d = s->lin[4].out, a2d = s->lin[5].out;
  float *h1 = (float *)malloc(sizeof(float) * (size_t)rows * (h1d + h2d + 1) + sizeof(float) * 4096);
  float *h2 = h1 + (size_t)rows * h1d;
  float *sc = h2 + (size_t)rows * h2d;
  float *tmp = sc + rows; /* scratch >= 4096 floats */
  float *tmp2 = tmp + 1024, *att_in = tmp + 2048;
  for (int r = 0; r < rows; ++r) {
    linear_f(s->lin[0].w, s->lin[0].b, D, d1a, x + (size_t)r * stride, tmp, 1);
    linear_f(s->lin[1].w, s->lin[1].b, d1a, h1d, tmp, h1 + (size_t)r * h1d, 1); /* last_relu */
    linear_f(s->lin[2].w, s->lin[2].b, h1d, d2a, h1 + (size_t)r * h1d, tmp, 1);
    linear_f(s->lin[3].w, s->lin[3].b, d2a, h2d, tmp, h2 + (size_t)r * h2d, 0);
  }
  float *g = tmp2 + 512;
  if (s->with_global) { /* sarl.py:51-60 torch.mean over the entity dimension */
    for (int k = 0; k < h1d; ++k) {
      double acc = 0.0;
      for (int r = 0; r < rows; ++r) acc += (double)h1[(size_t)r * h1d + k];
      g[k] = (float)(acc / (double)rows);
    }
  }
  for (int r = 0; r < rows; ++r) {
    memcpy(att_in, h1 + (size_t)r * h1d, sizeof(float) * h1d);
    if (s->with_global) memcpy(att_in + h1d, g, sizeof(float) * h1d);
    linear_f(s->lin[4].w, s->lin[4].b, s->lin[4].in, a1d, att_in, tmp, 1);
    linear_f(s->lin[5].w, s->lin[5].b, a1d, a2d, tmp, tmp2, 1);
    linear_f(s->lin[6].w, s->lin[6].b, a2d, 1, tmp2, &sc[r], 0);
  }
  /* sarl.py:69-70 masked softmax without max subtraction */
  double sum = 0.0;
  for (int r = 0; r < rows; ++r) {
    const float e = (sc[r] != 0.0f) ? expf(sc[r]) : 0.0f;
    sc[r] = e;
    sum += (double)e;
  }
  const float fsum = (float)sum;
  if (attn)                          /* sarl.py:71 attention_weights = weights[0, :, 0] */
    for (int r = 0; r < rows; ++r) attn[r] = sc[r] / fsum;
  float *joint = tmp; /* [self_dim + h2d] */
  for (int k = 0; k < s->self_dim; ++k) joint[k] = x[k];
  for (int k = 0; k < h2d; ++k) {
    double acc = 0.0;
    for (int r = 0; r < rows; ++r) acc += (double)(sc[r] / fsum) * (double)h2[(size_t)r * h2d + k];
    joint[s->self_dim + k] = (float)acc;
  }
  float *m1 = tmp2, *m2 = att_in, *m3 = tmp2 + 512;
  float v;
  linear_f(s->lin[7].w, s->lin[7].b, s->lin[7].in, s->lin[7].out, joint, m1, 1);
  linear_f(s->lin[8].w, s->lin[8].b, s->lin[8].in, s->lin[8].out, m1, m2, 1);
  linear_f(s->lin[9].w, s->lin[9].b, s->lin[9].in, s->lin[9].out, m2, m3, 1);
  linear_f(s->lin[10].w, s->lin[10].b, s->lin[10].in, 1, m3, &v, 0);
  free(h1);
  return v;
}

int ebc_ref_set_attention_output(ebc_sim *s, float *attn) {
  if (!s) return EBC_ERR_INVALID;
  s->attn_out = attn;
  return EBC_OK;
}

/* K4 */
int ebc_ref_value(ebc_sim *s, const float *vin, int64_t n_states, const int32_t *row_count,
                  float *values) {
  if (!s || !s->have_weights) return fail(s, EBC_ERR_UNBOUND, "ebc_value: weights not set");
  const ebc_config *c = &s->cfg;
  for (int i = 0; i < 11; ++i)
    if (s->lin[i].out > 512 || s->lin[i].in > 1024) return fail(s, EBC_ERR_INVALID, "ebc_value: layer too wide for the oracle scratch");
  const int n = c->max_humans + c->max_statics, D = s->D, A = c->n_actions;
  if (!row_count && !s->bound) return fail(s, EBC_ERR_UNBOUND, "ebc_value: state not bound");
#pragma omp parallel for schedule(dynamic, 16) num_threads(g_threads)
  for (int64_t i = 0; i < n_states; ++i) {
    int rows;
    if (row_count) rows = row_count[i];
    else {
      const int64_t e = i / A;
      rows = s->st.hum_count[e] + s->st.stat_count[e];
    }
    float *attn = s->attn_out ? s->attn_out + (size_t)i * n : NULL;
    if (attn) memset(attn, 0, sizeof(float) * (size_t)n);
    values[i] = value_forward(s, vin + (size_t)i * n * D, rows, D, attn);
  }
  return EBC_OK;
}

/* K5: multi_human_rl.py:25-26,72-82. */
int ebc_ref_select(ebc_sim *s, const double *reward, const float *values, double *action_values,
                   int32_t *argmax, uint8_t *nan_flag) {
  if (!s || !s->bound) return fail(s, EBC_ERR_UNBOUND, "ebc_select: state not bound");
  const ebc_config *c = &s->cfg;
  const int A = c->n_actions;
  for (int e = 0; e < c->n_episodes; ++e) {
    const float *rp = s->st.rob_pv + (size_t)e * 4;
    const float *rg = s->st.rob_gr + (size_t)e * 4;
    const double disc = pow(c->gamma, c->time_step * (double)rg[2]);
    double best = -INFINITY;
    int arg = -1;
    for (int a = 0; a < A; ++a) {
      const double v = reward[(size_t)e * A + a] + disc * (double)values[(size_t)e * A + a];
      if (action_values) action_values[(size_t)e * A + a] = v;
      if (v > best) { best = v; arg = a; }
    }
    if (nan_flag) nan_flag[e] = (uint8_t)(arg < 0);
    if (arg < 0) arg = 0;
    /* policy.py:43-54 reach_destination -> zero action */
    const double dy = (double)rp[1] - (double)rg[1], dx = (double)rp[0] - (double)rg[0];
    if (sqrt(dy * dy + dx * dx) < (double)rg[3]) arg = 0;
    argmax[e] = arg;
  }
  return EBC_OK;
}

/* K2: env.step(update=True), env.py:388-466 + compute_step_update :340-386 + agent.py:202-228. */
int ebc_ref_step(ebc_sim *s, const int32_t *action_idx, const double *action, const uint8_t *active,
                 double *reward, uint8_t *done, uint8_t *event, double *dmin, double *dist_to_goal) {
  if (!s || !s->bound) return fail(s, EBC_ERR_UNBOUND, "ebc_step: state not bound");
  if ((action_idx == NULL) == (action == NULL)) return fail(s, EBC_ERR_INVALID, "ebc_step: give exactly one of action_idx / action");
  if (action_idx && !s->have_actions) return fail(s, EBC_ERR_UNBOUND, "ebc_step: action table not set");
  const ebc_config *c = &s->cfg;
  const int Hm = c->max_humans;
  const double dt = c->time_step;
  const ebc_stats *sx = &s->stats;
  if (!active && sx->alive) active = sx->alive;
  int finished = 0;
#pragma omp parallel for schedule(static) num_threads(g_threads) reduction(+ : finished)
  for (int e = 0; e < c->n_episodes; ++e) {
    if (active && !active[e]) continue;
    double a0, a1;
    if (action_idx) {
      int ai = action_idx[e];
      if (ai < 0 || ai >= c->n_actions) ai = 0;
      a0 = s->actions[ai * 2]; a1 = s->actions[ai * 2 + 1];
    } else { a0 = action[(size_t)e * 2]; a1 = action[(size_t)e * 2 + 1]; }
    outcome_t o;
    evaluate_action(s, e, a0, a1, &o);
    if (reward) reward[e] = o.reward;
    if (done) done[e] = (uint8_t)o.done;
    if (event) event[e] = (uint8_t)o.event;
    if (dmin) { dmin[(size_t)e * 3] = o.dmin[0]; dmin[(size_t)e * 3 + 1] = o.dmin[1]; dmin[(size_t)e * 3 + 2] = o.dmin[2]; }
    if (dist_to_goal) dist_to_goal[e] = o.dist_to_goal;
    /* robot.step(action): agent.py:202-228 */
    float *rp = s->st.rob_pv + (size_t)e * 4;
    rp[0] = (float)o.end_x;
    rp[1] = (float)o.end_y;
    if (c->robot_kinematics == EBC_KIN_HOLONOMIC) {
      rp[2] = (float)a0;
      rp[3] = (float)a1;
    } else {
      const double th = fmod((double)s->st.rob_theta[e] + a1, 2.0 * M_PI);
      const double thw = th < 0.0 ? th + 2.0 * M_PI : th; /* Python % is non-negative */
      s->st.rob_theta[e] = (float)thw;
      rp[2] = (float)(a0 * cos(thw));
      rp[3] = (float)(a0 * sin(thw));
    }
    /* humans: agent.step(ActionXY) holonomic */
    const int H = s->st.hum_count[e];
    float *pv = s->st.hum_pv + (size_t)e * Hm * 4;
    const float *nv = s->st.hum_nv + (size_t)e * Hm * 2;
    for (int h = 0; h < H; ++h) {
      pv[h * 4] = (float)((double)pv[h * 4] + (double)nv[h * 2] * dt);
      pv[h * 4 + 1] = (float)((double)pv[h * 4 + 1] + (double)nv[h * 2 + 1] * dt);
      pv[h * 4 + 2] = nv[h * 2];
      pv[h * 4 + 3] = nv[h * 2 + 1];
    }
    s->st.time[e] += dt;
    if (sx->alive) {
      const double disc = sx->discount[e];
      sx->cum_reward[e] += disc * o.reward; /* explorer.py:189-193 */
      sx->discount[e] = disc * pow(c->gamma, dt * (double)s->st.rob_gr[(size_t)e * 4 + 2]);
      sx->steps[e] += 1;
      if (o.event == EBC_EV_DANGER) { /* explorer.py:47-49 */
        sx->too_close[e] += 1;
        sx->min_dist_sum[e] += o.min_dist;
      }
      if (o.done) {
        sx->final_event[e] = (uint8_t)o.event;
        sx->alive[e] = 0;
        finished += 1;
      }
    }
  }
  if (sx->alive) sx->alive_count[0] -= finished;
  return EBC_OK;
}

/* env.reset from a scene pool: simulator/env.py:128-205 (state hand-over only). */
/* ---- angular local map --------------------------------------------------------------------
 * simulator/env.py:468-568 calculate_angular_map_distances: one vertex seen from one robot corner ("edge"), then
 * the outline between this vertex and every earlier one of the same sweep, interpolated sector by sector. */
typedef struct { double max_range, min_angle, max_angle, res; int dim; } amap_t;

static void amap_calc(const amap_t *m, const double vertex[2], const double edge[2], double cs, double sn, double *vec,
                      int *rad_indeces, double (*locations)[2], int *n_seen) {
  double px = (vertex[0] - edge[0]) * cs + (vertex[1] - edge[1]) * sn;              /* env.py:474-479 */
  double py = (vertex[1] - edge[1]) * cs - (vertex[0] - edge[0]) * sn;
  const double phi = atan2(py, px);
  const int rad_idx = (int)((phi - m->min_angle) / m->res);                         /* int(): truncation */
  const double distance = sqrt(px * px + py * py);
  if (rad_idx >= 0 && rad_idx < m->dim && distance < vec[rad_idx]) vec[rad_idx] = distance;
  for (int k = 0; k < *n_seen; ++k) {                                                /* env.py:488-566 */
    const int old = rad_indeces[k];
    const double *loc = locations[k];
    int wrapped, idx_diff;
    if ((double)abs(rad_idx - old) > M_PI / m->res) {
      wrapped = 1;
      idx_diff = rad_idx > old ? m->dim - rad_idx + old : m->dim - old + rad_idx;
    } else {
      wrapped = 0;
      idx_diff = abs(rad_idx - old);
    }
    for (int i = 0; i < idx_diff; ++i) {
      const double t = (double)i / (double)idx_diff;
      if ((rad_idx < old && !wrapped) || (rad_idx > old && wrapped)) {
        if (rad_idx + i >= 0 && rad_idx + i < m->dim) {
          const double qx = vertex[0] + t * (loc[0] - vertex[0]) - edge[0];
          const double qy = vertex[1] + t * (loc[1] - vertex[1]) - edge[1];
          px = qx * cs + qy * sn;
          py = qy * cs - qx * sn;
          const double d = sqrt(px * px + py * py);
          if (d < vec[rad_idx + i]) vec[rad_idx + i] = d;
        }
      } else {
        if (old + i >= 0 && old + i < m->dim) {
          const double qx = loc[0] + t * (vertex[0] - loc[0]) - edge[0];
          const double qy = loc[1] + t * (vertex[1] - loc[1]) - edge[1];
          px = qx * cs + qy * sn;
          py = qy * cs - qx * sn;
          const double d = sqrt(px * px + py * py);
          if (d < vec[old + i]) vec[old + i] = d;
        }
      }
    }
  }
  rad_indeces[*n_seen] = rad_idx;
  locations[*n_seen][0] = vertex[0];
  locations[*n_seen][1] = vertex[1];
  ++*n_seen;
}

/* simulator/env.py:570-628 get_local_map_angular for every episode of the bound state. */
int ebc_ref_local_map_angular(ebc_sim *s, const ebc_angular_map *map, const double *poly_xy, const int32_t *poly_count,
                              double *out) {
  if (!s || !s->bound) return fail(s, EBC_ERR_UNBOUND, "ebc_local_map_angular: state not bound");
  if (!map || !poly_count || !out || map->dim < 1 || map->dim > 256 || map->max_polys < 0 ||
      (map->max_polys > 0 && !poly_xy))
    return fail(s, EBC_ERR_INVALID, "ebc_local_map_angular: bad argument");
  amap_t m;
  m.max_range = map->max_range; m.min_angle = map->min_angle; m.max_angle = map->max_angle; m.dim = map->dim;
  m.res = (map->max_angle - map->min_angle) / (double)map->dim;
  static const double SGN[4][2] = {{-1, -1}, {1, -1}, {-1, 1}, {1, 1}};               /* env.py:592-594 */
  for (int e = 0; e < s->cfg.n_episodes; ++e) {
    double *vec = out + (size_t)e * m.dim;
    for (int i = 0; i < m.dim; ++i) vec[i] = m.max_range;
    const double rx = s->st.rob_pv[(size_t)e * 4], ry = s->st.rob_pv[(size_t)e * 4 + 1];
    const double rad = s->st.rob_gr[(size_t)e * 4 + 3], theta = s->st.rob_theta[e];
    const double cs = cos(theta), sn = sin(theta);
    double edges[4][2];
    for (int k = 0; k < 4; ++k) { edges[k][0] = rx + SGN[k][0] * rad; edges[k][1] = ry + SGN[k][1] * rad; }
    int P = poly_count[e];
    if (P > map->max_polys) P = map->max_polys;
    int idx[4], n_seen;
    double loc[4][2];
    for (int o = 0; o < P; ++o) {                                                      /* env.py:596-609 */
      const double *poly = poly_xy + ((size_t)e * map->max_polys + o) * 8;
      for (int k = 0; k < 4; ++k) {
        n_seen = 0;
        for (int v = 0; v < 4; ++v) amap_calc(&m, poly + 2 * v, edges[k], cs, sn, vec, idx, loc, &n_seen);
      }
    }
    for (int o = 0; o < P; ++o) {                                                      /* env.py:611-620 */
      const double *poly = poly_xy + ((size_t)e * map->max_polys + o) * 8;
      for (int v = 0; v < 4; ++v) {
        n_seen = 0;
        for (int k = 0; k < 4; ++k) amap_calc(&m, poly + 2 * v, edges[k], cs, sn, vec, idx, loc, &n_seen);
      }
    }
    if (map->normalize)
      for (int i = 0; i < m.dim; ++i) vec[i] /= m.max_range;                           /* env.py:622-623 */
  }
  return EBC_OK;
}

/* ---- binary grid sub-map ([map] use_grid_map = true) ---------------------------------------------
 * simulator/env.py:630-692 get_local_map + :694-708 rotate_grid_around_center.  The rotation is OpenCV's
 * (opencv-python, third party, not under /root/reference, no pinned version): cv2.getRotationMatrix2D and
 * cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT, borderValue 1) on a float64 image, restated from OpenCV 4.x
 * modules/imgproc/src/imgwarp.cpp (getRotationMatrix2D, invertAffineTransform part of warpAffine, WarpAffineInvoker,
 * remapBilinear<Cast<double,double>, RemapNoVec, float>, initInterTab2D) and pinned to cv2 4.13 by
 * tests/golden/make_local_map_golden.py (bit-equal on every cell of the goldens, before thresholding). */
static int scene_map_is_free(const ebc_sim *s, int e, int ix, int iy) {          /* scene.map[ix][iy] (1 free, 0 wall) */
  const int16_t *rc = s->st.rect + (size_t)e * s->cfg.max_rects * 4;
  const int R = s->cfg.max_rects ? s->st.rect_count[e] : 0;
  for (int j = 0; j < R && j < s->cfg.max_rects; ++j)
    if (ix >= rc[4 * j] && ix < rc[4 * j + 2] && iy >= rc[4 * j + 1] && iy < rc[4 * j + 3]) return 0;
  return 1;
}

int ebc_ref_local_map_grid(ebc_sim *s, const ebc_grid_map *map, uint8_t *out) {
  if (!s || !s->bound) return fail(s, EBC_ERR_UNBOUND, "ebc_local_map_grid: state not bound");
  const ebc_config *c = &s->cfg;
  const int G = (int)rint(c->map_size_m / c->map_resolution);                      /* scene_generator.py:310-311 */
  if (!map || !out || map->size < 1 || map->size > 192 || map->size > G ||
      map->size != (int)rint(map->submap_size_m / c->map_resolution))
    return fail(s, EBC_ERR_INVALID, "ebc_local_map_grid: bad argument (size = round(submap_size_m / map_resolution), 1..192, <= map cells)");
  const int S = map->size;                                                          /* env.py:643 */
  float tab[32][2];                                                                 /* initInterTab1D, INTER_LINEAR */
  for (int i = 0; i < 32; ++i) {
    const float x = (float)i * (1.0f / 32);
    tab[i][0] = 1.0f - x;
    tab[i][1] = x;
  }
  for (int e = 0; e < c->n_episodes; ++e) {
    uint8_t *o = out + (size_t)e * S * S;
    const double px = s->st.rob_pv[(size_t)e * 4], py = s->st.rob_pv[(size_t)e * 4 + 1], theta = s->st.rob_theta[e];
    const int cx = (int)rint((px + c->map_size_m / 2.0) / c->map_resolution);       /* env.py:637-642 (round half even) */
    const int cy = (int)rint((py + c->map_size_m / 2.0) / c->map_resolution);
    int six = (int)rint((double)cx - floor((double)S / 2.0));                        /* env.py:645-648 */
    int siy = (int)rint((double)cy - floor((double)S / 2.0));
    int eix = six + S - 1, eiy = siy + S - 1;
    const int max_idx = G - 1;                                                      /* env.py:652-653 (square map) */
    int sgx = 0, sgy = 0, egx = S - 1, egy = S - 1;                                  /* env.py:655-658 */
    if (six < 0) { sgx = -six; six = 0; }                                            /* env.py:660-672 */
    else if (eix > max_idx) { egx = egx - (eix - max_idx); eix = max_idx; }
    if (siy < 0) { sgy = -siy; siy = 0; }
    else if (eiy > max_idx) { egy = egy - (eiy - max_idx); eiy = max_idx; }
    if (sgy > egy || siy > eiy || six > eix || sgx > egx) {                          /* env.py:674-680: all ones */
      memset(o, 1, (size_t)S * S);
      continue;
    }
    /* grid[sgx:egx, sgy:egy] = map[six:eix, siy:eiy] (env.py:682-684; the slices exclude their end index) */
#define GRID_AT(i, j)                                                                                         \
  (((i) < 0 || (i) >= S || (j) < 0 || (j) >= S) ? 1.0 /* borderValue */                                       \
   : ((i) >= sgx && (i) < egx && (j) >= sgy && (j) < egy) ? (double)scene_map_is_free(s, e, six + (i)-sgx, siy + (j)-sgy) \
                                                           : 1.0)
    const double angle = (-theta + M_PI / 2) * 180 / M_PI;                           /* env.py:686 */
    /* cv2.getRotationMatrix2D(center = (rows / 2, cols / 2), angle, 1) */
    const double a = angle * (M_PI / 180), alpha = cos(a), beta = sin(a), ctr = (double)S / 2.0;
    double M[6] = {alpha, beta, (1 - alpha) * ctr - beta * ctr, -beta, alpha, beta * ctr + (1 - alpha) * ctr};
    /* warpAffine without WARP_INVERSE_MAP: invert */
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    const double A11 = M[4] * D, A22 = M[0] * D;
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
    const double b1 = -M[0] * M[2] - M[1] * M[5], b2 = -M[3] * M[2] - M[4] * M[5];
    M[2] = b1; M[5] = b2;
    const int AB_SCALE = 1 << 10, round_delta = AB_SCALE / 32 / 2;
    for (int y = 0; y < S; ++y) {                                                   /* dst row = first index */
      const int X0 = (int)lrint((M[1] * y + M[2]) * AB_SCALE) + round_delta;
      const int Y0 = (int)lrint((M[4] * y + M[5]) * AB_SCALE) + round_delta;
      for (int x = 0; x < S; ++x) {
        const int X = (X0 + (int)lrint(M[0] * x * AB_SCALE)) >> 5, Y = (Y0 + (int)lrint(M[3] * x * AB_SCALE)) >> 5;
        const int sx = X >> 5, sy = Y >> 5, fx = X & 31, fy = Y & 31;
        const float w0 = tab[fy][0] * tab[fx][0], w1 = tab[fy][0] * tab[fx][1];
        const float w2 = tab[fy][1] * tab[fx][0], w3 = tab[fy][1] * tab[fx][1];
        const double v = GRID_AT(sy, sx) * w0 + GRID_AT(sy, sx + 1) * w1 + GRID_AT(sy + 1, sx) * w2 +
                         GRID_AT(sy + 1, sx + 1) * w3;
        o[(size_t)y * S + x] = v > 0.9 ? 1 : 0;                                      /* env.py:687-689 */
      }
    }
#undef GRID_AT
  }
  return EBC_OK;
}

int ebc_ref_reset(ebc_sim *s, const ebc_state *pool, int32_t pool_size, const int32_t *pool_index,
                  const uint8_t *mask) {
  if (!s || !s->bound) return fail(s, EBC_ERR_UNBOUND, "ebc_reset: state not bound");
  if (!pool || pool_size < 1) return fail(s, EBC_ERR_INVALID, "ebc_reset: bad pool");
  const ebc_config *c = &s->cfg;
  const int Hm = c->max_humans, Sm = c->max_statics, Rm = c->max_rects;
  for (int e = 0; e < c->n_episodes; ++e) {
    if (mask && !mask[e]) continue;
    int q = pool_index ? pool_index[e] : e;
    if (q < 0 || q >= pool_size) q = ((q % pool_size) + pool_size) % pool_size;
    memcpy(s->st.hum_pv + (size_t)e * Hm * 4, pool->hum_pv + (size_t)q * Hm * 4, sizeof(float) * 4 * Hm);
    memcpy(s->st.hum_gr + (size_t)e * Hm * 4, pool->hum_gr + (size_t)q * Hm * 4, sizeof(float) * 4 * Hm);
    memcpy(s->st.hum_type + (size_t)e * Hm, pool->hum_type + (size_t)q * Hm, Hm);
    memset(s->st.hum_nv + (size_t)e * Hm * 2, 0, sizeof(float) * 2 * Hm);
    s->st.hum_count[e] = pool->hum_count[q];
    if (Sm) memcpy(s->st.stat + (size_t)e * Sm * 4, pool->stat + (size_t)q * Sm * 4, sizeof(float) * 4 * Sm);
    s->st.stat_count[e] = pool->stat_count[q];
    if (Rm) memcpy(s->st.rect + (size_t)e * Rm * 4, pool->rect + (size_t)q * Rm * 4, sizeof(int16_t) * 4 * Rm);
    s->st.rect_count[e] = pool->rect_count[q];
    memcpy(s->st.rob_pv + (size_t)e * 4, pool->rob_pv + (size_t)q * 4, sizeof(float) * 4);
    memcpy(s->st.rob_gr + (size_t)e * 4, pool->rob_gr + (size_t)q * 4, sizeof(float) * 4);
    s->st.rob_theta[e] = pool->rob_theta[q];
    s->st.time[e] = pool->time[q];
    if (c->max_obst && s->st.obst && pool->obst && s->st.obst_count && pool->obst_count) {
      memcpy(s->st.obst + (size_t)e * c->max_obst, pool->obst + (size_t)q * c->max_obst,
             sizeof(ebc_obst_vertex) * (size_t)c->max_obst);
      s->st.obst_count[e] = pool->obst_count[q];
    }
  }
  return EBC_OK;
}
