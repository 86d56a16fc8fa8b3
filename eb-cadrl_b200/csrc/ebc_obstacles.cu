// ebc_obstacles.cu — host side of the ORCA obstacle half-planes: RVO2's addObstacle +
// processObstacles for one episode, producing the packed vertex records the kernels stage in shared memory.
//
// Replaces, for the caller simulator/policy/orca_obstacles.py:102-107,
//     for obstacle in global_map: sim.addObstacle(obstacle)
//     sim.processObstacles()
// i.e. RVOSimulator::addObstacle (one Obstacle per polygon vertex: point, unit direction to the next vertex,
// convexity, prev / next links) and KdTree::buildObstacleTree (a BSP over the obstacle edges whose splitting
// line is the edge that best balances the two sides; edges that straddle it are cut in two, which creates new
// vertices).  RVO2 walks that tree per agent, near side first, and keeps the visible edges sorted by distance
// with ties in visiting order.  The kernels do not walk a tree: each vertex record carries the set of its tree
// ancestors and on which side of each it hangs (two 64-bit masks), from which a warp derives both the range
// pruning and the visiting order of any two nodes with a few bit operations (ebc_sim.cu: obstacle_lines_warp).
// Pure host code, fp32 like RVO2; no CUDA calls.
#include <math.h>
#include <string.h>

#include <utility>
#include <vector>

#include "ebc_internal.cuh"

namespace {

constexpr float kEps = 0.00001f;   // RVO_EPSILON

struct Vtx {
  float x, y, ux, uy;
  int next, prev;
  bool convex;
  uint64_t anc = 0, anc_left = 0;
};

inline float det(float ax, float ay, float bx, float by) { return ax * by - ay * bx; }
// RVO2 leftOf(a, b, c) = det(a - c, b - a)
inline float left_of(const Vtx &a, const Vtx &b, const Vtx &c) { return det(a.x - c.x, a.y - c.y, b.x - a.x, b.y - a.y); }

class TreeBuilder {
 public:
  TreeBuilder(std::vector<Vtx> &v, size_t cap) : v_(v), cap_(cap) {}
  bool overflow() const { return overflow_; }

  // KdTree::buildObstacleTreeRecursive over `edges` (vertex ids; edge = vertex -> its next)
  void build(const std::vector<int> &edges, uint64_t anc, uint64_t anc_left) {
    if (edges.empty() || overflow_) return;
    const size_t n = edges.size();
    typedef std::pair<size_t, size_t> Score;          // (max(left, right), min(left, right)), smaller is better
    size_t best = 0;
    Score best_score(n, n);
    for (size_t i = 0; i < n; ++i) {
      size_t l = 0, r = 0;
      const Vtx &a = v_[edges[i]], &b = v_[a.next];
      for (size_t j = 0; j < n; ++j) {
        if (i == j) continue;
        const Vtx &p = v_[edges[j]], &q = v_[p.next];
        const float sp = left_of(a, b, p), sq = left_of(a, b, q);
        if (sp >= -kEps && sq >= -kEps) ++l;
        else if (sp <= kEps && sq <= kEps) ++r;
        else { ++l; ++r; }
        if (Score(l > r ? l : r, l > r ? r : l) >= best_score) break;
      }
      const Score sc(l > r ? l : r, l > r ? r : l);
      if (sc < best_score) { best_score = sc; best = i; }
    }
    const int node = edges[best];
    std::vector<int> lefts, rights;
    for (size_t j = 0; j < n; ++j) {
      if (j == best) continue;
      const int pj = edges[j];
      const Vtx a = v_[node], b = v_[a.next];
      const Vtx p = v_[pj], q = v_[p.next];
      const float sp = left_of(a, b, p), sq = left_of(a, b, q);
      if (sp >= -kEps && sq >= -kEps) lefts.push_back(pj);
      else if (sp <= kEps && sq <= kEps) rights.push_back(pj);
      else {
        // cut edge p -> q where it meets the line through a -> b
        if (v_.size() >= cap_) { overflow_ = true; return; }
        const float t = det(b.x - a.x, b.y - a.y, p.x - a.x, p.y - a.y) / det(b.x - a.x, b.y - a.y, p.x - q.x, p.y - q.y);
        Vtx cut;
        cut.x = p.x + t * (q.x - p.x);
        cut.y = p.y + t * (q.y - p.y);
        cut.ux = p.ux; cut.uy = p.uy;
        cut.prev = pj; cut.next = p.next;
        cut.convex = true;
        const int id = (int)v_.size();
        v_[p.next].prev = id;
        v_[pj].next = id;
        v_.push_back(cut);
        if (sp > 0.0f) { lefts.push_back(pj); rights.push_back(id); }
        else { rights.push_back(pj); lefts.push_back(id); }
      }
    }
    v_[node].anc = anc;
    v_[node].anc_left = anc_left;
    const uint64_t bit = (uint64_t)1 << node;
    build(lefts, anc | bit, anc_left | bit);
    build(rights, anc | bit, anc_left);
  }

 private:
  std::vector<Vtx> &v_;
  size_t cap_;
  bool overflow_ = false;
};

}  // namespace

extern "C" int ebc_pack_obstacles(const float *xy, const int32_t *poly_size, int32_t n_poly, ebc_obst_vertex *out,
                                  int32_t cap, int32_t *n_out) {
  if (!poly_size || !out || !n_out || n_poly < 0 || cap < 0 || (n_poly > 0 && !xy)) return EBC_ERR_INVALID;
  if (cap > 64) cap = 64;
  std::vector<Vtx> v;
  v.reserve(64);
  size_t off = 0;
  for (int p = 0; p < n_poly; ++p) {                  // RVOSimulator::addObstacle
    const int m = poly_size[p];
    if (m < 2 || v.size() + (size_t)m > (size_t)cap) return EBC_ERR_INVALID;
    const float *q = xy + 2 * off;
    const int base = (int)v.size();
    for (int i = 0; i < m; ++i) {
      const int nx = (i + 1) % m, pr = (i + m - 1) % m;
      Vtx o;
      o.x = q[2 * i]; o.y = q[2 * i + 1];
      o.next = base + nx; o.prev = base + pr;
      const float dx = q[2 * nx] - o.x, dy = q[2 * nx + 1] - o.y;
      const float inv = 1.0f / sqrtf(dx * dx + dy * dy);
      o.ux = dx * inv; o.uy = dy * inv;
      if (m == 2) o.convex = true;
      else o.convex = det(q[2 * pr] - q[2 * nx], q[2 * pr + 1] - q[2 * nx + 1], o.x - q[2 * pr], o.y - q[2 * pr + 1]) >= 0.0f;
      v.push_back(o);
    }
    off += (size_t)m;
  }
  std::vector<int> all(v.size());
  for (size_t i = 0; i < v.size(); ++i) all[i] = (int)i;
  TreeBuilder tb(v, (size_t)cap);
  tb.build(all, 0, 0);                                // RVOSimulator::processObstacles
  if (tb.overflow()) return EBC_ERR_INVALID;
  for (size_t i = 0; i < v.size(); ++i) {
    ebc_obst_vertex r;
    memset(&r, 0, sizeof(r));
    r.px = v[i].x; r.py = v[i].y; r.ux = v[i].ux; r.uy = v[i].uy;
    r.next = (int16_t)v[i].next; r.prev = (int16_t)v[i].prev; r.convex = v[i].convex ? 1 : 0;
    r.anc = v[i].anc; r.anc_left = v[i].anc_left;
    out[i] = r;
  }
  *n_out = (int32_t)v.size();
  return EBC_OK;
}
