// lbvh_traverse.cuh -- closest-hit / any-hit traversal of the packed LBVH (replaces Embree's
// BVHNIntersector1 reached through rtcIntersect1 / rtcOccluded1, cpp/embree_connect.cpp:218,
// cpp/embree3/rtcore_scene.h:128, and scales geometry.adb:266-323 IntersectMeshBF beyond a handful of triangles).
//
// Node (32 B = 2 x float4 = one DRAM sector, two 16-byte vector loads): both child boxes quantised to 8 bits per bound
// inside the node's own box, conservatively (a decoded box contains the exact one):
//   q0 = (origin.x, origin.y, origin.z, bits: ex | ey << 8 | ez << 16 | leaf flags << 24)   step_k = 2^(e_k - 127)
//   q1 = bits (c0.lo.xyz, c0.hi.x), (c0.hi.yz, c1.lo.xy), (c1.lo.z, c1.hi.xyz), gamma      bound = origin + byte * step
//   children = (gamma, gamma + 1): node index, or -- leaf flag set -- sorted triangle slot (one triangle per leaf)
// Triangle slot (48 B = 3 x float4): (A.xyz, bits prim id), (e1.xyz, 0), (e2.xyz, 0) with e1 = B - A, e2 = C - A.
//
// Semantics: nearest hit of the reference's one-sided Moeller-Trumbore test (geometry.adb:231-263) inside the
// open window (t_min, t_max); exact ties go to the lowest ORIGINAL triangle index.  Box tests only cull: boxes
// are padded at build time and the slab test is evaluated conservatively, so the result equals testing every
// triangle.  TWO_SIDED (legacy seam only, selectable): both faces are hit, like Embree's rtcIntersect1.
// ANY_HIT (gcore_bvh_occluded_batch): the walk ends at the first accepted hit, like rtcOccluded1.
#pragma once

namespace art {

// One stack entry per level at most (the far child is pushed while descending).  A Karras tree over 63-bit Morton codes with the
// sorted position as tie-break is at most 63 + 32 levels deep, so 96 entries cannot overflow; the guards below are belt and braces.
#define LBVH_STACK 96

ART_D bool slab_hit(float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz,
                    f3 o, f3 inv, float t_lo, float t_hi, float& t_entry) {
  float x0 = (bminx - o.x) * inv.x, x1 = (bmaxx - o.x) * inv.x;
  float y0 = (bminy - o.y) * inv.y, y1 = (bmaxy - o.y) * inv.y;
  float z0 = (bminz - o.z) * inv.z, z1 = (bmaxz - o.z) * inv.z;
  // fminf/fmaxf drop NaNs (0 * inf on axis-parallel rays lying in a slab plane) -> never culls wrongly
  float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), t_lo));
  float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), t_hi));
  t_entry = tn;
  return tn <= tf * 1.00001f + 1.0e-6f;
}

struct BvhStats { unsigned int nodes, tris; };     // nodes visited (32 B each), triangle slots tested (48 B each)

struct BvhNode2 { float b0[6], b1[6]; int c0, c1; };          // decoded: child boxes (min.xyz, max.xyz), child links (>= 0 node, < 0: ~slot)

ART_D BvhNode2 bvh_fetch_node(const BvhView& bv, int node) {
#if LBVH_NODE_BYTES == 64
  // experiment layout: both child boxes in full precision (n0 = c0.min.xyz c0.max.x, n1 = c0.max.yz c1.min.xy, n2 = c1.min.z c1.max.xyz,
  // n3 = gamma | flags) -- no decode arithmetic, twice the bytes
  const float4 n0 = __ldg(&bv.nodes[4 * node + 0]);
  const float4 n1 = __ldg(&bv.nodes[4 * node + 1]);
  const float4 n2 = __ldg(&bv.nodes[4 * node + 2]);
  const float4 n3 = __ldg(&bv.nodes[4 * node + 3]);
  BvhNode2 n;
  n.b0[0] = n0.x; n.b0[1] = n0.y; n.b0[2] = n0.z; n.b0[3] = n0.w; n.b0[4] = n1.x; n.b0[5] = n1.y;
  n.b1[0] = n1.z; n.b1[1] = n1.w; n.b1[2] = n2.x; n.b1[3] = n2.y; n.b1[4] = n2.z; n.b1[5] = n2.w;
  const int gamma = __float_as_int(n3.x); const unsigned int w0 = __float_as_uint(n3.y);
  n.c0 = (w0 & 1u) ? ~gamma : gamma;
  n.c1 = (w0 & 2u) ? ~(gamma + 1) : (gamma + 1);
  return n;
#else
  const float4 q0 = __ldg(&bv.nodes[2 * node + 0]);
  const float4 q1 = __ldg(&bv.nodes[2 * node + 1]);
  const unsigned int w0 = __float_as_uint(q0.w), a = __float_as_uint(q1.x), b = __float_as_uint(q1.y), c = __float_as_uint(q1.z);
  const float sx = __uint_as_float((w0 & 255u) << 23), sy = __uint_as_float(((w0 >> 8) & 255u) << 23), sz = __uint_as_float(((w0 >> 16) & 255u) << 23);
  BvhNode2 n;
  n.b0[0] = q0.x + (float)(a & 255u) * sx;         n.b0[1] = q0.y + (float)((a >> 8) & 255u) * sy;  n.b0[2] = q0.z + (float)((a >> 16) & 255u) * sz;
  n.b0[3] = q0.x + (float)(a >> 24) * sx;          n.b0[4] = q0.y + (float)(b & 255u) * sy;         n.b0[5] = q0.z + (float)((b >> 8) & 255u) * sz;
  n.b1[0] = q0.x + (float)((b >> 16) & 255u) * sx; n.b1[1] = q0.y + (float)(b >> 24) * sy;          n.b1[2] = q0.z + (float)(c & 255u) * sz;
  n.b1[3] = q0.x + (float)((c >> 8) & 255u) * sx;  n.b1[4] = q0.y + (float)((c >> 16) & 255u) * sy; n.b1[5] = q0.z + (float)(c >> 24) * sz;
  const int gamma = __float_as_int(q1.w);
  n.c0 = (w0 & (1u << 24)) ? ~gamma : gamma;
  n.c1 = (w0 & (2u << 24)) ? ~(gamma + 1) : (gamma + 1);
  return n;
#endif
}

// Moeller-Trumbore without the max(det, 1e-25) back-face rejection: det of either sign (Embree semantics)
ART_D bool intersect_triangle_two_sided(f3 o, f3 d, f3 A, f3 edge1, f3 edge2, float t_min, float t_max, float& t_o, float& u_o, float& v_o) {
  f3 pvec = cross(d, edge2);
  f3 tvec = o - A;
  float det = dot(edge1, pvec);
  if (det == 0.0f) return false;
  float invDet = 1.0f / det;
  float v = dot(tvec, pvec) * invDet;
  if (!(v > 0.0f)) return false;
  f3 qvec = cross(tvec, edge1);
  float u = dot(qvec, d) * invDet;
  if (!(u > 0.0f && u + v < 1.0f)) return false;
  float t = dot(edge2, qvec) * invDet;
  if (t > t_min && t < t_max) { t_o = t; u_o = u; v_o = v; return true; }
  return false;
}

// Both child boxes of `node` against one ray: h0/h1 = the ray's window (t_lo, t_hi) overlaps the child's box, e0/e1 = entry parameters.
// The 32-byte layout is tested without materialising the boxes: plane parameter t = (origin + q * 2^e - o) * inv is evaluated as
// fma(q, 2^e * inv, (origin - o) * inv) -- one fused operation per plane instead of multiply, add, subtract, multiply.  Culling only has
// to be conservative (the triangle tests decide the hit, bit for bit): the rounding differs from slab_hit's by a few ulp of the plane
// parameter, far inside the 1e-5 relative / 1e-6 absolute slack below; 0 * inf on axis-parallel rays gives NaN, which fminf/fmaxf drop.
#ifndef LBVH_FUSED_SLABS
#define LBVH_FUSED_SLABS 1
#endif
// 1/d for the slab tests.  An axis-parallel ray has 1/d = +-inf; in the fused form inf * q + (-inf) is NaN on every plane of an axis the
// ray runs beside, the axis drops out of the test and a walk that should touch a line of boxes touches a whole slab of the tree (measured:
// one such ray in 5e7 -- a cosine sample with r2 = 0 off an axis-aligned wall -- held a 16 ms launch for 1.5 s).  Clamped to +-1e18 the
// plane parameters keep their sign and stay finite for any scene below 1e20 units; they only differ from the unclamped ones where
// |plane - o| < 1e-12, below the resolution of the coordinates.
ART_D f3 bvh_slab_inv_of(f3 inv) {
  const float big = 1.0e18f;
  return mk3(fminf(fmaxf(inv.x, -big), big), fminf(fmaxf(inv.y, -big), big), fminf(fmaxf(inv.z, -big), big));
}
ART_D void bvh_node_children(const BvhView& bv, int node, f3 o, f3 inv, float t_lo, float t_hi,
                             int& c0, int& c1, bool& h0, bool& h1, float& e0, float& e1) {
#if LBVH_NODE_BYTES == 64 || !LBVH_FUSED_SLABS
  const BvhNode2 nd = bvh_fetch_node(bv, node);
  c0 = nd.c0; c1 = nd.c1;
  h0 = slab_hit(nd.b0[0], nd.b0[1], nd.b0[2], nd.b0[3], nd.b0[4], nd.b0[5], o, inv, t_lo, t_hi, e0);
  h1 = slab_hit(nd.b1[0], nd.b1[1], nd.b1[2], nd.b1[3], nd.b1[4], nd.b1[5], o, inv, t_lo, t_hi, e1);
#else
  const float4 q0 = __ldg(&bv.nodes[2 * node + 0]);
  const float4 q1 = __ldg(&bv.nodes[2 * node + 1]);
  const unsigned int w0 = __float_as_uint(q0.w), a = __float_as_uint(q1.x), b = __float_as_uint(q1.y), c = __float_as_uint(q1.z);
  const float kx = __uint_as_float((w0 & 255u) << 23) * inv.x, ky = __uint_as_float(((w0 >> 8) & 255u) << 23) * inv.y,
              kz = __uint_as_float(((w0 >> 16) & 255u) << 23) * inv.z;
  const float bx = (q0.x - o.x) * inv.x, by = (q0.y - o.y) * inv.y, bz = (q0.z - o.z) * inv.z;
  {
    const float x0 = __fmaf_rn((float)(a & 255u), kx, bx),         x1 = __fmaf_rn((float)(a >> 24), kx, bx);
    const float y0 = __fmaf_rn((float)((a >> 8) & 255u), ky, by),  y1 = __fmaf_rn((float)(b & 255u), ky, by);
    const float z0 = __fmaf_rn((float)((a >> 16) & 255u), kz, bz), z1 = __fmaf_rn((float)((b >> 8) & 255u), kz, bz);
    const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), t_lo));
    const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), t_hi));
    e0 = tn; h0 = tn <= __fmaf_rn(tf, 1.00001f, 1.0e-6f);
  }
  {
    const float x0 = __fmaf_rn((float)((b >> 16) & 255u), kx, bx), x1 = __fmaf_rn((float)((c >> 8) & 255u), kx, bx);
    const float y0 = __fmaf_rn((float)(b >> 24), ky, by),          y1 = __fmaf_rn((float)((c >> 16) & 255u), ky, by);
    const float z0 = __fmaf_rn((float)(c & 255u), kz, bz),         z1 = __fmaf_rn((float)(c >> 24), kz, bz);
    const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), t_lo));
    const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), t_hi));
    e1 = tn; h1 = tn <= __fmaf_rn(tf, 1.00001f, 1.0e-6f);
  }
  const int gamma = __float_as_int(q1.w);
  c0 = (w0 & (1u << 24)) ? ~gamma : gamma;
  c1 = (w0 & (2u << 24)) ? ~(gamma + 1) : (gamma + 1);
#endif
}

template <bool TWO_SIDED>
ART_D void bvh_leaf_test(const BvhView& bv, int first, int count, f3 o, f3 d, float t_min,
                         float& best_t, int& best_prim, float& best_u, float& best_v, int& best_slot) {
#pragma unroll 1
  for (int s = first; s < first + count; s++) {
    float4 a  = __ldg(&bv.tris[3 * s + 0]);
    float4 e1 = __ldg(&bv.tris[3 * s + 1]);
    float4 e2 = __ldg(&bv.tris[3 * s + 2]);
    int prim = __float_as_int(a.w);
    float t, u, v;
    // closed upper bound so that equal-t candidates reach the tie rule
    const bool hit = TWO_SIDED ? intersect_triangle_two_sided(o, d, mk3(a.x, a.y, a.z), mk3(e1.x, e1.y, e1.z), mk3(e2.x, e2.y, e2.z), t_min, ART_FLT_LAST, t, u, v)
                               : intersect_triangle(o, d, mk3(a.x, a.y, a.z), mk3(e1.x, e1.y, e1.z), mk3(e2.x, e2.y, e2.z), t_min, ART_FLT_LAST, t, u, v);
    if (hit) {
      if (t < best_t || (t == best_t && best_prim >= 0 && prim < best_prim)) { best_t = t; best_prim = prim; best_u = u; best_v = v; best_slot = s; }
    }
  }
}

// closest hit; prim = -1 when nothing is hit inside (t_min, t_max); slot = position of the winner in the Morton-ordered
// triangle array (for callers that want its edges)
template <bool STATS, bool TWO_SIDED = false, bool ANY_HIT = false>
ART_D void bvh_closest_slot_t(const BvhView& bv, f3 o, f3 d, f3 inv, float t_min, float t_max, int& prim_o, float& t_o, float& u_o, float& v_o, int& slot_o,
                              BvhStats& st) {
  float best_t = t_max; int best_prim = -1; float best_u = 0.0f, best_v = 0.0f; int best_slot = -1;
  inv = bvh_slab_inv_of(inv);
  if (bv.ntris > 0) {
    if (bv.root < 0) {
      bvh_leaf_test<TWO_SIDED>(bv, 0, bv.ntris, o, d, t_min, best_t, best_prim, best_u, best_v, best_slot);
    } else {
      int stack[LBVH_STACK];
      int sp = 0;
      int node = bv.root;
      for (;;) {
        if (STATS) st.nodes++;
        int c0, c1; float e0, e1; bool h0, h1;
        bvh_node_children(bv, node, o, inv, t_min, best_t, c0, c1, h0, h1, e0, e1);
        if (STATS) { if (h0 && c0 < 0) st.tris += 1; if (h1 && c1 < 0) st.tris += 1; }
        if (h0 && c0 < 0) { bvh_leaf_test<TWO_SIDED>(bv, ~c0, 1, o, d, t_min, best_t, best_prim, best_u, best_v, best_slot); h0 = false; }
        if (h1 && c1 < 0) { bvh_leaf_test<TWO_SIDED>(bv, ~c1, 1, o, d, t_min, best_t, best_prim, best_u, best_v, best_slot); h1 = false; }
        if (ANY_HIT && best_prim >= 0) break;                           // occlusion query: any hit inside the window will do
        if (h0 && h1) {
          // near child first, far child on the stack
          if (e1 < e0) { int t = c0; c0 = c1; c1 = t; }
          if (sp < LBVH_STACK) stack[sp++] = c1;
          node = c0;
        } else if (h0) node = c0;
        else if (h1) node = c1;
        else { if (sp == 0) break; node = stack[--sp]; }
      }
    }
  }
  prim_o = best_prim;
  t_o = best_t; u_o = best_u; v_o = best_v; slot_o = best_slot;
}

ART_D void bvh_closest_slot(const BvhView& bv, f3 o, f3 d, f3 inv, float t_min, float t_max, int& prim_o, float& t_o, float& u_o, float& v_o, int& slot_o) {
  BvhStats st;
  bvh_closest_slot_t<false>(bv, o, d, inv, t_min, t_max, prim_o, t_o, u_o, v_o, slot_o, st);
}

ART_D void bvh_closest(const BvhView& bv, f3 o, f3 d, f3 inv, float t_min, float t_max, int& prim_o, float& t_o, float& u_o, float& v_o) {
  int slot;
  bvh_closest_slot(bv, o, d, inv, t_min, t_max, prim_o, t_o, u_o, v_o, slot);
}

// ---- resumable closest-hit traversal for the render kernel: the same walk in slices of at most `max_steps` nodes, so
// that no stage waits for the longest ray of the block.  The traversal state is small and lives OUTSIDE the thread: the
// cursor (next node, stack depth) and the running best hit are per path slot in shared memory, the first LBVH_SHORT_STACK
// stack entries in shared memory too (a short stack covers almost every ray: the far children pending on a near-first
// walk), deeper entries in a per-slot global overflow area.  Any thread can therefore continue any ray -- each slice is
// run by the threads of a dense queue of pending rays, not by the slots' owners (render_wave.cuh).
#define LBVH_SHORT_STACK 8
#define LBVH_OVERFLOW (LBVH_STACK - LBVH_SHORT_STACK)

struct BvhRayState {
  int   node;                       // next node to visit; -1: finished; -2: single-leaf tree, not started
  int   sp;
  float best_t; int best_prim; float best_u, best_v;
};

ART_D void bvh_state_begin(const BvhView& bv, float t_max, BvhRayState& c) {
  c.node = (bv.ntris > 0 && bv.root >= 0) ? bv.root : ((bv.ntris > 0) ? -2 : -1);
  c.sp = 0; c.best_t = t_max; c.best_prim = -1; c.best_u = 0.0f; c.best_v = 0.0f;
}

// smem_stack: this ray's column of the shared short stack (entry k at smem_stack[k * stride]); ovf: this ray's overflow area
// returns true when the traversal is complete (c.best_* hold the answer; best_prim < 0: nothing hit)
ART_D bool bvh_state_advance(const BvhView& bv, f3 o, f3 d, f3 inv, float t_min, int max_steps, BvhRayState& c,
                             int* smem_stack, int stride, int* ovf, BvhStats& st) {
  int dummy_slot;
  if (c.node == -2) { bvh_leaf_test<false>(bv, 0, bv.ntris, o, d, t_min, c.best_t, c.best_prim, c.best_u, c.best_v, dummy_slot); c.node = -1; }
  if (c.node < 0) return true;
  int node = c.node, sp = c.sp;
  float best_t = c.best_t, best_u = c.best_u, best_v = c.best_v; int best_prim = c.best_prim;
  bool done = false;
#pragma unroll 1
  for (int step = 0; step < max_steps; step++) {
    st.nodes++;
    int c0, c1; float e0, e1; bool h0, h1;
    bvh_node_children(bv, node, o, inv, t_min, best_t, c0, c1, h0, h1, e0, e1);
    if (h0 && c0 < 0) { st.tris += 1; bvh_leaf_test<false>(bv, ~c0, 1, o, d, t_min, best_t, best_prim, best_u, best_v, dummy_slot); h0 = false; }
    if (h1 && c1 < 0) { st.tris += 1; bvh_leaf_test<false>(bv, ~c1, 1, o, d, t_min, best_t, best_prim, best_u, best_v, dummy_slot); h1 = false; }
    if (h0 && h1) {
      if (e1 < e0) { int t = c0; c0 = c1; c1 = t; }                    // near child first, far child on the stack
      if (sp < LBVH_SHORT_STACK) { smem_stack[sp * stride] = c1; sp++; }
      else if (sp < LBVH_STACK) { ovf[sp - LBVH_SHORT_STACK] = c1; sp++; }
      node = c0;
    } else if (h0) node = c0;
    else if (h1) node = c1;
    else {
      if (sp == 0) { done = true; break; }
      sp--;
      node = (sp < LBVH_SHORT_STACK) ? smem_stack[sp * stride] : ovf[sp - LBVH_SHORT_STACK];
    }
  }
  c.node = done ? -1 : node; c.sp = sp;
  c.best_t = best_t; c.best_prim = best_prim; c.best_u = best_u; c.best_v = best_v;
  return done;
}

// ---- the same walk with the leaves postponed (render_wave.cuh's traversal warps): a warp that tests a triangle the moment ONE lane
// reaches a leaf runs the ~90-instruction test (and waits for its 48 bytes) at 1-3 lanes almost every step.  Here a step handles inner
// nodes only; children that are leaves are parked in pend0/pend1 (links < 0; 0 = none) and the warp tests parked leaves together, when
// enough lanes hold one.  The set of nodes visited, their order and the candidate set are those of bvh_state_advance: the children
// hit at a node are decided before its leaves are tested there as well, and best_t is current again before the next node is fetched.
// node: next inner node, -1 = walk complete.
ART_D void bvh_walk_step(const BvhView& bv, f3 o, f3 inv, float t_min, float best_t, int& node, int& sp, int& pend0, int& pend1,
                         int* smem_stack, int stride, int* ovf) {
  int c0, c1; float e0, e1; bool h0, h1;
  bvh_node_children(bv, node, o, inv, t_min, best_t, c0, c1, h0, h1, e0, e1);
  if (h0 && c0 < 0) { pend0 = c0; h0 = false; }
  if (h1 && c1 < 0) { if (pend0 == 0) pend0 = c1; else pend1 = c1; h1 = false; }
  if (h0 && h1) {
    if (e1 < e0) { int t = c0; c0 = c1; c1 = t; }                      // near child first, far child on the stack
    if (sp < LBVH_SHORT_STACK) { smem_stack[sp * stride] = c1; sp++; }
    else if (sp < LBVH_STACK) { ovf[sp - LBVH_SHORT_STACK] = c1; sp++; }
    node = c0;
  } else if (h0) node = c0;
  else if (h1) node = c1;
  else if (sp == 0) node = -1;
  else {
    sp--;
    node = (sp < LBVH_SHORT_STACK) ? smem_stack[sp * stride] : ovf[sp - LBVH_SHORT_STACK];
  }
}

}  // namespace art
