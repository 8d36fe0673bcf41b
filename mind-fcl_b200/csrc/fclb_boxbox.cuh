// fclb_boxbox.cuh -- Box-Box SAT + face clipping, up to 4 contacts.
//
// Behavioural contract: include/fcl/narrowphase/detail/primitive_shape_algorithm/
// box_box-inl.h: boxBox2 (:214-807), lineClosestApproach (:51-72),
// intersectRectQuad2 (:75-135), cullPoints2 (:138-211), boxBoxIntersect (:824-846).
// This is what fcl::collide(Box, Box) runs (gjk_solver-inl.h:210); it is NOT
// GJK/EPA.  Per-thread; the small clip buffers index dynamically
// and live in local memory (L1-resident).
//
// Known platform dependence: cullPoints2 calls an unqualified atan2(), i.e. the
// C double routine even for S = float; CUDA's atan2(double) is not correctly
// rounded (<= 2 ulp), so an angle comparison on a knife edge can select a
// different subset of the same contact polygon.  Only reached when a face
// contact has more than 4 clip points.
#pragma once
#include "fclb_primitives_intersect.cuh"

namespace fclb {

template <typename S>
struct num_limits;
template <>
struct num_limits<float> {
  static FCLB_DI float eps() { return 1.1920928955078125e-07f; }
  static FCLB_DI float max() { return 3.402823466e+38f; }
  static FCLB_DI float pi() { return 3.14159265358979323846f; }
};
template <>
struct num_limits<double> {
  static FCLB_DI double eps() { return 2.220446049250313e-16; }
  static FCLB_DI double max() { return 1.7976931348623157e+308; }
  static FCLB_DI double pi() { return 3.14159265358979323846; }
};

// box_box-inl.h:75-135.  h[2] half sizes, p[8] quad corners; returns the number
// of points written to ret[16] (at most 8).
template <typename S>
FCLB_DI int intersectRectQuad2(const S h[2], const S p[8], S ret[16]) {
  // q = current polygon (nq points), r = chopped polygon (nr points).  The
  // reference starts with q = p, r = ret and then ping-pongs r between a
  // scratch buffer and ret.
  int nq = 4, nr = 0;
  S buffer[16];
  const S* q = p;
  S* r = ret;
  for (int dir = 0; dir <= 1; ++dir) {
    for (int sign = -1; sign <= 1; sign += 2) {
      const S sg = S(sign);
      const S* pq = q;
      S* pr = r;
      nr = 0;
      #pragma unroll 1
      for (int i = nq; i > 0; --i) {
        if (sg * pq[dir] < h[dir]) {
          pr[0] = pq[0];
          pr[1] = pq[1];
          pr += 2;
          nr++;
          if (nr & 8) {
            q = r;
            goto done;
          }
        }
        const S* nextq = (i > 1) ? pq + 2 : q;
        if ((sg * pq[dir] < h[dir]) ^ (sg * nextq[dir] < h[dir])) {
          pr[1 - dir] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (sg * h[dir] - pq[dir]);
          pr[dir] = sg * h[dir];
          pr += 2;
          nr++;
          if (nr & 8) {
            q = r;
            goto done;
          }
        }
        pq += 2;
      }
      q = r;
      r = (q == ret) ? buffer : ret;
      nq = nr;
    }
  }
done:
  if (q != ret)
    #pragma unroll 1
    for (int i = 0; i < nr * 2; i++) ret[i] = q[i];
  return nr;
}

// box_box-inl.h:138-211
template <typename S>
FCLB_DI void cullPoints2(int n, const S p[], int m, int i0, int iret[]) {
  S a, cx, cy, q;
  if (n == 1) {
    cx = p[0];
    cy = p[1];
  } else if (n == 2) {
    cx = S(0.5) * (p[0] + p[2]);
    cy = S(0.5) * (p[1] + p[3]);
  } else {
    a = 0;
    cx = 0;
    cy = 0;
    #pragma unroll 1
    for (int i = 0; i < n - 1; ++i) {
      q = p[i * 2] * p[i * 2 + 3] - p[i * 2 + 2] * p[i * 2 + 1];
      a += q;
      cx += q * (p[i * 2] + p[i * 2 + 2]);
      cy += q * (p[i * 2 + 1] + p[i * 2 + 3]);
    }
    q = p[n * 2 - 2] * p[1] - p[0] * p[n * 2 - 1];
    if (fabs_(a + q) > num_limits<S>::eps())
      a = S(1) / (S(3) * (a + q));
    else
      a = S(1e18f);
    cx = a * (cx + q * (p[n * 2 - 2] + p[0]));
    cy = a * (cy + q * (p[n * 2 - 1] + p[1]));
  }
  S A[8];
  #pragma unroll 1
  for (int i = 0; i < n; ++i) A[i] = S(atan2(double(p[i * 2 + 1] - cy), double(p[i * 2] - cx)));
  int avail[8];
  #pragma unroll 1
  for (int i = 0; i < n; ++i) avail[i] = 1;
  avail[i0] = 0;
  iret[0] = i0;
  int k = 1;
  const S pi = num_limits<S>::pi();
  #pragma unroll 1
  for (int j = 1; j < m; ++j) {
    a = S(j) * (S(2) * pi / S(m)) + A[i0];
    if (a > pi) a -= S(2) * pi;
    S maxdiff = S(1e9), diff;
    iret[k] = i0;
    #pragma unroll 1
    for (int i = 0; i < n; ++i) {
      if (avail[i]) {
        diff = fabs_(A[i] - a);
        if (diff > pi) diff = S(2) * pi - diff;
        if (diff < maxdiff) {
          maxdiff = diff;
          iret[k] = i;
        }
      }
    }
    avail[iret[k]] = 0;
    k++;
  }
}

// box_box-inl.h:214-807 with maxc = 4 (boxBoxIntersect, :824-846).
// Returns return_code (0 = separated); *n_contacts contacts are written.
// SAT_ONLY: stop after the 15-axis separating test (the return code is already final; the contact
// generation that follows is what the two-phase box-box kernel runs on the compacted colliding queries).
template <typename S, bool SAT_ONLY = false>
FCLB_DI int boxBox2(const V3<S>& side1, const Pose<S>& tf1, const V3<S>& side2, const Pose<S>& tf2, ContactPt<S> out[4],
                    int* n_contacts) {
  const S fudge_factor = S(1.05);
  const M3<S>& R1 = tf1.R;
  const M3<S>& R2 = tf2.R;
  const V3<S> T1 = tf1.t, T2 = tf2.t;
  *n_contacts = 0;
  int maxc = 4;

  const V3<S> p = T2 - T1;
  const V3<S> pp = mulMtV(R1, p);
  const V3<S> A = side1 * S(0.5);
  const V3<S> B = side2 * S(0.5);
  const M3<S> R = mulMtM(R1, R2);
  M3<S> Q;
#pragma unroll
  for (int i = 0; i < 9; i++) Q.m[i] = fabs_(R.m[i]);

  int best_col_id = -1;
  int normalR = 0;  // 0 none, 1 = R1, 2 = R2
  S tmp = 0, s2, l;
  S s = -num_limits<S>::max();
  int invert_normal = 0, code = 0;
  V3<S> normalC = zero3<S>();

#define FCLB_BB_FACE(TMP, RAD, COL, WHICH, CODE) \
  tmp = (TMP);                                   \
  s2 = fabs_(tmp) - (RAD);                       \
  if (s2 > 0) return 0;                          \
  if (s2 > s) {                                  \
    s = s2;                                      \
    best_col_id = (COL);                         \
    normalR = (WHICH);                           \
    invert_normal = (tmp < 0);                   \
    code = (CODE);                               \
  }
  FCLB_BB_FACE(pp.x, dot(row(Q, 0), B) + A.x, 0, 1, 1)
  FCLB_BB_FACE(pp.y, dot(row(Q, 1), B) + A.y, 1, 1, 2)
  FCLB_BB_FACE(pp.z, dot(row(Q, 2), B) + A.z, 2, 1, 3)
  FCLB_BB_FACE(dot(col(R2, 0), p), dot(col(Q, 0), A) + B.x, 0, 2, 4)
  FCLB_BB_FACE(dot(col(R2, 1), p), dot(col(Q, 1), A) + B.y, 1, 2, 5)
  FCLB_BB_FACE(dot(col(R2, 2), p), dot(col(Q, 2), A) + B.z, 2, 2, 6)
#undef FCLB_BB_FACE

  const S eps = num_limits<S>::eps();
  {
    const S amax = fmax_(fmax_(A.x, A.y), A.z);  // maxCoeff: value only
    const S bmax = fmax_(fmax_(B.x, B.y), B.z);
    const S scale_factor = fmax_(fmax_(amax, bmax), S(1.0)) * S(10) * eps;
#pragma unroll
    for (int i = 0; i < 9; i++) Q.m[i] += scale_factor;
  }

  V3<S> n;
#define FCLB_BB_EDGE(TMP, RAD, NX, NY, NZ, CODE) \
  tmp = (TMP);                                   \
  s2 = fabs_(tmp) - (RAD);                       \
  if (s2 > 0) return 0;                          \
  n = mk<S>((NX), (NY), (NZ));                   \
  l = norm(n);                                   \
  if (l > eps) {                                 \
    s2 /= l;                                     \
    if (s2 * fudge_factor > s) {                 \
      s = s2;                                    \
      best_col_id = -1;                          \
      normalC = n / l;                           \
      invert_normal = (tmp < 0);                 \
      code = (CODE);                             \
    }                                            \
  }
  // u1 x (v1,v2,v3)
  FCLB_BB_EDGE(pp.z * R(1, 0) - pp.y * R(2, 0), A.y * Q(2, 0) + A.z * Q(1, 0) + B.y * Q(0, 2) + B.z * Q(0, 1), S(0),
               -R(2, 0), R(1, 0), 7)
  FCLB_BB_EDGE(pp.z * R(1, 1) - pp.y * R(2, 1), A.y * Q(2, 1) + A.z * Q(1, 1) + B.x * Q(0, 2) + B.z * Q(0, 0), S(0),
               -R(2, 1), R(1, 1), 8)
  FCLB_BB_EDGE(pp.z * R(1, 2) - pp.y * R(2, 2), A.y * Q(2, 2) + A.z * Q(1, 2) + B.x * Q(0, 1) + B.y * Q(0, 0), S(0),
               -R(2, 2), R(1, 2), 9)
  // u2 x (v1,v2,v3)
  FCLB_BB_EDGE(pp.x * R(2, 0) - pp.z * R(0, 0), A.x * Q(2, 0) + A.z * Q(0, 0) + B.y * Q(1, 2) + B.z * Q(1, 1), R(2, 0),
               S(0), -R(0, 0), 10)
  FCLB_BB_EDGE(pp.x * R(2, 1) - pp.z * R(0, 1), A.x * Q(2, 1) + A.z * Q(0, 1) + B.x * Q(1, 2) + B.z * Q(1, 0), R(2, 1),
               S(0), -R(0, 1), 11)
  FCLB_BB_EDGE(pp.x * R(2, 2) - pp.z * R(0, 2), A.x * Q(2, 2) + A.z * Q(0, 2) + B.x * Q(1, 1) + B.y * Q(1, 0), R(2, 2),
               S(0), -R(0, 2), 12)
  // u3 x (v1,v2,v3)
  FCLB_BB_EDGE(pp.y * R(0, 0) - pp.x * R(1, 0), A.x * Q(1, 0) + A.y * Q(0, 0) + B.y * Q(2, 2) + B.z * Q(2, 1), -R(1, 0),
               R(0, 0), S(0), 13)
  FCLB_BB_EDGE(pp.y * R(0, 1) - pp.x * R(1, 1), A.x * Q(1, 1) + A.y * Q(0, 1) + B.x * Q(2, 2) + B.z * Q(2, 0), -R(1, 1),
               R(0, 1), S(0), 14)
  FCLB_BB_EDGE(pp.y * R(0, 2) - pp.x * R(1, 2), A.x * Q(1, 2) + A.y * Q(0, 2) + B.x * Q(2, 1) + B.y * Q(2, 0), -R(1, 2),
               R(0, 2), S(0), 15)
#undef FCLB_BB_EDGE

  if (!code) return 0;
  if (SAT_ONLY) return code;

  V3<S> normal;
  if (best_col_id != -1)
    normal = (normalR == 1) ? col(R1, best_col_id) : col(R2, best_col_id);
  else
    normal = mulMV(R1, normalC);
  if (invert_normal) normal = -normal;
  const S depth = -s;

  if (code > 6) {
    V3<S> pa = T1;
    S sign;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      sign = (dot(col(R1, j), normal) > 0) ? S(1) : S(-1);
      pa = pa + col(R1, j) * (comp(A, j) * sign);
    }
    V3<S> pb = T2;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      sign = (dot(col(R2, j), normal) > 0) ? S(-1) : S(1);
      pb = pb + col(R2, j) * (comp(B, j) * sign);
    }
    const V3<S> ua = col(R1, (code - 7) / 3);
    const V3<S> ub = col(R2, (code - 7) % 3);
    // lineClosestApproach, :51-72
    S alpha, beta;
    {
      const V3<S> pd = pb - pa;
      const S uaub = dot(ua, ub);
      const S q1 = dot(ua, pd);
      const S q2 = -dot(ub, pd);
      S d = S(1) - uaub * uaub;
      if (d <= S(0.0001f)) {
        alpha = 0;
        beta = 0;
      } else {
        d = S(1) / d;
        alpha = (q1 + uaub * q2) * d;
        beta = (uaub * q1 + q2) * d;
      }
    }
    pa = pa + ua * alpha;
    pb = pb + ub * beta;
    out[0].normal = normal;
    out[0].pos = (pa + pb) * S(0.5);
    out[0].depth = depth;
    *n_contacts = 1;
    return code;
  }

  // face-something contact: reference face 'a', incident face 'b'
  const bool ref1 = (code <= 3);
  const M3<S>& Ra = ref1 ? R1 : R2;
  const M3<S>& Rb = ref1 ? R2 : R1;
  const V3<S> pa = ref1 ? T1 : T2;
  const V3<S> pb = ref1 ? T2 : T1;
  const V3<S> Sa = ref1 ? A : B;
  const V3<S> Sb = ref1 ? B : A;
  const V3<S> normal2 = ref1 ? normal : -normal;
  const V3<S> nr = mulMtV(Rb, normal2);
  const V3<S> anr = mk<S>(fabs_(nr.x), fabs_(nr.y), fabs_(nr.z));
  int lanr, a1, a2;
  if (anr.y > anr.x) {
    if (anr.y > anr.z) {
      a1 = 0;
      lanr = 1;
      a2 = 2;
    } else {
      a1 = 0;
      a2 = 1;
      lanr = 2;
    }
  } else {
    if (anr.x > anr.z) {
      lanr = 0;
      a1 = 1;
      a2 = 2;
    } else {
      a1 = 0;
      a2 = 1;
      lanr = 2;
    }
  }
  V3<S> center;
  if (comp(nr, lanr) < 0)
    center = pb - pa + col(Rb, lanr) * comp(Sb, lanr);
  else
    center = pb - pa - col(Rb, lanr) * comp(Sb, lanr);

  const int codeN = ref1 ? (code - 1) : (code - 4);
  int code1, code2;
  if (codeN == 0) {
    code1 = 1;
    code2 = 2;
  } else if (codeN == 1) {
    code1 = 0;
    code2 = 2;
  } else {
    code1 = 0;
    code2 = 1;
  }

  S quad[8];
  S c1, c2, m11, m12, m21, m22;
  c1 = dot(col(Ra, code1), center);
  c2 = dot(col(Ra, code2), center);
  V3<S> tempRac = col(Ra, code1);
  m11 = dot(col(Rb, a1), tempRac);
  m12 = dot(col(Rb, a2), tempRac);
  tempRac = col(Ra, code2);
  m21 = dot(col(Rb, a1), tempRac);
  m22 = dot(col(Rb, a2), tempRac);
  {
    const S k1 = m11 * comp(Sb, a1);
    const S k2 = m21 * comp(Sb, a1);
    const S k3 = m12 * comp(Sb, a2);
    const S k4 = m22 * comp(Sb, a2);
    quad[0] = c1 - k1 - k3;
    quad[1] = c2 - k2 - k4;
    quad[2] = c1 - k1 + k3;
    quad[3] = c2 - k2 + k4;
    quad[4] = c1 + k1 + k3;
    quad[5] = c2 + k2 + k4;
    quad[6] = c1 + k1 - k3;
    quad[7] = c2 + k2 - k4;
  }
  S rect[2];
  rect[0] = comp(Sa, code1);
  rect[1] = comp(Sa, code2);

  S ret[16];
  const int n_intersect = intersectRectQuad2(rect, quad, ret);
  if (n_intersect < 1) return code;

  V3<S> points[8];
  S dep[8];
  const S det1 = S(1.f) / (m11 * m22 - m12 * m21);
  m11 *= det1;
  m12 *= det1;
  m21 *= det1;
  m22 *= det1;
  int cnum = 0;
  #pragma unroll 1
  for (int j = 0; j < n_intersect; ++j) {
    const S k1 = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
    const S k2 = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
    points[cnum] = center + col(Rb, a1) * k1 + col(Rb, a2) * k2;
    dep[cnum] = comp(Sa, codeN) - dot(normal2, points[cnum]);
    if (dep[cnum] >= 0) {
      ret[cnum * 2] = ret[j * 2];
      ret[cnum * 2 + 1] = ret[j * 2 + 1];
      cnum++;
    }
  }
  if (cnum < 1) return code;

  if (maxc > cnum) maxc = cnum;
  if (maxc < 1) maxc = 1;
  int iret[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  if (cnum > maxc) {
    int i1 = 0;
    S maxdepth = dep[0];
    #pragma unroll 1
    for (int i = 1; i < cnum; ++i) {
      if (dep[i] > maxdepth) {
        maxdepth = dep[i];
        i1 = i;
      }
    }
    cullPoints2(cnum, ret, maxc, i1, iret);
    cnum = maxc;
  }
  #pragma unroll 1
  for (int j = 0; j < cnum; ++j) {
    const int i = iret[j];
    out[j].normal = normal;
    if (code < 4)
      out[j].pos = points[i] + pa + normal * (dep[i] / S(2));
    else
      out[j].pos = points[i] + pa - normal * (dep[i] / S(2));
    out[j].depth = dep[i];
  }
  *n_contacts = cnum;
  return code;
}

}  // namespace fclb
