// geometry.cu -- batched DLT triangulation and reprojection residuals on sm_100a.
//
// triangulate_kernel : cv::triangulatePoints + float32 de-homogenise, as used by reconstruct()
//                      (OpenCV_SFM/NViewReconstuct.cpp:1146-1156; TwoViewReconstruct.cpp:249).
// residual_kernel    : ReprojectCost::operator() (NViewReconstuct.cpp:151-183) for a batch of
//                      observations, optional ceres::HuberLoss cost (:1184).
//
// Both are HBM-streaming kernels: one thread per point / observation, coalesced vector loads,
// all state in registers, camera tables in shared memory.
#include <cuda_runtime.h>

#include "../../include/sfm_b200.h"
#include <float.h>
#include <stdint.h>

namespace sfm {

constexpr int kMaxViewsSmem = 64;   // projection matrices staged in shared memory

// ---------------------------------------------------------------------------------------
// Per point: A (2V x 4, float64) with rows x*P[2]-P[0], y*P[2]-P[1] per view (the matrix
// cv::triangulatePoints builds); the wanted vector is the right singular vector of the
// smallest singular value = eigenvector of the smallest eigenvalue of M = A^T A (4x4, SPD).
// adj(M) = sum_i (prod_{j != i} lambda_j) v_i v_i^T is dominated by v_4 v_4^T with relative
// weight rho = lambda_4/lambda_3 for the rest, so x = adj(M) e_k (k = largest diagonal cofactor)
// followed by two more products with adj(M) converges to v_4 like rho^3: enough for consistent
// rays (rho ~ 1e-8).  The reference also triangulates ALL matches of later pairs
// (NViewReconstuct.cpp:1441), mismatches included, whose rays do not meet (rho up to ~0.3 on
// the bundled data): the last product doubles as a convergence test, and the rows that fail it
// take null_vector_slow(): adj(M) squared 14 times (rho^16384), which is the SVD answer for
// every system whose two smallest singular values are distinguishable in float64.
// ~190 double FMAs per point on the fast path instead of a Jacobi SVD.
struct Sym4 {
  double m00, m01, m02, m03, m11, m12, m13, m22, m23, m33;
};

__device__ __forceinline__ void sym4_add_row(Sym4& m, double a0, double a1, double a2, double a3) {
  m.m00 = fma(a0, a0, m.m00); m.m01 = fma(a0, a1, m.m01); m.m02 = fma(a0, a2, m.m02);
  m.m03 = fma(a0, a3, m.m03); m.m11 = fma(a1, a1, m.m11); m.m12 = fma(a1, a2, m.m12);
  m.m13 = fma(a1, a3, m.m13); m.m22 = fma(a2, a2, m.m22); m.m23 = fma(a2, a3, m.m23);
  m.m33 = fma(a3, a3, m.m33);
}

// adjugate of a symmetric 4x4 (10 unique cofactors) through 2x2 minors
__device__ __forceinline__ Sym4 sym4_adjugate(const Sym4& a) {
  // rows 0,1 minors (columns i<j): s; rows 2,3 minors: c   (M symmetric: a10=a01 ...)
  const double a00 = a.m00, a01 = a.m01, a02 = a.m02, a03 = a.m03;
  const double a10 = a.m01, a11 = a.m11, a12 = a.m12, a13 = a.m13;
  const double a20 = a.m02, a21 = a.m12, a22 = a.m22, a23 = a.m23;
  const double a30 = a.m03, a31 = a.m13, a32 = a.m23, a33 = a.m33;
  const double s0 = a00 * a11 - a10 * a01;
  const double s1 = a00 * a12 - a10 * a02;
  const double s2 = a00 * a13 - a10 * a03;
  const double s3 = a01 * a12 - a11 * a02;
  const double s4 = a01 * a13 - a11 * a03;
  const double s5 = a02 * a13 - a12 * a03;
  const double c5 = a22 * a33 - a32 * a23;
  const double c4 = a21 * a33 - a31 * a23;
  const double c3 = a21 * a32 - a31 * a22;
  const double c2 = a20 * a33 - a30 * a23;
  const double c1 = a20 * a32 - a30 * a22;
  const double c0 = a20 * a31 - a30 * a21;
  Sym4 r;
  r.m00 = a11 * c5 - a12 * c4 + a13 * c3;
  r.m01 = -a01 * c5 + a02 * c4 - a03 * c3;
  r.m02 = a31 * s5 - a32 * s4 + a33 * s3;
  r.m03 = -a21 * s5 + a22 * s4 - a23 * s3;
  r.m11 = a00 * c5 - a02 * c2 + a03 * c1;
  r.m12 = -a30 * s5 + a32 * s2 - a33 * s1;
  r.m13 = a20 * s5 - a22 * s2 + a23 * s1;
  r.m22 = a30 * s4 - a31 * s2 + a33 * s0;
  r.m23 = -a20 * s4 + a21 * s2 - a23 * s0;
  r.m33 = a20 * s3 - a21 * s1 + a22 * s0;
  return r;
}

__device__ __forceinline__ void sym4_mul(const Sym4& a, double& x0, double& x1, double& x2,
                                         double& x3) {
  const double y0 = a.m00 * x0 + a.m01 * x1 + a.m02 * x2 + a.m03 * x3;
  const double y1 = a.m01 * x0 + a.m11 * x1 + a.m12 * x2 + a.m13 * x3;
  const double y2 = a.m02 * x0 + a.m12 * x1 + a.m22 * x2 + a.m23 * x3;
  const double y3 = a.m03 * x0 + a.m13 * x1 + a.m23 * x2 + a.m33 * x3;
  x0 = y0; x1 = y1; x2 = y2; x3 = y3;
}

// 2^-e for x = m * 2^e (m in [1,2)): exact scaling factor from the exponent bits; 1 for
// zero / denormal / non-finite input.
__device__ __forceinline__ double pow2_inv(double x) {
  const int hi = __double2hiint(x);
  const int e = (hi >> 20) & 0x7ff;
  const int ne = 2046 - e;                       // biased exponent of 2^-(e - 1023)
  return (e > 0 && e < 2046) ? __hiloint2double(ne << 20, 0) : 1.0;
}

// Rows whose fast iteration has not converged (lambda_4/lambda_3 not small: noisy or mismatched
// rays): B <- B^2 (rescaled by an exact power of two) squares the eigenvalue ratio each time.
// (tr(B)^2 - tr(B^2)) / 2 is the sum of the principal 2x2 minors of B ~ mu_1 mu_2, so its ratio
// to tr(B)^2 measures the current eigenvalue ratio mu_2/mu_1 of B for free: the loop stops as
// soon as B is rank one to kTriRank (2-4 squarings for noisy inliers, at most kTriSquarings,
// i.e. rho^16384, for rays that miss each other).  The column of the largest diagonal entry is
// then v_4 to float64 accuracy.  Out of line: rare, and its second matrix must not cost the
// fast path registers.
constexpr double kTriTol = 1e-9;
constexpr double kTriRank = 1e-12;
constexpr int kTriSquarings = 14;
struct Vec4d {
  double x0, x1, x2, x3;
};
__device__ __noinline__ Vec4d null_vector_slow(Sym4 b) {   // by value in, by value out: registers
  for (int k = 0; k < kTriSquarings; ++k) {
    const double sc = pow2_inv(b.m00 + b.m11 + b.m22 + b.m33);   // trace > 0: B is PSD
    b.m00 *= sc; b.m01 *= sc; b.m02 *= sc; b.m03 *= sc; b.m11 *= sc;
    b.m12 *= sc; b.m13 *= sc; b.m22 *= sc; b.m23 *= sc; b.m33 *= sc;
    Sym4 r;
    r.m00 = b.m00 * b.m00 + b.m01 * b.m01 + b.m02 * b.m02 + b.m03 * b.m03;
    r.m01 = b.m00 * b.m01 + b.m01 * b.m11 + b.m02 * b.m12 + b.m03 * b.m13;
    r.m02 = b.m00 * b.m02 + b.m01 * b.m12 + b.m02 * b.m22 + b.m03 * b.m23;
    r.m03 = b.m00 * b.m03 + b.m01 * b.m13 + b.m02 * b.m23 + b.m03 * b.m33;
    r.m11 = b.m01 * b.m01 + b.m11 * b.m11 + b.m12 * b.m12 + b.m13 * b.m13;
    r.m12 = b.m01 * b.m02 + b.m11 * b.m12 + b.m12 * b.m22 + b.m13 * b.m23;
    r.m13 = b.m01 * b.m03 + b.m11 * b.m13 + b.m12 * b.m23 + b.m13 * b.m33;
    r.m22 = b.m02 * b.m02 + b.m12 * b.m12 + b.m22 * b.m22 + b.m23 * b.m23;
    r.m23 = b.m02 * b.m03 + b.m12 * b.m13 + b.m22 * b.m23 + b.m23 * b.m33;
    r.m33 = b.m03 * b.m03 + b.m13 * b.m13 + b.m23 * b.m23 + b.m33 * b.m33;
    const double tb = b.m00 + b.m11 + b.m22 + b.m33, tr = r.m00 + r.m11 + r.m22 + r.m33;
    b = r;
    if (tb * tb - tr <= kTriRank * tb * tb) break;               // B was already rank one
  }
  double best = b.m00;
  Vec4d v = {b.m00, b.m01, b.m02, b.m03};
  if (b.m11 > best) { best = b.m11; v.x0 = b.m01; v.x1 = b.m11; v.x2 = b.m12; v.x3 = b.m13; }
  if (b.m22 > best) { best = b.m22; v.x0 = b.m02; v.x1 = b.m12; v.x2 = b.m22; v.x3 = b.m23; }
  if (b.m33 > best) { best = b.m33; v.x0 = b.m03; v.x1 = b.m13; v.x2 = b.m23; v.x3 = b.m33; }
  return v;
}

// Projection matrices of up to kMaxViewsParam views travel as a kernel parameter: the fp64
// FMAs then take them straight from the constant bank (no shared-memory loads, no registers).
constexpr int kMaxViewsParam = 8;
struct ProjParam {
  double p[kMaxViewsParam * 12];
};

template <bool kParamP, int kV>     // kV > 0: view count known at compile time (the reference: 2)
__global__ void __launch_bounds__(256)
triangulate_kernel(const __grid_constant__ ProjParam pp, const float* __restrict__ P,
                   const float* __restrict__ xy, int n_views, int64_t n_pts,
                   float* __restrict__ X4, double* __restrict__ xyz) {
  __shared__ double sP[kParamP ? 1 : kMaxViewsSmem * 12];
  const bool p_in_smem = !kParamP && n_views <= kMaxViewsSmem;
  if (p_in_smem) {
    for (int i = threadIdx.x; i < n_views * 12; i += blockDim.x) sP[i] = static_cast<double>(P[i]);
    __syncthreads();
  }
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_pts;
       i += stride) {
    Sym4 m = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int v = 0; v < (kV > 0 ? kV : n_views); ++v) {
      const float2 p = __ldcs(reinterpret_cast<const float2*>(xy) + v * n_pts + i);
      const double x = p.x, y = p.y;
      double q[12];
#pragma unroll
      for (int k = 0; k < 12; ++k)
        q[k] = kParamP ? pp.p[v * 12 + k]
                       : (p_in_smem ? sP[v * 12 + k] : static_cast<double>(P[v * 12 + k]));
      sym4_add_row(m, x * q[8] - q[0], x * q[9] - q[1], x * q[10] - q[2], x * q[11] - q[3]);
      sym4_add_row(m, y * q[8] - q[4], y * q[9] - q[5], y * q[10] - q[6], y * q[11] - q[7]);
    }
    // scale so that cofactors (cubic in M) and three products with adj(M) stay far from overflow
    // and underflow for any pixel scale: one exact power of two from the exponent of the trace
    // (no division); after it |adj| <= 6 and nothing in between needs renormalising
    const double tr = m.m00 + m.m11 + m.m22 + m.m33;
    const double sc = pow2_inv(tr);
    m.m00 *= sc; m.m01 *= sc; m.m02 *= sc; m.m03 *= sc; m.m11 *= sc;
    m.m12 *= sc; m.m13 *= sc; m.m22 *= sc; m.m23 *= sc; m.m33 *= sc;
    const Sym4 adj = sym4_adjugate(m);
    // column k of the largest diagonal cofactor (|v4[k]| largest): no cancellation in the start
    double x0 = adj.m00, x1 = adj.m01, x2 = adj.m02, x3 = adj.m03, best = fabs(adj.m00);
    int k = 0;
    if (fabs(adj.m11) > best) { best = fabs(adj.m11); k = 1; x0 = adj.m01; x1 = adj.m11; x2 = adj.m12; x3 = adj.m13; }
    if (fabs(adj.m22) > best) { best = fabs(adj.m22); k = 2; x0 = adj.m02; x1 = adj.m12; x2 = adj.m22; x3 = adj.m23; }
    if (fabs(adj.m33) > best) { best = fabs(adj.m33); k = 3; x0 = adj.m03; x1 = adj.m13; x2 = adj.m23; x3 = adj.m33; }
    sym4_mul(adj, x0, x1, x2, x3);
    {
      // last product = convergence test.  y = adj x is parallel to x iff x_k y_i - x_i y_k = 0
      // for all i, k the component where v4 is largest (x_k y_k carries the scale):
      // sum_i (x_k y_i - x_i y_k)^2 <= tol^2 (x_k y_k)^2 bounds the direction change by ~2 tol.
      double y0 = x0, y1 = x1, y2 = x2, y3 = x3;
      sym4_mul(adj, y0, y1, y2, y3);
      const double xk = k == 0 ? x0 : (k == 1 ? x1 : (k == 2 ? x2 : x3));
      const double yk = k == 0 ? y0 : (k == 1 ? y1 : (k == 2 ? y2 : y3));
      const double c0 = xk * y0 - x0 * yk, c1 = xk * y1 - x1 * yk;
      const double c2 = xk * y2 - x2 * yk, c3 = xk * y3 - x3 * yk;
      const double cross2 = c0 * c0 + c1 * c1 + c2 * c2 + c3 * c3;
      const double ref = kTriTol * xk * yk;
      x0 = y0; x1 = y1; x2 = y2; x3 = y3;
      if (!(cross2 <= ref * ref)) {
        const Vec4d v = null_vector_slow(adj);
        x0 = v.x0; x1 = v.x1; x2 = v.x2; x3 = v.x3;
      }
    }
    best = fmax(fmax(fabs(x0), fabs(x1)), fmax(fabs(x2), fabs(x3)));
    const double inv2 = pow2_inv(best);
    x0 *= inv2; x1 *= inv2; x2 *= inv2; x3 *= inv2;
    const double n2 = x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3;   // in [1, 4] after the scaling
    // 1/sqrt(n2): float seed + one Newton step in double (2^-22 -> 2^-43: the result is rounded to
    // float anyway); a zero vector (degenerate input) stays zero
    double inv = 0.0;
    if (n2 > 0.0) {
      const double r0 = static_cast<double>(rsqrtf(static_cast<float>(n2)));
      inv = r0 * (1.5 - 0.5 * n2 * r0 * r0);
    }
    // cv::triangulatePoints returns the points' dtype: float32 (pts2d are Point2f, :1147)
    const float f0 = static_cast<float>(x0 * inv), f1 = static_cast<float>(x1 * inv);
    const float f2 = static_cast<float>(x2 * inv), f3 = static_cast<float>(x3 * inv);
    if (X4 != nullptr) {
      __stcs(X4 + i, f0);
      __stcs(X4 + n_pts + i, f1);
      __stcs(X4 + 2 * n_pts + i, f2);
      __stcs(X4 + 3 * n_pts + i, f3);
    }
    if (xyz != nullptr) {
      // pt4d_homo /= pt4d_homo(3) (:1154) on a Mat_<float> is OpenCV's a.convertTo(a, -1, 1./b):
      // every element is MULTIPLIED by float(1.0 / w), not divided by w (this form reproduces
      // 1835 of the 1847 two-view points of the bundled Viewer/structure.yml bit for bit, the
      // quotient form 758); then Point3f -> Point3d (:1155).  __frcp_rn(w) = RN_float(1/w) equals
      // float(RN_double(1/w)) except when 1/w lies within 2^-54 of a float rounding boundary.
      const float rw = __frcp_rn(f3);
      __stcs(xyz + 3 * i + 0, static_cast<double>(f0 * rw));
      __stcs(xyz + 3 * i + 1, static_cast<double>(f1 * rw));
      __stcs(xyz + 3 * i + 2, static_cast<double>(f2 * rw));
    }
  }
}

// ---------------------------------------------------------------------------------------
// Camera table: rotation matrix of ceres::AngleAxisRotatePoint (same two branches around
// theta^2 <= DBL_EPSILON) + translation, 12 doubles per camera.
__global__ void camera_table_kernel(const double* __restrict__ ext, int n_cam,
                                    double* __restrict__ cam) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cam) return;
  const double w0 = ext[6 * c + 0], w1 = ext[6 * c + 1], w2 = ext[6 * c + 2];
  const double theta2 = w0 * w0 + w1 * w1 + w2 * w2;
  double R[9];
  if (theta2 > DBL_EPSILON) {
    const double theta = sqrt(theta2);
    const double ct = cos(theta), st = sin(theta);
    const double ti = 1.0 / theta;
    const double a = w0 * ti, b = w1 * ti, g = w2 * ti;
    const double oc = 1.0 - ct;
    // p = X cos + (w x X) sin + w (w.X)(1-cos)
    R[0] = ct + a * a * oc;      R[1] = -g * st + a * b * oc; R[2] = b * st + a * g * oc;
    R[3] = g * st + b * a * oc;  R[4] = ct + b * b * oc;      R[5] = -a * st + b * g * oc;
    R[6] = -b * st + g * a * oc; R[7] = a * st + g * b * oc;  R[8] = ct + g * g * oc;
  } else {
    // p = X + w x X
    R[0] = 1.0; R[1] = -w2; R[2] = w1;
    R[3] = w2;  R[4] = 1.0; R[5] = -w0;
    R[6] = -w1; R[7] = w0;  R[8] = 1.0;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) cam[12 * c + k] = R[k];
  cam[12 * c + 9] = ext[6 * c + 3];
  cam[12 * c + 10] = ext[6 * c + 4];
  cam[12 * c + 11] = ext[6 * c + 5];
}

__device__ __forceinline__ double huber_rho(double s, double delta) {
  // ceres::HuberLoss(a): rho(s) = s for s <= a^2, 2 a sqrt(s) - a^2 otherwise
  if (delta <= 0.0) return s;
  const double b = delta * delta;
  return s <= b ? s : 2.0 * delta * sqrt(s) - b;
}

__device__ __forceinline__ double2 residual_one(double fx, double fy, double cx, double cy,
                                                const double* __restrict__ cam,
                                                const double* __restrict__ pts, int c, int j,
                                                float2 o) {
  const double* R = cam + 12 * static_cast<int64_t>(c);
  const double* Xp = pts + 3 * static_cast<int64_t>(j);
  const double X = __ldg(Xp), Y = __ldg(Xp + 1), Z = __ldg(Xp + 2);
  const double p0 = __ldg(R + 0) * X + __ldg(R + 1) * Y + __ldg(R + 2) * Z + __ldg(R + 9);
  const double p1 = __ldg(R + 3) * X + __ldg(R + 4) * Y + __ldg(R + 5) * Z + __ldg(R + 10);
  const double p2 = __ldg(R + 6) * X + __ldg(R + 7) * Y + __ldg(R + 8) * Z + __ldg(R + 11);
  // xp = p0 / p2, yp = p1 / p2 (:166-167); the two quotients are kept as true divisions so
  // that the result rounds exactly like the reference's double arithmetic
  const double x = p0 / p2, y = p1 / p2;
  return make_double2(fx * x + cx - static_cast<double>(o.x),
                      fy * y + cy - static_cast<double>(o.y));
}

__global__ void __launch_bounds__(256)
residual_kernel(double fx, double fy, double cx, double cy, const double* __restrict__ cam,
                const double* __restrict__ pts, const int32_t* __restrict__ cam_idx,
                const int32_t* __restrict__ pt_idx, const float* __restrict__ obs_xy,
                int64_t n_obs, double huber_delta, double* __restrict__ resid,
                double* __restrict__ block_cost) {
  double cost = 0.0;
  // two observations per thread and iteration: 8-byte index loads, one 16-byte observation
  // load, two 16-byte residual stores in flight
  const int64_t n2 = n_obs >> 1;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  // software pipeline: the loads of the next iteration are issued before the (dependent) point
  // gathers of this one
  int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int2 c = make_int2(0, 0), j = make_int2(0, 0);
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k < n2) {
    c = __ldcs(reinterpret_cast<const int2*>(cam_idx) + k);
    j = __ldcs(reinterpret_cast<const int2*>(pt_idx) + k);
    o = __ldcs(reinterpret_cast<const float4*>(obs_xy) + k);
  }
  while (k < n2) {
    const int64_t kn = k + stride;
    int2 cn = c, jn = j;
    float4 on = o;
    if (kn < n2) {
      cn = __ldcs(reinterpret_cast<const int2*>(cam_idx) + kn);
      jn = __ldcs(reinterpret_cast<const int2*>(pt_idx) + kn);
      on = __ldcs(reinterpret_cast<const float4*>(obs_xy) + kn);
    }
    const double2 r0 = residual_one(fx, fy, cx, cy, cam, pts, c.x, j.x, make_float2(o.x, o.y));
    const double2 r1 = residual_one(fx, fy, cx, cy, cam, pts, c.y, j.y, make_float2(o.z, o.w));
    if (resid != nullptr) {
      __stcs(reinterpret_cast<double2*>(resid) + 2 * k, r0);
      __stcs(reinterpret_cast<double2*>(resid) + 2 * k + 1, r1);
    }
    if (block_cost != nullptr)
      cost += huber_rho(r0.x * r0.x + r0.y * r0.y, huber_delta) +
              huber_rho(r1.x * r1.x + r1.y * r1.y, huber_delta);
    k = kn;
    c = cn;
    j = jn;
    o = on;
  }
  if ((n_obs & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t k = n_obs - 1;
    const float2 o = reinterpret_cast<const float2*>(obs_xy)[k];
    const double2 r = residual_one(fx, fy, cx, cy, cam, pts, cam_idx[k], pt_idx[k], o);
    if (resid != nullptr) reinterpret_cast<double2*>(resid)[k] = r;
    if (block_cost != nullptr) cost += huber_rho(r.x * r.x + r.y * r.y, huber_delta);
  }
  if (block_cost != nullptr) {
    __shared__ double s_part[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = cost;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += s_part[w];
      block_cost[blockIdx.x] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Camera-major observation lists (what bundle_adjustment() builds, NViewReconstuct.cpp:1187-1211)
// visit the point table once per camera; processed in list order every camera streams the whole
// table from HBM again (24 B per observation on top of the 32 algorithmic ones: measured 0.58 of
// the copy rate at 8 views).  When cam_idx is sorted the kernel below walks the list in
// (chunk of kSegChunk positions) x (camera) order instead, so the cameras read one chunk of
// points while it is still in L2.  Results are written to the caller's positions: the order of
// the residual vector does not change.
constexpr int kSegChunk = 2048;

// flag |= 1 when cam_idx is not non-decreasing
__global__ void __launch_bounds__(256)
sorted_check_kernel(const int32_t* __restrict__ cam_idx, int64_t n_obs, uint32_t* __restrict__ flag) {
  bool bad = false;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k + 1 < n_obs; k += stride)
    bad |= __ldg(cam_idx + k) > __ldg(cam_idx + k + 1);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

// seg[c] = first observation of camera c (lower bound in the sorted cam_idx), seg[n_cam] = n_obs
__global__ void segment_table_kernel(const int32_t* __restrict__ cam_idx, int64_t n_obs, int n_cam,
                                     int64_t* __restrict__ seg) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_cam) return;
  int64_t lo = 0, hi = n_obs;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (cam_idx[mid] < c) lo = mid + 1; else hi = mid;
  }
  seg[c] = lo;
}

__global__ void __launch_bounds__(256)
residual_seg_kernel(double fx, double fy, double cx, double cy, const double* __restrict__ cam,
                    const double* __restrict__ pts, const int64_t* __restrict__ seg, int n_cam,
                    int64_t max_seg, const int32_t* __restrict__ pt_idx,
                    const float* __restrict__ obs_xy, double huber_delta,
                    double* __restrict__ resid, double* __restrict__ block_cost) {
  double cost = 0.0;
  const int n_chunks = static_cast<int>((max_seg + kSegChunk - 1) / kSegChunk);
  const int n_blocks = n_chunks * n_cam;                          // (chunk, camera) blocks, camera fastest
  for (int b = blockIdx.x; b < n_blocks; b += gridDim.x) {
    const int c = b % n_cam;
    const int64_t k0 = __ldg(seg + c) + static_cast<int64_t>(b / n_cam) * kSegChunk;
    const int64_t k1 = min(k0 + kSegChunk, __ldg(seg + c + 1));
    if (((k0 | k1) & 1) == 0) {
      // two observations per thread and iteration: 8-byte index loads, one 16-byte observation
      // load, two 16-byte residual stores in flight
      // software pipeline: the index / observation loads of the next iteration are issued before
      // the (dependent) point gathers of this one
      const int64_t kend = k1 >> 1;
      int64_t k = (k0 >> 1) + threadIdx.x;
      int2 j = make_int2(0, 0);
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < kend) {
        j = __ldcs(reinterpret_cast<const int2*>(pt_idx) + k);
        o = __ldcs(reinterpret_cast<const float4*>(obs_xy) + k);
      }
      while (k < kend) {
        const int64_t kn = k + 256;
        int2 jn = j;
        float4 on = o;
        if (kn < kend) {
          jn = __ldcs(reinterpret_cast<const int2*>(pt_idx) + kn);
          on = __ldcs(reinterpret_cast<const float4*>(obs_xy) + kn);
        }
        const double2 r0 = residual_one(fx, fy, cx, cy, cam, pts, c, j.x, make_float2(o.x, o.y));
        const double2 r1 = residual_one(fx, fy, cx, cy, cam, pts, c, j.y, make_float2(o.z, o.w));
        if (resid != nullptr) {
          __stcs(reinterpret_cast<double2*>(resid) + 2 * k, r0);
          __stcs(reinterpret_cast<double2*>(resid) + 2 * k + 1, r1);
        }
        if (block_cost != nullptr)
          cost += huber_rho(r0.x * r0.x + r0.y * r0.y, huber_delta) +
                  huber_rho(r1.x * r1.x + r1.y * r1.y, huber_delta);
        k = kn;
        j = jn;
        o = on;
      }
    } else {
      for (int64_t k = k0 + threadIdx.x; k < k1; k += 256) {
        const float2 o = __ldcs(reinterpret_cast<const float2*>(obs_xy) + k);
        const double2 r = residual_one(fx, fy, cx, cy, cam, pts, c, __ldcs(pt_idx + k), o);
        if (resid != nullptr) __stcs(reinterpret_cast<double2*>(resid) + k, r);
        if (block_cost != nullptr) cost += huber_rho(r.x * r.x + r.y * r.y, huber_delta);
      }
    }
  }
  if (block_cost != nullptr) {
    __shared__ double s_part[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = cost;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += s_part[w];
      block_cost[blockIdx.x] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Jacobians of ReprojectCost (NViewReconstuct.cpp:151-183) with respect to its three
// parameter blocks, in the order Ceres sees them (:1202-1209): intrinsic (fx, fy, cx, cy),
// extrinsic (angle-axis, translation), point.  This is what AutoDiffCostFunction<ReprojectCost,
// 2, 4, 6, 3> derives with Jets; here it is the closed form of the same function, including
// the small-angle branch of ceres::AngleAxisRotatePoint (p = X + w x X).
//
// Second camera table: 8 doubles per camera {w_hat(3), theta, sin, cos, big-angle flag, 0}.
__global__ void camera_jac_table_kernel(const double* __restrict__ ext, int n_cam,
                                        double* __restrict__ tab) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cam) return;
  const double w0 = ext[6 * c + 0], w1 = ext[6 * c + 1], w2 = ext[6 * c + 2];
  const double theta2 = w0 * w0 + w1 * w1 + w2 * w2;
  double* o = tab + 8 * c;
  if (theta2 > DBL_EPSILON) {
    const double theta = sqrt(theta2), ti = 1.0 / theta;
    o[0] = w0 * ti; o[1] = w1 * ti; o[2] = w2 * ti;
    o[3] = theta; o[4] = sin(theta); o[5] = cos(theta); o[6] = 1.0;
  } else {
    o[0] = w0; o[1] = w1; o[2] = w2;
    o[3] = 0.0; o[4] = 0.0; o[5] = 1.0; o[6] = 0.0;
  }
  o[7] = 0.0;
}

constexpr int kJacThreads = 128;
constexpr int kJacCols = 13;                       // 4 + 6 + 3 parameters
constexpr int kJacRow = 2 * kJacCols;              // doubles per observation
constexpr int kJacPad = kJacRow + 1;               // shared-memory row stride (bank spread)

// One observation per thread; the 26 doubles of every observation are staged in shared
// memory so that the block writes its contiguous 128 x 208-byte output range coalesced.
__global__ void __launch_bounds__(kJacThreads)
jacobian_kernel(double fx, double fy, double cx, double cy, const double* __restrict__ cam,
                const double* __restrict__ jtab, const double* __restrict__ pts,
                const int32_t* __restrict__ cam_idx, const int32_t* __restrict__ pt_idx,
                const float* __restrict__ obs_xy, int64_t n_obs, double* __restrict__ resid,
                double* __restrict__ jac) {
  __shared__ double s_j[kJacThreads * kJacPad];
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * kJacThreads; base < n_obs;
       base += static_cast<int64_t>(gridDim.x) * kJacThreads) {
    const int64_t k = base + threadIdx.x;
    if (k < n_obs) {
      const int c = cam_idx[k], j = pt_idx[k];
      const double* R = cam + 12 * static_cast<int64_t>(c);
      const double* W = jtab + 8 * static_cast<int64_t>(c);
      const double* Xp = pts + 3 * static_cast<int64_t>(j);
      const double X[3] = {__ldg(Xp), __ldg(Xp + 1), __ldg(Xp + 2)};
      double Rm[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) Rm[i] = __ldg(R + i);
      const double p0 = Rm[0] * X[0] + Rm[1] * X[1] + Rm[2] * X[2] + __ldg(R + 9);
      const double p1 = Rm[3] * X[0] + Rm[4] * X[1] + Rm[5] * X[2] + __ldg(R + 10);
      const double p2 = Rm[6] * X[0] + Rm[7] * X[1] + Rm[8] * X[2] + __ldg(R + 11);
      const double x = p0 / p2, y = p1 / p2, iz = 1.0 / p2;
      if (resid != nullptr) {
        const float2 o = reinterpret_cast<const float2*>(obs_xy)[k];
        reinterpret_cast<double2*>(resid)[k] = make_double2(fx * x + cx - static_cast<double>(o.x),
                                                            fy * y + cy - static_cast<double>(o.y));
      }
      // d(residual)/d(p): rows (fx/z, 0, -fx x/z), (0, fy/z, -fy y/z)
      const double a00 = fx * iz, a02 = -fx * x * iz, a11 = fy * iz, a12 = -fy * y * iz;
      double* o = s_j + threadIdx.x * kJacPad;
      // intrinsic block
      o[0] = x;   o[1] = 0.0; o[2] = 1.0; o[3] = 0.0;
      o[kJacCols + 0] = 0.0; o[kJacCols + 1] = y; o[kJacCols + 2] = 0.0; o[kJacCols + 3] = 1.0;
      // extrinsic block: angle-axis
      const double wh[3] = {__ldg(W), __ldg(W + 1), __ldg(W + 2)};
      const double theta = __ldg(W + 3), st = __ldg(W + 4), ct = __ldg(W + 5);
      const bool big = __ldg(W + 6) != 0.0;
      const double wxX[3] = {wh[1] * X[2] - wh[2] * X[1], wh[2] * X[0] - wh[0] * X[2],
                             wh[0] * X[1] - wh[1] * X[0]};
      const double d = wh[0] * X[0] + wh[1] * X[1] + wh[2] * X[2];
      const double ti = big ? 1.0 / theta : 0.0, oc = 1.0 - ct;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        // e_q x X
        const double exX[3] = {q == 1 ? X[2] : (q == 2 ? -X[1] : 0.0),
                               q == 2 ? X[0] : (q == 0 ? -X[2] : 0.0),
                               q == 0 ? X[1] : (q == 1 ? -X[0] : 0.0)};
        double dp[3];
        if (big) {
          // g = (e_q - w_hat w_hat_q) / theta;  dp = -s w_q X + s (g x X) + c w_q (w_hat x X)
          //      + (1 - c) (g d + w_hat (g . X)) + s w_q d w_hat
          const double wq = wh[q];
          const double gX = (X[q] - d * wq) * ti;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const double g_i = ((i == q ? 1.0 : 0.0) - wh[i] * wq) * ti;
            const double gxX_i = (exX[i] - wq * wxX[i]) * ti;
            dp[i] = -st * wq * X[i] + st * gxX_i + ct * wq * wxX[i] + oc * (g_i * d + wh[i] * gX) +
                    st * wq * d * wh[i];
          }
        } else {
          dp[0] = exX[0]; dp[1] = exX[1]; dp[2] = exX[2];
        }
        o[4 + q] = a00 * dp[0] + a02 * dp[2];
        o[kJacCols + 4 + q] = a11 * dp[1] + a12 * dp[2];
      }
      // extrinsic block: translation (dp/dt = I)
      o[7] = a00; o[8] = 0.0; o[9] = a02;
      o[kJacCols + 7] = 0.0; o[kJacCols + 8] = a11; o[kJacCols + 9] = a12;
      // point block (dp/dX = R)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        o[10 + q] = a00 * Rm[q] + a02 * Rm[6 + q];
        o[kJacCols + 10 + q] = a11 * Rm[3 + q] + a12 * Rm[6 + q];
      }
    }
    __syncthreads();
    const int64_t rows = min(static_cast<int64_t>(kJacThreads), n_obs - base);
    double* dst = jac + base * kJacRow;
    for (int64_t e = threadIdx.x; e < rows * kJacRow; e += kJacThreads)
      __stcs(dst + e, s_j[(e / kJacRow) * kJacPad + (e % kJacRow)]);
    __syncthreads();
  }
}

// fixed-order sum of the per-block partial costs (deterministic), result = 0.5 * sum
__global__ void __launch_bounds__(256)
cost_sum_kernel(const double* __restrict__ block_cost, int n, double* __restrict__ out) {
  __shared__ double s[256];
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) t += block_cost[i];
  s[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = 0.5 * s[0];
}

// ---------------------------------------------------------------------------------------
// estimate_normals (NViewReconstuct.cpp:551-599) + PCAFitPlane (:601-690): for every point the
// K nearest other points (brute force, fp64 distances as the reference's Pt3dDist compares them),
// the covariance of those K neighbours about their own mean, its eigenvector of smallest
// eigenvalue, flipped so that normal . centroid <= 0 (towards the camera at the origin, :672-677)
// and normalised.  One thread per point; candidates stream through shared memory in tiles; the
// running top-K is a sorted register array (insertion only when a candidate beats the K-th).
constexpr int kNrmThreads = 128;
constexpr int kNrmTile = 256;
constexpr int kNrmKMax = 16;

__device__ __forceinline__ void jacobi_rotate(double (&A)[3][3], double (&V)[3][3], int p, int q) {
  if (fabs(A[p][q]) < 1e-300) return;
  const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
  const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
  const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
  for (int k = 0; k < 3; ++k) {                       // A <- A J
    const double akp = A[k][p], akq = A[k][q];
    A[k][p] = c * akp - s * akq;
    A[k][q] = s * akp + c * akq;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {                       // A <- J^T A
    const double apk = A[p][k], aqk = A[q][k];
    A[p][k] = c * apk - s * aqk;
    A[q][k] = s * apk + c * aqk;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {                       // V <- V J
    const double vkp = V[k][p], vkq = V[k][q];
    V[k][p] = c * vkp - s * vkq;
    V[k][q] = s * vkp + c * vkq;
  }
}

__global__ void __launch_bounds__(kNrmThreads)
normals_kernel(const double* __restrict__ pts, int n, int K, double* __restrict__ normals) {
  __shared__ double s_p[kNrmTile][3];
  const int i = blockIdx.x * kNrmThreads + threadIdx.x;
  const int ic = min(i, n - 1);
  const double px = pts[3 * ic], py = pts[3 * ic + 1], pz = pts[3 * ic + 2];
  double bd[kNrmKMax];
  int bi[kNrmKMax];
#pragma unroll
  for (int k = 0; k < kNrmKMax; ++k) { bd[k] = DBL_MAX; bi[k] = -1; }
  double worst = DBL_MAX;                              // current K-th smallest distance^2
  for (int j0 = 0; j0 < n; j0 += kNrmTile) {
    __syncthreads();
    for (int e = threadIdx.x; e < kNrmTile * 3; e += kNrmThreads)
      if (j0 * 3 + e < n * 3) (&s_p[0][0])[e] = pts[static_cast<size_t>(j0) * 3 + e];
    __syncthreads();
    const int m = min(kNrmTile, n - j0);
    for (int r = 0; r < m; ++r) {
      const double dx = px - s_p[r][0], dy = py - s_p[r][1], dz = pz - s_p[r][2];
      // (dx*dx + dy*dy) + dz*dz without contraction: the reference's expression order (:470-472)
      const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      if (d < worst && j0 + r != ic) {
        // sorted insertion; entries beyond K-1 are never read
        double cd = d;
        int ci = j0 + r;
#pragma unroll
        for (int k = 0; k < kNrmKMax; ++k) {
          if (k < K && cd < bd[k]) {
            const double td = bd[k]; const int ti = bi[k];
            bd[k] = cd; bi[k] = ci;
            cd = td; ci = ti;
          }
        }
        worst = DBL_MAX;
#pragma unroll
        for (int k = 0; k < kNrmKMax; ++k)
          if (k == K - 1) worst = bd[k];
      }
    }
  }
  if (i >= n) return;
  // PCAFitPlane: mean and covariance of the K neighbours (the point itself is not included)
  double mx = 0.0, my = 0.0, mz = 0.0;
#pragma unroll
  for (int k = 0; k < kNrmKMax; ++k)
    if (k < K) { mx += pts[3 * bi[k]]; my += pts[3 * bi[k] + 1]; mz += pts[3 * bi[k] + 2]; }
  const double inv = 1.0 / static_cast<double>(K);
  mx *= inv; my *= inv; mz *= inv;
  double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
  for (int k = 0; k < kNrmKMax; ++k)
    if (k < K) {
      const double x = pts[3 * bi[k]] - mx, y = pts[3 * bi[k] + 1] - my, z = pts[3 * bi[k] + 2] - mz;
      A[0][0] += x * x; A[1][1] += y * y; A[2][2] += z * z;
      A[0][1] += x * y; A[0][2] += x * z; A[1][2] += y * z;
    }
  A[0][0] *= inv; A[1][1] *= inv; A[2][2] *= inv;
  A[0][1] *= inv; A[0][2] *= inv; A[1][2] *= inv;
  A[1][0] = A[0][1]; A[2][0] = A[0][2]; A[2][1] = A[1][2];
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 12; ++sweep) {           // cyclic Jacobi: quadratic convergence
    jacobi_rotate(A, V, 0, 1);
    jacobi_rotate(A, V, 0, 2);
    jacobi_rotate(A, V, 1, 2);
  }
  int mn = 0;
  if (A[1][1] < A[mn][mn]) mn = 1;
  if (A[2][2] < A[mn][mn]) mn = 2;
  double a = mn == 0 ? V[0][0] : (mn == 1 ? V[0][1] : V[0][2]);
  double b = mn == 0 ? V[1][0] : (mn == 1 ? V[1][1] : V[1][2]);
  double c = mn == 0 ? V[2][0] : (mn == 1 ? V[2][1] : V[2][2]);
  if (a * mx + b * my + c * mz > 0.0) { a = -a; b = -b; c = -c; }     // :672-677
  const double den = sqrt(a * a + b * b + c * c);
  normals[3 * i] = a / den;
  normals[3 * i + 1] = b / den;
  normals[3 * i + 2] = c / den;
}

cudaError_t launch_normals(const double* pts, int n, int K, double* normals, cudaStream_t s) {
  normals_kernel<<<(n + kNrmThreads - 1) / kNrmThreads, kNrmThreads, 0, s>>>(pts, n, K, normals);
  return cudaGetLastError();
}

// Bare fp64 issue-rate probe: 8 independent DFMA chains per thread, registers only.  The
// triangulation kernel is bound by this pipe, not by HBM.
__global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, double seed, double* sink) {
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = seed + k + threadIdx.x;
  const double m = 1.0 + seed * 1e-9, c = seed * 1e-7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
  }
  double t = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) t += a[k];
  if (t == 12345.678) sink[0] = t;      // never true: keeps the chains alive
}

cudaError_t launch_fp64_peak(int iters, int n_sms, double* sink, cudaStream_t s) {
  fp64_peak_kernel<<<n_sms * 8, 256, 0, s>>>(iters, 1.0, sink);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------- launchers
int geometry_grid(int64_t n, int n_sms) {
  const int64_t blocks = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(n_sms) * 8;   // 8 resident 256-thread CTAs per SM
  return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// get_matched_points (NViewReconstuct.cpp:989-1003) on the device: match m of a pair picks
// kp_q[queryIdx].pt and kp_t[trainIdx].pt; sel (nullable) lists the matches that survive
// maskout_points (:943).  Output is view-major [2][n][2], the triangulation kernel's layout.
__global__ void gather_matched_points_kernel(const sfm_match_t* __restrict__ matches,
                                             const int32_t* __restrict__ sel, int64_t n,
                                             const float2* __restrict__ kp_q,
                                             const float2* __restrict__ kp_t,
                                             float2* __restrict__ xy) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 m = *reinterpret_cast<const int4*>(&matches[sel ? sel[i] : i]);
  xy[i] = kp_q[m.x];
  xy[n + i] = kp_t[m.y];
}

cudaError_t launch_gather_matched_points(const sfm_match_t* matches, const int32_t* sel, int64_t n,
                                         const float* kp_q, const float* kp_t, float* xy,
                                         cudaStream_t s) {
  if (n > 0)
    gather_matched_points_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(
        matches, sel, n, reinterpret_cast<const float2*>(kp_q),
        reinterpret_cast<const float2*>(kp_t), reinterpret_cast<float2*>(xy));
  return cudaGetLastError();
}

cudaError_t launch_triangulate(const float* P, const float* P_host, const float* xy, int n_views,
                               int64_t n_pts, float* X4, double* xyz, int n_sms, cudaStream_t s) {
  ProjParam pp;
  if (P_host != nullptr && n_views <= kMaxViewsParam) {
    for (int i = 0; i < n_views * 12; ++i) pp.p[i] = static_cast<double>(P_host[i]);
    // view count as a template argument: the loop over the views unrolls and every projection
    // entry becomes a constant-bank operand of its DFMA (the generic loop fetches them with LDC)
    const int grid = geometry_grid(n_pts, n_sms);
    switch (n_views) {
#define SFM_TRI_CASE(V) \
      case V: triangulate_kernel<true, V><<<grid, 256, 0, s>>>(pp, P, xy, n_views, n_pts, X4, xyz); break;
      SFM_TRI_CASE(2) SFM_TRI_CASE(3) SFM_TRI_CASE(4) SFM_TRI_CASE(5) SFM_TRI_CASE(6) SFM_TRI_CASE(7) SFM_TRI_CASE(8)
#undef SFM_TRI_CASE
      default: triangulate_kernel<true, 0><<<grid, 256, 0, s>>>(pp, P, xy, n_views, n_pts, X4, xyz);
    }
  } else {
    triangulate_kernel<false, 0><<<geometry_grid(n_pts, n_sms), 256, 0, s>>>(pp, P, xy, n_views, n_pts, X4, xyz);
  }
  return cudaGetLastError();
}

// Range check of the observation tables on the device (they are uploaded anyway): *flag != 0
// when an observation names a camera or point that does not exist.
__global__ void __launch_bounds__(256)
validate_indices_kernel(const int32_t* __restrict__ cam_idx, const int32_t* __restrict__ pt_idx,
                        int64_t n_obs, int n_cam, int64_t n_pts, uint32_t* __restrict__ flag) {
  bool bad = false;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < n_obs; k += stride) {
    const int c = __ldg(cam_idx + k), j = __ldg(pt_idx + k);
    bad |= (c < 0) | (c >= n_cam) | (j < 0) | (j >= n_pts);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

cudaError_t launch_validate_indices(const int32_t* cam_idx, const int32_t* pt_idx, int64_t n_obs,
                                    int n_cam, int64_t n_pts, uint32_t* flag, int n_sms, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(flag, 0, 4, s);
  if (e != cudaSuccess) return e;
  if (n_obs > 0)
    validate_indices_kernel<<<geometry_grid(n_obs, n_sms), 256, 0, s>>>(cam_idx, pt_idx, n_obs, n_cam, n_pts, flag);
  return cudaGetLastError();
}

cudaError_t launch_camera_table(const double* ext, int n_cam, double* cam, cudaStream_t s) {
  camera_table_kernel<<<(n_cam + 127) / 128, 128, 0, s>>>(ext, n_cam, cam);
  return cudaGetLastError();
}

cudaError_t launch_jacobians(const double intr[4], const double* ext, int n_cam, double* cam,
                             double* jtab, const double* pts, const int32_t* cam_idx,
                             const int32_t* pt_idx, const float* obs_xy, int64_t n_obs,
                             double* resid, double* jac, int n_sms, bool tables, cudaStream_t s) {
  if (tables) {
    camera_table_kernel<<<(n_cam + 127) / 128, 128, 0, s>>>(ext, n_cam, cam);
    camera_jac_table_kernel<<<(n_cam + 127) / 128, 128, 0, s>>>(ext, n_cam, jtab);
  }
  const int64_t blocks = (n_obs + kJacThreads - 1) / kJacThreads;
  const int grid = static_cast<int>(blocks < 16ll * n_sms ? (blocks > 0 ? blocks : 1) : 16ll * n_sms);
  jacobian_kernel<<<grid, kJacThreads, 0, s>>>(intr[0], intr[1], intr[2], intr[3], cam, jtab, pts,
                                               cam_idx, pt_idx, obs_xy, n_obs, resid, jac);
  return cudaGetLastError();
}

// Camera-major probe: flag (device) != 0 afterwards when cam_idx is not sorted; otherwise seg holds
// the n_cam + 1 segment offsets.
cudaError_t launch_order_probe(const int32_t* cam_idx, int64_t n_obs, int n_cam, uint32_t* flag,
                               int64_t* seg, int n_sms, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(flag, 0, 4, s);
  if (e != cudaSuccess) return e;
  if (n_obs > 1) sorted_check_kernel<<<geometry_grid(n_obs, n_sms), 256, 0, s>>>(cam_idx, n_obs, flag);
  segment_table_kernel<<<(n_cam + 1 + 127) / 128, 128, 0, s>>>(cam_idx, n_obs, n_cam, seg);
  return cudaGetLastError();
}

// seg != nullptr (camera-major list, see residual_seg_kernel): max_seg = longest camera segment
cudaError_t launch_residuals(const double intr[4], const double* cam, const double* pts,
                             const int32_t* cam_idx, const int32_t* pt_idx, const float* obs_xy,
                             int64_t n_obs, double huber_delta, double* resid, double* block_cost,
                             double* cost_out, int grid, cudaStream_t s, const int64_t* seg = nullptr,
                             int n_cam = 0, int64_t max_seg = 0) {
  if (seg != nullptr && n_cam > 1)
    residual_seg_kernel<<<grid, 256, 0, s>>>(intr[0], intr[1], intr[2], intr[3], cam, pts, seg, n_cam,
                                             max_seg, pt_idx, obs_xy, huber_delta, resid, block_cost);
  else
    residual_kernel<<<grid, 256, 0, s>>>(intr[0], intr[1], intr[2], intr[3], cam, pts, cam_idx,
                                         pt_idx, obs_xy, n_obs, huber_delta, resid, block_cost);
  if (block_cost != nullptr) cost_sum_kernel<<<1, 256, 0, s>>>(block_cost, grid, cost_out);
  return cudaGetLastError();
}

}  // namespace sfm
