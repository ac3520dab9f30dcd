// Kernels for the step in front of bundle adjustment:
//   * undistort_kernel     == CameraData.undistort_points  (reference cameras/camera_array.py:135-174, which calls
//                             cv2.undistortPoints / cv2.fisheye.undistortPoints on float32 copies of the points)
//   * tri_* kernels        == triangulate_image_points     (reference core/point_data.py:122-229): group observations
//                             by a 64-bit composite key, DLT system per group, smallest right singular vector.
// Algorithmic traffic is 36 B (undistort) / 24 B (DLT, gathered) per observation, but on B200 both are bound by the fp64
// pipe (iterative inverse distortion; 4x4 eigen-solve), not by HBM: profiles/r01/ncu_summary_triangulate.jsonl.
#pragma once
#include <cstdint>

#include "cb_device.cuh"

namespace cb {

struct UndistCam {
  double fx, fy, cx, cy, skew;
  double d[12];  // pinhole: k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 ; fisheye: k1..k4
  int fisheye;
  int pad;
};

__device__ __forceinline__ float2 load_px(const double* p, long long i) {
  const double2 v = reinterpret_cast<const double2*>(p)[i];
  return make_float2((float)v.x, (float)v.y);
}
__device__ __forceinline__ float2 load_px(const float* p, long long i) { return reinterpret_cast<const float2*>(p)[i]; }

// float32 in -> double arithmetic -> float32 out, exactly the precision contract of the reference call.
// TIN = double (caller's array, rounded to float32 here) or float (already rounded on the host while staging).
template <typename TIN>
__global__ void undistort_kernel(const UndistCam* __restrict__ cams, const int* __restrict__ obs_cam,
                                 const TIN* __restrict__ xy_in, double* __restrict__ xy_out, long long n,
                                 int to_pixels) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const UndistCam& c = cams[obs_cam ? obs_cam[i] : 0];
  const float2 in = load_px(xy_in, i);
  const double u = (double)in.x, v = (double)in.y;
  double x, y;
  if (c.fisheye) {
    const double pwx = (u - c.cx) / c.fx, pwy = (v - c.cy) / c.fy;
    double theta_d = sqrt(pwx * pwx + pwy * pwy);
    const double half_pi = 1.5707963267948966;
    theta_d = fmin(fmax(-half_pi, theta_d), half_pi);
    bool converged = false;
    double th = theta_d, scale = 0.0;
    if (theta_d > 1e-8) {
      for (int j = 0; j < 10; ++j) {
        const double t2 = th * th, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
        const double k0 = c.d[0] * t2, k1 = c.d[1] * t4, k2 = c.d[2] * t6, k3 = c.d[3] * t8;
        const double fix = (th * (1 + k0 + k1 + k2 + k3) - theta_d) / (1 + 3 * k0 + 5 * k1 + 7 * k2 + 9 * k3);
        th -= fix;
        if (fabs(fix) < 1e-8) {
          converged = true;
          break;
        }
      }
      scale = tan(th) / theta_d;
    } else {
      converged = true;
    }
    const bool flipped = (theta_d < 0 && th > 0) || (theta_d > 0 && th < 0);
    if (converged && !flipped) {
      x = pwx * scale;
      y = pwy * scale;
    } else {
      x = -1000000.0;
      y = -1000000.0;
    }
  } else {
    const double x0 = (u - c.cx) / c.fx, y0 = (v - c.cy) / c.fy;
    x = x0;
    y = y0;
#pragma unroll 1
    for (int j = 0; j < 5; ++j) {
      const double r2 = x * x + y * y;
      const double icdist =
          (1 + ((c.d[7] * r2 + c.d[6]) * r2 + c.d[5]) * r2) / (1 + ((c.d[4] * r2 + c.d[1]) * r2 + c.d[0]) * r2);
      if (icdist < 0) {
        x = x0;
        y = y0;
        break;
      }
      const double dx = 2 * c.d[2] * x * y + c.d[3] * (r2 + 2 * x * x) + c.d[8] * r2 + c.d[9] * r2 * r2;
      const double dy = c.d[2] * (r2 + 2 * y * y) + 2 * c.d[3] * x * y + c.d[10] * r2 + c.d[11] * r2 * r2;
      x = (x0 - dx) * icdist;
      y = (y0 - dy) * icdist;
    }
  }
  if (to_pixels) {
    const double px = c.fx * x + c.skew * y + c.cx, py = c.fy * y + c.cy;
    x = px;
    y = py;
  }
  reinterpret_cast<double2*>(xy_out)[i] = make_double2((double)(float)x, (double)(float)y);
}

// ---- grouping --------------------------------------------------------------------------------------------------
__global__ void tri_iota_kernel(int* __restrict__ v, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) v[i] = (int)i;
}

// counts rows whose camera index is outside [0, n_cams) or whose key is negative
__global__ void tri_validate_kernel(const int* __restrict__ cam, const long long* __restrict__ key, long long n,
                                    int n_cams, int* __restrict__ bad) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool b = cam[i] < 0 || cam[i] >= n_cams || (key && key[i] < 0);
  if (b) atomicAdd(bad, 1);
}

__global__ void tri_heads_kernel(const unsigned long long* __restrict__ k, long long n, int* __restrict__ head) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || k[i] != k[i - 1]) ? 1 : 0;
}

// gid = inclusive scan of head; start[gid-1] = i at heads; start[n_groups] = n
__global__ void tri_starts_kernel(const int* __restrict__ head, const int* __restrict__ gid, long long n,
                                  int* __restrict__ start) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (head[i]) start[gid[i] - 1] = (int)i;
  if (i == n - 1) start[gid[i]] = (int)n;
}

// ---- DLT ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long tri_mix(unsigned long long x, unsigned long long salt) {
  x += salt;
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27; x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}

// Eigenvector of the smallest eigenvalue of a symmetric 4x4 (cyclic Jacobi; quadratic convergence, <= 12 sweeps).
__device__ __forceinline__ void sym4_min_eigvec(double a[4][4], double out[4]) {
  double V[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
#pragma unroll 1
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0.0, dg = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      dg += a[i][i] * a[i][i];
#pragma unroll
      for (int j = i + 1; j < 4; ++j) off += a[i][j] * a[i][j];
    }
    if (off <= 1e-34 * dg) break;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        const double apq = a[p][q];
        if (apq != 0.0) {
          const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
          const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          const double c = rsqrt(t * t + 1.0), s = t * c;
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // A <- A J
            const double akp = a[k][p], akq = a[k][q];
            a[k][p] = c * akp - s * akq;
            a[k][q] = s * akp + c * akq;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // A <- J^T A
            const double apk = a[p][k], aqk = a[q][k];
            a[p][k] = c * apk - s * aqk;
            a[q][k] = s * apk + c * aqk;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const double vkp = V[k][p], vkq = V[k][q];
            V[k][p] = c * vkp - s * vkq;
            V[k][q] = s * vkp + c * vkq;
          }
        }
      }
  }
  int m = 0;
  double best = a[0][0];
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (a[i][i] < best) {
      best = a[i][i];
      m = i;
    }
#pragma unroll
  for (int k = 0; k < 4; ++k) out[k] = (m == 0) ? V[k][0] : (m == 1) ? V[k][1] : (m == 2) ? V[k][2] : V[k][3];
}

constexpr int TRI_THREADS = 256;

// One group of observations (same (sync, object, keypoint) key) per TRI_LANES lanes (8 for the 2-6 views of a
// charuco corner, 32 when groups average more than 16 rows).  Each lane accumulates the normal matrix
// M = sum_rows (x P2 - P0)(x P2 - P0)^T + (y P2 - P1)(y P2 - P1)^T  of the DLT system over its rows, an
// xor-butterfly leaves the identical sum on all lanes, the smallest eigenvector is de-homogenised.
// `proj` is the [n_cams][3][4] table (shared-memory copy when it fits).
template <int TRI_LANES>
__global__ void __launch_bounds__(TRI_THREADS)
tri_dlt_kernel(const double* __restrict__ proj, int n_cams, int proj_in_smem, const int* __restrict__ start,
               const int* __restrict__ rows, const int* __restrict__ obs_cam, const double* __restrict__ obs_xy,
               int n_groups, double* __restrict__ xyz, int* __restrict__ count, int* __restrict__ rep_row,
               unsigned long long* __restrict__ sig) {
  extern __shared__ double s_proj[];
  // shared-memory copy with a stride of 13 doubles per camera: lanes read DIFFERENT cameras' rows, and a stride of 12 puts
  // the same entry of consecutive cameras into 4 of the 16 eight-byte banks (round 1: 4.1 M bank conflicts per launch)
  constexpr int PSTRIDE_SM = 13;
  if (proj_in_smem) {
    for (int i = threadIdx.x; i < n_cams * 12; i += blockDim.x) s_proj[(i / 12) * PSTRIDE_SM + i % 12] = proj[i];
    __syncthreads();
  }
  const double* P = proj_in_smem ? s_proj : proj;
  const int pstride = proj_in_smem ? PSTRIDE_SM : 12;
  const int lane = threadIdx.x & (TRI_LANES - 1);
  const long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / TRI_LANES;
  const bool live = g < n_groups;
  const int b = live ? start[g] : 0, e = live ? start[g + 1] : 0;
  double m[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) m[k] = 0.0;
  unsigned long long h1 = 0, h2 = 0;
  for (int i = b + lane; i < e; i += TRI_LANES) {
    const int r = rows[i];
    const int c = obs_cam[r];
    const double2 xy = reinterpret_cast<const double2*>(obs_xy)[r];
    const double* Pc = P + (size_t)pstride * (size_t)c;
    double a[4], bb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double p2 = Pc[8 + k];
      a[k] = xy.x * p2 - Pc[k];
      bb[k] = xy.y * p2 - Pc[4 + k];
    }
    int t = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = p; q < 4; ++q) {
        m[t] = fma(a[p], a[q], fma(bb[p], bb[q], m[t]));
        ++t;
      }
    h1 += tri_mix((unsigned long long)(unsigned)c, 0x9e3779b97f4a7c15ULL);
    h2 += tri_mix((unsigned long long)(unsigned)c, 0xd1b54a32d192ed03ULL);
  }
#pragma unroll
  for (int s = TRI_LANES / 2; s > 0; s >>= 1) {
#pragma unroll
    for (int k = 0; k < 10; ++k) m[k] += __shfl_xor_sync(0xffffffffu, m[k], s);
    h1 += __shfl_xor_sync(0xffffffffu, h1, s);
    h2 += __shfl_xor_sync(0xffffffffu, h2, s);
  }
  if (!live || lane != 0) return;
  const int n = e - b;
  count[g] = n;
  rep_row[g] = rows[b];
  sig[2 * g] = h1;
  sig[2 * g + 1] = h2;
  if (n < 2) {
    xyz[3 * g] = xyz[3 * g + 1] = xyz[3 * g + 2] = __longlong_as_double(0x7ff8000000000000LL);
    return;
  }
  double A[4][4];
  int t = 0;
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = p; q < 4; ++q) {
      A[p][q] = m[t];
      A[q][p] = m[t];
      ++t;
    }
  double w[4];
  sym4_min_eigvec(A, w);
  xyz[3 * g + 0] = w[0] / w[3];
  xyz[3 * g + 1] = w[1] / w[3];
  xyz[3 * g + 2] = w[2] / w[3];
}

}  // namespace cb
