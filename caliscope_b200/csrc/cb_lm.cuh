// Device-resident Levenberg-Marquardt state machine and the Jacobian-free point kernels.
//
// Round-2 structure of one LM trial (every kernel reads the damping, the current-buffer index and the
// stop flag from LmState in device memory, so the host never has to look at a number between launches and
// the whole trial replays as one CUDA graph):
//
//   pt_pass_kernel        per point: recompute r, Jp, Jc of its observations from (pixel, camera table, point),
//                         V = sum Jp^T Jp, g = sum Jp^T r, Marquardt scale, 3x3 damped Cholesky, t = L^-1 g and
//                         Z = (Jc^T Jp) L^-T streamed straight into the k-major Schur factor   [reprojection.py:171-205]
//   schur_syrk_kernel     Z Z^T (+ Z t), split-K                                               (cb_kernels.cuh)
//   schur_finalize(_peer) S = U - Z Z^T, b = g_c - Z t (+ the all-reduce over NVLink peers)
//   reduced_prep_kernel   damping, gradient norm, gtol / max_nfev tests, block-Jacobi inverses
//   pcg_cluster_kernel    reduced camera solve
//   cam_step_kernel       camera step, bounds clamp, predicted reduction, camera table of the trial point
//   pt_backsub_kernel     dX = -L^-T (t + L sum Jp^T (Jc dc)), again recomputed from the observation list
//   resjac_kernel<P,5>    camera-major pass at the TRIAL point: cost, U_c, g_c (these are the next linearisation's
//                         camera blocks if the step is accepted)
//   trial_reduce_kernel   chunk partials -> per-camera blocks, trial sums, and (single GPU) the accept/reject decision
//   lm_decide_kernel      (multi GPU) 4-double all-reduce over peer memory + the decision
//
// The decision follows scipy's TRF bookkeeping (site-packages/scipy/optimize/_lsq/trf.py:465-560,
// common.py:705-717): nfev / njev / nit count the same events, termination statuses 0..4 are scipy's.
#pragma once
#include "cb_kernels.cuh"

namespace cb {

constexpr int LM_ERR_NONFINITE_X0 = 1;   // scipy: "Residuals are not finite in the initial point."
constexpr int LM_ERR_STUCK_NONFINITE = 2;  // damping saturated and every trial non-finite: stop instead of spinning

struct LmLogRow {
  double nit, nfev, cost, cost_new, ratio, lam, step, gnorm, pcg;
};

// sum over the LANES lanes of a sub-warp group (groups are aligned, so xor offsets below LANES stay inside one)
template <int LANES>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// Point pass.  LANES lanes per point (32 / LANES points per warp), persistent grid-stride over points so the camera
// table is staged into shared memory once per CTA.  DUPS: repeated (camera, point) rows exist (static objects seen
// in many frames); they are adjacent in the point-major list and the first lane of a run sums the run.
// Points tied by rigid-distance constraints (pt_comp >= 0) only get V, g and the untransformed W = Jc^T Jp written to
// their Zt rows; comp_build_kernel eliminates the component.
// ---------------------------------------------------------------------------------------------
// CAMSM: the camera table is staged in shared memory (it fits: <= 64 KB).  A template parameter, not a run-time pointer
// select: with a pointer that may be shared or global the compiler emits GENERIC loads (LD.E) for the ~36 table reads per
// observation, which wait on the long scoreboard like global loads (ncu: 55 % of the stalls of the first version).
template <int P, int LANES, bool DUPS, bool CAMSM>
__global__ void __launch_bounds__(PT_WARPS * 32, 2)
pt_pass_kernel(const LmState* __restrict__ st, const int* __restrict__ pt_start, const int* __restrict__ pm_cam,
               const double2* __restrict__ pm_xy, const int* __restrict__ pt_comp, int n_pts, int n_cams,
               CPtr2 camtab2, CPtr2 xp2, double* __restrict__ V6, double* __restrict__ gp,
               double* __restrict__ Dp2, double* __restrict__ Linv6, double* __restrict__ tvec,
               double* __restrict__ Zt, size_t LD, unsigned long long* __restrict__ gmax_bits) {
  extern __shared__ __align__(16) double pt_sm[];
  __shared__ double wmax[PT_WARPS];
  if (st->done) return;
  const int cur = st->cur;
  const double lam = st->lam;
  const int loss = st->loss;
  const double fscale = st->fscale;
  const double* __restrict__ gtab = camtab2.p[cur];
  const double* xp4 = xp2.p[cur];
  constexpr int cstride = CAMSM ? CT_SMEM : CT_SIZE;
  if constexpr (CAMSM) {
    for (int i = threadIdx.x; i < n_cams * CT_SIZE; i += blockDim.x) pt_sm[(i / CT_SIZE) * CT_SMEM + i % CT_SIZE] = gtab[i];
    __syncthreads();
  }
  // camera table entry of camera c: shared (LDS) or global (LDG), decided at compile time
  auto cam_entry = [&](int c) -> const double* {
    if constexpr (CAMSM) return pt_sm + (size_t)c * cstride;
    else return gtab + (size_t)c * cstride;
  };
  constexpr int GPW = 32 / LANES;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int gl = lane % LANES, grp = lane / LANES;
  double gm = 0.0;
  for (int j0 = (blockIdx.x * PT_WARPS + wid) * GPW; j0 < n_pts; j0 += gridDim.x * PT_WARPS * GPW) {
    const int j = j0 + grp;
    const bool valid = j < n_pts;
    int s = 0, e = 0;
    double X0 = 0.0, X1 = 0.0, X2 = 0.0, X3;
    if (valid) {
      s = pt_start[j]; e = pt_start[j + 1];
      ld256nc(xp4 + 4 * (size_t)j, X0, X1, X2, X3);
    }
    (void)X3;
    const bool in_comp = valid && pt_comp != nullptr && pt_comp[j] >= 0;
    // ---- phase 1: V, g over the point's observations (residual and d f / d X only)
    double v[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) v[k] = 0.0;
    // (index / pixel of the next row are requested before the current one is evaluated: the kernel is latency bound --
    // 16 warps per SM, long_scoreboard 60 % of the stalls in the first version)
    int cam_n = 0;
    double2 xy_n = make_double2(0.0, 0.0);
    if (s + gl < e) { cam_n = pm_cam[s + gl]; xy_n = pm_xy[s + gl]; }
    for (int pos = s + gl; pos < e; pos += LANES) {
      const int cam = cam_n;
      const double2 xy = xy_n;
      if (pos + LANES < e) { cam_n = pm_cam[pos + LANES]; xy_n = pm_xy[pos + LANES]; }
      double f[2], JX[6];
      obs_res_jx(cam_entry(cam), X0, X1, X2, xy.x, xy.y, loss, fscale, f, JX);
      v[0] += JX[0] * JX[0] + JX[3] * JX[3]; v[1] += JX[0] * JX[1] + JX[3] * JX[4]; v[2] += JX[0] * JX[2] + JX[3] * JX[5];
      v[3] += JX[1] * JX[1] + JX[4] * JX[4]; v[4] += JX[1] * JX[2] + JX[4] * JX[5]; v[5] += JX[2] * JX[2] + JX[5] * JX[5];
      v[6] += JX[0] * f[0] + JX[3] * f[1]; v[7] += JX[1] * f[0] + JX[4] * f[1]; v[8] += JX[2] * f[0] + JX[5] * f[1];
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) v[k] = group_sum<LANES>(v[k]);
    double D[3] = {1.0, 1.0, 1.0};
    if (valid) {
      const double* d = Dp2 + (size_t)j * 3;
      D[0] = fmax(d[0], v[0]); D[1] = fmax(d[1], v[3]); D[2] = fmax(d[2], v[5]);
    }
    double Li[6];
    if (in_comp) { Li[0] = 1.0; Li[1] = 0.0; Li[2] = 1.0; Li[3] = 0.0; Li[4] = 0.0; Li[5] = 1.0; }
    else chol3_inv(v, D, lam, Li);
    if (valid && gl == 0) {
#pragma unroll
      for (int k = 0; k < 6; ++k) V6[(size_t)j * 6 + k] = v[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) { gp[(size_t)j * 3 + k] = v[6 + k]; Dp2[(size_t)j * 3 + k] = D[k]; }
      if (!in_comp) {
#pragma unroll
        for (int k = 0; k < 6; ++k) Linv6[(size_t)j * 6 + k] = Li[k];
        tvec[3 * (size_t)j + 0] = Li[0] * v[6];
        tvec[3 * (size_t)j + 1] = Li[1] * v[6] + Li[2] * v[7];
        tvec[3 * (size_t)j + 2] = Li[3] * v[6] + Li[4] * v[7] + Li[5] * v[8];
        gm = fmax(gm, fmax(fabs(v[6]), fmax(fabs(v[7]), fabs(v[8]))));
      }
    }
    // ---- phase 2: Z = (Jc^T Jp) Linv^T per (camera, point) pair, rows 3j..3j+2 of the k-major factor
    if (s + gl < e) { cam_n = pm_cam[s + gl]; xy_n = pm_xy[s + gl]; }
    for (int pos = s + gl; pos < e; pos += LANES) {
      int cam;
      double f[2], JX[6], Jc[2 * P];
      double z[3][P];
      if constexpr (!DUPS) {
        cam = cam_n;
        const double2 xy = xy_n;
        if (pos + LANES < e) { cam_n = pm_cam[pos + LANES]; xy_n = pm_xy[pos + LANES]; }
        obs_jac<P>(cam_entry(cam), X0, X1, X2, xy.x, xy.y, loss, fscale, f, JX, Jc);
        const double q00 = JX[0] * Li[0], q01 = JX[0] * Li[1] + JX[1] * Li[2], q02 = JX[0] * Li[3] + JX[1] * Li[4] + JX[2] * Li[5];
        const double q10 = JX[3] * Li[0], q11 = JX[3] * Li[1] + JX[4] * Li[2], q12 = JX[3] * Li[3] + JX[4] * Li[4] + JX[5] * Li[5];
#pragma unroll
        for (int p = 0; p < P; ++p) {
          z[0][p] = fma(Jc[p], q00, Jc[P + p] * q10);
          z[1][p] = fma(Jc[p], q01, Jc[P + p] * q11);
          z[2][p] = fma(Jc[p], q02, Jc[P + p] * q12);
        }
      } else {
        cam = pm_cam[pos];
        if (pos > s && pm_cam[pos - 1] == cam) continue;  // not the first row of its (point, camera) run
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int p = 0; p < P; ++p) z[a][p] = 0.0;
        int r = pos;
        do {
          const double2 xy = pm_xy[r];
          obs_jac<P>(cam_entry(cam), X0, X1, X2, xy.x, xy.y, loss, fscale, f, JX, Jc);
          const double q00 = JX[0] * Li[0], q01 = JX[0] * Li[1] + JX[1] * Li[2], q02 = JX[0] * Li[3] + JX[1] * Li[4] + JX[2] * Li[5];
          const double q10 = JX[3] * Li[0], q11 = JX[3] * Li[1] + JX[4] * Li[2], q12 = JX[3] * Li[3] + JX[4] * Li[4] + JX[5] * Li[5];
#pragma unroll
          for (int p = 0; p < P; ++p) {
            z[0][p] = fma(Jc[p], q00, fma(Jc[P + p], q10, z[0][p]));
            z[1][p] = fma(Jc[p], q01, fma(Jc[P + p], q11, z[1][p]));
            z[2][p] = fma(Jc[p], q02, fma(Jc[P + p], q12, z[2][p]));
          }
          ++r;
        } while (r < e && pm_cam[r] == cam);
      }
      (void)f;
      double* z0 = Zt + (3 * (size_t)j) * LD + (size_t)cam * P;
      if constexpr (P == 6) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          double2* dst = reinterpret_cast<double2*>(z0 + a * LD);  // cam*48 B and LD*8 B are 16-byte multiples
          dst[0] = make_double2(z[a][0], z[a][1]);
          dst[1] = make_double2(z[a][2], z[a][3]);
          dst[2] = make_double2(z[a][4], z[a][5]);
        }
      } else {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int p = 0; p < P; ++p) z0[a * LD + p] = z[a][p];
      }
    }
  }
  gm = warp_max(gm);
  if (lane == 0) wmax[wid] = gm;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int w = 0; w < PT_WARPS; ++w) m = fmax(m, wmax[w]);
    if (m > 0.0) atomicMax(gmax_bits, (unsigned long long)__double_as_longlong(m));
  }
}

// ---------------------------------------------------------------------------------------------
// Point back-substitution, Jacobian-free:  Z_j^T dc = L^-1 sum_obs Jp^T (Jc dc_cam), so
//   dX_j = -L^-T (t_j + L^-1 u),  u = sum_obs Jp^T (Jc dc)
// recomputed from the observation list (24 B / observation) instead of streaming the dense factor.
// Per-CTA partial sums of the predicted reduction / step / x norms (fixed grid => deterministic).
// ---------------------------------------------------------------------------------------------
template <int P, int LANES, bool CAMSM>
__global__ void __launch_bounds__(PT_WARPS * 32, 2)
pt_backsub_kernel(const LmState* __restrict__ st, const int* __restrict__ pt_start, const int* __restrict__ pm_cam,
                  const double2* __restrict__ pm_xy, const int* __restrict__ pt_comp, int n_pts, int n_cams, int nP,
                  CPtr2 camtab2, Ptr2 xp2, const double* __restrict__ dc,
                  const double* __restrict__ Linv6, const double* __restrict__ tvec, const double* __restrict__ gp,
                  const double* __restrict__ Dp2, double* __restrict__ dp_out, double* __restrict__ bpart,
                  int bpart_stride) {
  extern __shared__ __align__(16) double bs_sm[];
  __shared__ double wsum[3][PT_WARPS];
  if (st->done) return;
  const int cur = st->cur;
  const double lam = st->lam;
  const int loss = st->loss;
  const double fscale = st->fscale;
  double* dcs = bs_sm;
  const double* __restrict__ gtab = camtab2.p[cur];
  const double* xp4 = xp2.p[cur];
  double* xp4_new = xp2.p[cur ^ 1];
  for (int i = threadIdx.x; i < nP; i += blockDim.x) dcs[i] = dc[i];
  constexpr int cstride = CAMSM ? CT_SMEM : CT_SIZE;
  double* cs = bs_sm + ((nP + 3) & ~3);
  if constexpr (CAMSM)
    for (int i = threadIdx.x; i < n_cams * CT_SIZE; i += blockDim.x) cs[(i / CT_SIZE) * CT_SMEM + i % CT_SIZE] = gtab[i];
  __syncthreads();
  auto cam_entry = [&](int c) -> const double* {
    if constexpr (CAMSM) return cs + (size_t)c * cstride;
    else return gtab + (size_t)c * cstride;
  };
  constexpr int GPW = 32 / LANES;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int gl = lane % LANES, grp = lane / LANES;
  double pred = 0.0, st2 = 0.0, x2 = 0.0;
  for (int j0 = (blockIdx.x * PT_WARPS + wid) * GPW; j0 < n_pts; j0 += gridDim.x * PT_WARPS * GPW) {
    const int j = j0 + grp;
    const bool valid = j < n_pts && !(pt_comp != nullptr && pt_comp[j] >= 0);
    int s = 0, e = 0;
    double X0 = 0.0, X1 = 0.0, X2 = 0.0, X3;
    if (valid) {
      s = pt_start[j]; e = pt_start[j + 1];
      ld256nc(xp4 + 4 * (size_t)j, X0, X1, X2, X3);
    }
    (void)X3;
    double u0 = 0.0, u1 = 0.0, u2 = 0.0;
    int cam_n = 0;
    double2 xy_n = make_double2(0.0, 0.0);
    if (s + gl < e) { cam_n = pm_cam[s + gl]; xy_n = pm_xy[s + gl]; }
    for (int pos = s + gl; pos < e; pos += LANES) {
      const int cam = cam_n;
      const double2 xy = xy_n;
      if (pos + LANES < e) { cam_n = pm_cam[pos + LANES]; xy_n = pm_xy[pos + LANES]; }
      double f[2], JX[6], Jc[2 * P];
      obs_jac<P>(cam_entry(cam), X0, X1, X2, xy.x, xy.y, loss, fscale, f, JX, Jc);
      const double* d = dcs + cam * P;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int p = 0; p < P; ++p) { s0 = fma(Jc[p], d[p], s0); s1 = fma(Jc[P + p], d[p], s1); }
      u0 = fma(JX[0], s0, fma(JX[3], s1, u0));
      u1 = fma(JX[1], s0, fma(JX[4], s1, u1));
      u2 = fma(JX[2], s0, fma(JX[5], s1, u2));
    }
    u0 = group_sum<LANES>(u0); u1 = group_sum<LANES>(u1); u2 = group_sum<LANES>(u2);
    if (valid && gl == 0) {
      const double* Li = Linv6 + (size_t)j * 6;
      const double w0 = Li[0] * u0, w1 = Li[1] * u0 + Li[2] * u1, w2 = Li[3] * u0 + Li[4] * u1 + Li[5] * u2;
      const double v0 = tvec[3 * (size_t)j] + w0, v1 = tvec[3 * (size_t)j + 1] + w1, v2 = tvec[3 * (size_t)j + 2] + w2;
      const double d0 = -(Li[0] * v0 + Li[1] * v1 + Li[3] * v2);
      const double d1 = -(Li[2] * v1 + Li[4] * v2);
      const double d2 = -(Li[5] * v2);
      double* xn = xp4_new + 4 * (size_t)j;
      xn[0] = X0 + d0; xn[1] = X1 + d1; xn[2] = X2 + d2; xn[3] = 0.0;
      if (dp_out) { dp_out[3 * (size_t)j] = d0; dp_out[3 * (size_t)j + 1] = d1; dp_out[3 * (size_t)j + 2] = d2; }
      const double* D = Dp2 + 3 * (size_t)j;
      const double* g = gp + 3 * (size_t)j;
      const double e0 = D[0] > 0.0 ? D[0] : 1.0, e1 = D[1] > 0.0 ? D[1] : 1.0, e2 = D[2] > 0.0 ? D[2] : 1.0;
      pred += 0.5 * (d0 * (lam * e0 * d0 - g[0]) + d1 * (lam * e1 * d1 - g[1]) + d2 * (lam * e2 * d2 - g[2]));
      st2 += d0 * d0 + d1 * d1 + d2 * d2;
      x2 += X0 * X0 + X1 * X1 + X2 * X2;
    }
  }
  pred = warp_sum(pred); st2 = warp_sum(st2); x2 = warp_sum(x2);
  if (lane == 0) { wsum[0][wid] = pred; wsum[1][wid] = st2; wsum[2][wid] = x2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 0;
    for (int w = 0; w < PT_WARPS; ++w) { a += wsum[0][w]; b += wsum[1][w]; c += wsum[2][w]; }
    bpart[blockIdx.x] = a;
    bpart[bpart_stride + blockIdx.x] = b;
    bpart[2 * (size_t)bpart_stride + blockIdx.x] = c;
  }
}

// ---------------------------------------------------------------------------------------------
// After the (all-)reduce of [S | b | g_c | diag U | cost | per-rank point-gradient slots]: Marquardt scaling and
// damping of the reduced system, gradient norm, the start-of-iteration tests, block-Jacobi inverses.  One CTA.
// ---------------------------------------------------------------------------------------------
template <int P>
__device__ __forceinline__ void block_inverse_one(const double* __restrict__ S, int nP, int c, double* __restrict__ out) {
  double A[P][P], Li[P][P];
  for (int i = 0; i < P; ++i)
    for (int j = 0; j < P; ++j) A[i][j] = S[(size_t)(c * P + i) * nP + c * P + j];
  bool ok = true;
  for (int j = 0; j < P; ++j) {
    double d = A[j][j];
    for (int k = 0; k < j; ++k) d -= A[j][k] * A[j][k];
    if (!(d > 0.0)) { ok = false; d = 1.0; }
    d = sqrt(d);
    A[j][j] = d;
    for (int i = j + 1; i < P; ++i) {
      double s = A[i][j];
      for (int k = 0; k < j; ++k) s -= A[i][k] * A[j][k];
      A[i][j] = s / d;
    }
  }
  for (int j = 0; j < P; ++j) {
    Li[j][j] = 1.0 / A[j][j];
    for (int i = j + 1; i < P; ++i) {
      double s = 0.0;
      for (int k = j; k < i; ++k) s -= A[i][k] * Li[k][j];
      Li[i][j] = s / A[i][i];
    }
  }
  for (int i = 0; i < P; ++i)
    for (int j = 0; j < P; ++j) {
      double s = 0.0;
      if (ok) {
        for (int k = (i > j ? i : j); k < P; ++k) s += Li[k][i] * Li[k][j];
      } else {
        const double d = S[(size_t)(c * P + i) * nP + c * P + i];
        s = (i == j) ? (d > 0.0 ? 1.0 / d : 1.0) : 0.0;
      }
      out[i * P + j] = s;
    }
}

template <int P, bool WANT_MINV>
__device__ __forceinline__ void reduced_prep_body(LmState* __restrict__ st, int nP, int n_cams, int red_slots,
                                                  double* __restrict__ red, double* __restrict__ Dc2,
                                                  const unsigned char* __restrict__ active, double* __restrict__ Minv,
                                                  unsigned long long* __restrict__ gmax_bits, double* __restrict__ sc) {
  __shared__ double sh[32];
  const size_t nn = (size_t)nP * nP;
  const double lam = st->lam;
  double gm = 0.0;
  for (int i = threadIdx.x; i < nP; i += blockDim.x) {
    double d = fmax(Dc2[i], red[nn + 2 * (size_t)nP + i]);  // running max; idempotent when the point is unchanged
    Dc2[i] = d;
    red[(size_t)i * nP + i] += lam * (d > 0.0 ? d : 1.0);
    if (active[i]) gm = fmax(gm, fabs(red[nn + nP + i]));
  }
  const size_t slot0 = nn + 3 * (size_t)nP + 1;
  for (int s = threadIdx.x; s < red_slots; s += blockDim.x) gm = fmax(gm, red[slot0 + s]);
  gm = warp_max(gm);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = gm;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, sh[w]);
    sc[SC_GNORM_C] = m;
    const double cost = red[nn + 3 * (size_t)nP];
    sc[SC_COST] = cost;
    *gmax_bits = 0ull;  // consumed by the finalize kernel; the next point pass accumulates afresh
    st->epoch_big += 1;
    int done = 0;
    if (st->new_lin) {
      st->cost = cost;
      st->gnorm = m;
      if (st->njev == 1) {
        st->initial_cost = cost;
        if (!isfinite(cost)) { st->err = LM_ERR_NONFINITE_X0; done = 1; }
      }
      if (!done && m < st->gtol) { st->status = 1; done = 1; }
      if (!done) st->nit += 1;
    }
    if (!done && st->nfev >= st->max_nfev) { st->status = 0; done = 1; }
    if (done) st->done = 1;
  }
  __syncthreads();
  if constexpr (WANT_MINV)
    for (int c = threadIdx.x; c < n_cams; c += blockDim.x) block_inverse_one<P>(red, nP, c, Minv + (size_t)c * P * P);
}
template <int P>
__global__ void __launch_bounds__(256)
reduced_prep_kernel(LmState* __restrict__ st, int nP, int n_cams, int red_slots, double* __restrict__ red,
                    double* __restrict__ Dc2, const unsigned char* __restrict__ active, double* __restrict__ Minv,
                    unsigned long long* __restrict__ gmax_bits, double* __restrict__ sc) {
  if (st->done) return;
  reduced_prep_body<P, true>(st, nP, n_cams, red_slots, red, Dc2, active, Minv, gmax_bits, sc);
}

// ---------------------------------------------------------------------------------------------
// camera step: bounds clamp, effective step back into dc, predicted reduction (camera part), then the camera
// table of the trial point.  One CTA.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cam_step_body(const LmState* __restrict__ st, int nP, int n_cams, int P, Ptr2 xc2,
                                              double* __restrict__ dc, const double* __restrict__ lo,
                                              const double* __restrict__ hi, const double* __restrict__ gc_total,
                                              const double* __restrict__ Dc2, const unsigned char* __restrict__ active,
                                              const int* __restrict__ cam_flags, const double* __restrict__ cam_const,
                                              Ptr2 camtab2, double* __restrict__ sc) {
  __shared__ double sh[3][32];
  const int cur = st->cur;
  const double lam = st->lam;
  const double* xc = xc2.p[cur];
  double* xc_new = xc2.p[cur ^ 1];
  double pred = 0.0, st2 = 0.0, x2 = 0.0;
  for (int i = threadIdx.x; i < nP; i += blockDim.x) {
    const double x = xc[i];
    double xn = x + dc[i];
    xn = fmin(fmax(xn, lo[i]), hi[i]);
    if (!active[i]) xn = x;
    const double de = xn - x;
    dc[i] = de;
    xc_new[i] = xn;
    if (active[i]) {
      const double d = Dc2[i] > 0.0 ? Dc2[i] : 1.0;
      pred += 0.5 * de * (lam * d * de - gc_total[i]);
      st2 += de * de;
      x2 += x * x;
    }
  }
  pred = warp_sum(pred); st2 = warp_sum(st2); x2 = warp_sum(x2);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sh[0][wid] = pred; sh[1][wid] = st2; sh[2][wid] = x2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sh[0][w]; b += sh[1][w]; c += sh[2][w]; }
    sc[SC_PRED_C] = a; sc[SC_STEP2_C] = b; sc[SC_X2_C] = c;
  }
  for (int c = threadIdx.x; c < n_cams; c += blockDim.x)
    cam_prep_one(xc_new + (size_t)c * P, cam_const + (size_t)c * 9, cam_flags[c], camtab2.p[cur ^ 1] + (size_t)c * CT_SIZE);
}
__global__ void __launch_bounds__(256)
cam_step_kernel(const LmState* __restrict__ st, int nP, int n_cams, int P, Ptr2 xc2, double* __restrict__ dc,
                const double* __restrict__ lo, const double* __restrict__ hi, const double* __restrict__ gc_total,
                const double* __restrict__ Dc2, const unsigned char* __restrict__ active,
                const int* __restrict__ cam_flags, const double* __restrict__ cam_const, Ptr2 camtab2,
                double* __restrict__ sc) {
  if (st->done) return;
  cam_step_body(st, nP, n_cams, P, xc2, dc, lo, hi, gc_total, Dc2, active, cam_flags, cam_const, camtab2, sc);
}

// Small rigs (n_camera_params <= DIRECT_MAX_N): damping + head-of-iteration tests, the direct reduced solve and the camera
// step in ONE single-CTA kernel -- three launches and two kernel boundaries (~3.5 us each inside a graph) become one.
template <int P>
__global__ void __launch_bounds__(DIRECT_THREADS, 1)
small_rig_step_kernel(LmState* __restrict__ st, int nP, int n_cams, int red_slots, double* __restrict__ red,
                      double* __restrict__ Dc2, const unsigned char* __restrict__ active,
                      unsigned long long* __restrict__ gmax_bits, double* __restrict__ sc, Ptr2 xc2,
                      double* __restrict__ dc, const double* __restrict__ lo, const double* __restrict__ hi,
                      const int* __restrict__ cam_flags, const double* __restrict__ cam_const, Ptr2 camtab2) {
  extern __shared__ __align__(16) double dsm[];
  if (st->done) return;
  reduced_prep_body<P, false>(st, nP, n_cams, red_slots, red, Dc2, active, nullptr, gmax_bits, sc);
  __syncthreads();
  if (st->done) return;  // set by thread 0 before the barrier: gtol / max_nfev / non-finite start (uniform)
  const size_t nn = (size_t)nP * nP;
  dense_ldlt_body(dsm, red, red + nn, nP, dc, sc);
  __syncthreads();
  cam_step_body(st, nP, n_cams, P, xc2, dc, lo, hi, red + nn + nP, Dc2, active, cam_flags, cam_const, camtab2, sc);
}

// ---------------------------------------------------------------------------------------------
// The accept / reject decision of one trial (scipy trf.py:465-560 bookkeeping, Marquardt/Nielsen damping update).
// red2 = [cost_new, predicted reduction (points), |dX|^2 (points), |X|^2 (points)] summed over ranks.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void lm_decide(LmState* __restrict__ st, const double* __restrict__ sc,
                                          const double* __restrict__ red2, LmLogRow* __restrict__ log) {
  st->nfev += 1;
  st->pcg_total += (long long)sc[SC_PCG_ITS];
  const double cost = st->cost;
  const double cost_new = red2[0];
  const double pred = sc[SC_PRED_C] + red2[1];
  const double step2 = sc[SC_STEP2_C] + red2[2];
  const double x2 = sc[SC_X2_C] + red2[3];
  const bool pcg_bad = sc[SC_PCG_FLAG] != 0.0;
  const bool finite = isfinite(cost_new) && isfinite(pred) && !pcg_bad;
  const double actual = finite ? cost - cost_new : -1.0;
  const double ratio = (finite && pred > 0) ? actual / pred : -1.0;
  const double step_norm = sqrt(step2), x_norm = sqrt(x2);
  const bool ft = finite && actual < st->ftol * cost && ratio > 0.25;
  const bool xt = finite && step_norm < st->xtol * (st->xtol + x_norm);
  const int term = (ft && xt) ? 4 : ft ? 2 : xt ? 3 : 0;
  if (log != nullptr && st->n_log < st->log_cap) {
    LmLogRow& r = log[st->n_log];
    r.nit = (double)st->nit; r.nfev = (double)st->nfev; r.cost = cost; r.cost_new = cost_new; r.ratio = ratio;
    r.lam = st->lam; r.step = step_norm; r.gnorm = st->gnorm; r.pcg = sc[SC_PCG_ITS];
    st->n_log += 1;
  }
  double lam = st->lam;
  if (finite && actual > 0) {
    st->cur ^= 1;
    const double t = 2.0 * ratio - 1.0;
    lam = fmax(lam * fmax(1.0 / 3.0, 1.0 - t * t * t), 1e-15);
    st->nu = 2.0;
    st->cost = cost_new;
    st->new_lin = 1;
    st->bad_streak = 0;
    if (term) { st->status = term; st->done = 1; }
    else st->njev += 1;
  } else {
    lam = fmin(lam * st->nu, 1e12);
    st->nu *= 2.0;
    st->new_lin = 0;
    if (term) { st->status = term; st->done = 1; }
    if (!finite) {
      // every trial non-finite with the damping at its cap: nothing can change any more (scipy would raise or stop
      // on max_nfev = 100 n; stop now with status 0 instead of spinning for millions of evaluations)
      if (++st->bad_streak >= 6 && lam >= 1e12 && !st->done) { st->status = 0; st->err = LM_ERR_STUCK_NONFINITE; st->done = 1; }
    } else {
      st->bad_streak = 0;
    }
  }
  st->lam = lam;
}

// ---------------------------------------------------------------------------------------------
// Chunk partials of the camera pass -> per-camera packed U, g and the cost; block n_cams sums the point-step
// partials.  The last block to finish adds up the cost in fixed order and (MODE 1) takes the decision.
//   MODE 0: initial linearisation (writes slot cur, no decision)
//   MODE 1: trial point, single GPU: decision fused
//   MODE 2: trial point, several GPUs: red2 goes through the all-reduce, lm_decide_kernel follows
// ---------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(64)
trial_reduce_kernel(LmState* __restrict__ st, int mode, int n_cams, const int* __restrict__ cam_chunk_start,
                    const double* __restrict__ partial, Ptr2 Upk2, Ptr2 gc2, Ptr2 costsum2,
                    double* __restrict__ cam_cost, int n_extra_cost, const double* __restrict__ bpart, int bpart_n,
                    int bpart_stride, double* __restrict__ red2, unsigned int* __restrict__ counter,
                    const double* __restrict__ sc, LmLogRow* __restrict__ log) {
  using RT = RowT<P>;
  __shared__ double sh[2];
  __shared__ int s_last;
  if (st->done) return;
  const int sel = (mode == 0) ? st->cur : (st->cur ^ 1);
  const int c = blockIdx.x, k = threadIdx.x;
  if (c < n_cams) {
    if (k < RT::NACC) {
      double v = 0.0;
      for (int ch = cam_chunk_start[c]; ch < cam_chunk_start[c + 1]; ++ch) v += partial[(size_t)ch * RT::NACC + k];
      if (k < RT::NU) Upk2.p[sel][(size_t)c * RT::NU + k] = v;
      else if (k < RT::NU + P) gc2.p[sel][(size_t)c * P + (k - RT::NU)] = v;
      else cam_cost[c] = v;
    }
  } else if (mode != 0) {
    // three deterministic sums of bpart_n values each
    for (int q = 0; q < 3; ++q) {
      const double* src = bpart + (size_t)q * bpart_stride;
      double v = 0.0;
      for (int i = k; i < bpart_n; i += 64) v += src[i];
      v = warp_sum(v);
      if ((k & 31) == 0) sh[k >> 5] = v;
      __syncthreads();
      if (k == 0) red2[1 + q] = sh[0] + sh[1];
      __syncthreads();
    }
  }
  __syncthreads();
  if (k == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(counter, 1u);
    s_last = (prev == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double v = 0.0;
  for (int i = k; i < n_cams + n_extra_cost; i += 64) v += __ldcg(cam_cost + i);
  v = warp_sum(v);
  if ((k & 31) == 0) sh[k >> 5] = v;
  __syncthreads();
  if (k == 0) {
    const double cost = sh[0] + sh[1];
    costsum2.p[sel][0] = cost;
    *counter = 0u;
    if (mode != 0) {
      red2[0] = cost;
      if (mode == 1) {
        __threadfence();
        const double r2[4] = {cost, __ldcg(red2 + 1), __ldcg(red2 + 2), __ldcg(red2 + 3)};
        lm_decide(st, sc, r2, log);
      }
    }
  }
}

// Last node of the WHILE-loop body (device-loop mode): keep looping until the state machine says done.
__global__ void lm_loop_cond_kernel(const LmState* __restrict__ st, cudaGraphConditionalHandle h) {
  cudaGraphSetConditional(h, st->done ? 0u : 1u);
}

}  // namespace cb
