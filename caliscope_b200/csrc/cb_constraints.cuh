// Rigid-distance constraint rows (reference: src/caliscope/core/reprojection.py:112-117 residual,
// :207-226 Jacobian; built by capture_volume.py:446-516) inside the Schur-complement engine.
//
// A constraint row couples up to 8 world points (two width-4 endpoint groups), so the point block of
// the normal equations is no longer 3x3-block-diagonal.  Points linked by constraints form connected
// COMPONENTS (typically: the corners of one board in one frame); each component is eliminated as one
// dense block:  E_comp = blockdiag(V_q + lam D_q) + K^T K = L L^T,  Z_comp = W_comp L^-T,  t = L^-1 g.
// The component's rows of the k-major Schur factor Zt / tvec are the rows 3*point+axis of its own points
// (an L-transformed basis of the component's coordinates), so the SYRK, the reduced solve and the dense
// Zt^T dc product are unchanged; only the per-point build / back-substitution differ for these points.
#pragma once
#include "cb_kernels.cuh"

namespace cb {

constexpr int CC_THREADS = 256;
constexpr int CC_MAX_UNIQ = 8;
constexpr int CC_SMEM_DIM = 128;  // components up to this dimension are factorised in shared memory

struct ConstraintTables {
  int n_c, n_comp, n_dim_max;
  const int* c_nu;        // unique member points per constraint (<= 8)
  const int* c_gidx;      // [n_c][8] global point ids
  const int* c_lidx;      // [n_c][8] index within the component
  const double* c_coef;   // [n_c][8] (count in group a - count in group b) / 4
  const double* c_dist;
  const double* c_w;
  const int* comp_pt_start;  // [n_comp+1]
  const int* comp_pts;       // global point ids per component
  const int* comp_c_start;   // [n_comp+1]
  const int* comp_cons;      // constraint ids per component
  const long long* comp_L_off;  // offset (doubles) of the component's n x n factor
  const int* pt_comp;        // [n_pts] component id or -1
};

// residual, robust rescale and scaled direction per constraint; block partial sums of the cost.
//   diff = sum_u coef_u X_u ; r = (|diff| - d) w ; row of the Jacobian for point u: coef_u * dirw^T
//   with dirw = jscale * w * diff/|diff| (zero sub-gradient at coincident endpoints).
// COST_ONLY: only the partial sums.
// With `st` it is one step of the LM trial: evaluates at point buffer (st->cur ^ flip) and writes the scaled residual
// and direction into slot (st->cur ^ flip) of c_rs2 / c_dirw2 (the trial point's rows become the next linearisation's
// if the step is accepted).
template <bool COST_ONLY>
__global__ void __launch_bounds__(CC_THREADS)
constraint_eval_kernel(const LmState* __restrict__ st, int flip, ConstraintTables T, CPtr2 xp2, int loss, double fscale,
                       Ptr2 c_rs2, Ptr2 c_dirw2, double* __restrict__ raw_r, double* __restrict__ cost_part) {
  __shared__ double sh[CC_THREADS / 32];
  int sel = 0;
  if (st != nullptr) {
    if (st->done) return;
    sel = st->cur ^ flip;
    loss = st->loss;
    fscale = st->fscale;
  }
  const double* __restrict__ xp4 = xp2.p[sel];
  double* __restrict__ c_rs = c_rs2.p[sel];
  double* __restrict__ c_dirw = c_dirw2.p[sel];
  const int k = blockIdx.x * CC_THREADS + threadIdx.x;
  double cost = 0.0;
  if (k < T.n_c) {
    double d0 = 0.0, d1 = 0.0, d2 = 0.0;
    const int nu = T.c_nu[k];
    for (int u = 0; u < nu; ++u) {
      const double* X = xp4 + 4 * (size_t)T.c_gidx[k * CC_MAX_UNIQ + u];
      const double c = T.c_coef[k * CC_MAX_UNIQ + u];
      d0 = fma(c, X[0], d0); d1 = fma(c, X[1], d1); d2 = fma(c, X[2], d2);
    }
    const double nrm = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    const double w = T.c_w[k];
    double r = (nrm - T.c_dist[k]) * w;
    if (raw_r) raw_r[k] = r;
    if (COST_ONLY) {
      cost = robust_cost_only(loss, fscale, r);
    } else {
      double js;
      cost = robust_row(loss, fscale, r, js);
      const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
      c_rs[k] = r;
      if (raw_r) {  // unscaled direction for the Jacobian mirror
        c_dirw[3 * (size_t)k] = w * d0 * inv; c_dirw[3 * (size_t)k + 1] = w * d1 * inv; c_dirw[3 * (size_t)k + 2] = w * d2 * inv;
      } else {
        const double s = js * w * inv;
        c_dirw[3 * (size_t)k] = s * d0; c_dirw[3 * (size_t)k + 1] = s * d1; c_dirw[3 * (size_t)k + 2] = s * d2;
      }
    }
  }
  cost = warp_sum(cost);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = cost;
  __syncthreads();
  if (threadIdx.x == 0 && cost_part) {
    double s = 0.0;
    for (int w = 0; w < CC_THREADS / 32; ++w) s += sh[w];
    cost_part[blockIdx.x] = s;
  }
}

// forward substitution L y = rhs for NRHS right-hand sides held column-major in shared memory
// (rhs[c*n + i]); one warp per column, in place.  L: row-major n x n, lower triangle valid.
__device__ __forceinline__ void forward_solve_cols(const double* L, int n, double* rhs, int ncols, int wid, int lane,
                                                   int nwarps) {
  for (int c = wid; c < ncols; c += nwarps) {
    double* y = rhs + (size_t)c * n;
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int k = lane; k < i; k += 32) s = fma(L[(size_t)i * n + k], y[k], s);
      s = warp_sum(s);
      if (lane == 0) y[i] = (y[i] - s) / L[(size_t)i * n + i];
      __syncwarp();
    }
  }
}

// One CTA per component: assemble E and g, Marquardt scale, Cholesky, t = L^-1 g, Z = W L^-T per camera.
template <int P>
__global__ void __launch_bounds__(CC_THREADS)
comp_build_kernel(const LmState* __restrict__ st, ConstraintTables T, const int* __restrict__ pt_start,
                  const int* __restrict__ pm_cam, const double* __restrict__ V6, const double* __restrict__ gp,
                  double* __restrict__ Dp2, double* __restrict__ gpt, CPtr2 c_rs2, CPtr2 c_dirw2, int n_cams,
                  double* __restrict__ compL, double* __restrict__ tvec, double* __restrict__ Zt, size_t LD,
                  unsigned long long* __restrict__ gmax_bits) {
  extern __shared__ __align__(16) double csm[];
  if (st->done) return;
  const double lam = st->lam;
  const double* __restrict__ c_rs = c_rs2.p[st->cur];
  const double* __restrict__ c_dirw = c_dirw2.p[st->cur];
  const int comp = blockIdx.x;
  const int p0 = T.comp_pt_start[comp], m = T.comp_pt_start[comp + 1] - p0, n = 3 * m;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = CC_THREADS / 32;
  // shared layout: g[n] | Wt[P][n] | cam bitmap (ints) | E[n][n] if n <= CC_SMEM_DIM
  double* g = csm;
  double* Wt = g + T.n_dim_max;
  unsigned int* cambits = reinterpret_cast<unsigned int*>(Wt + (size_t)P * T.n_dim_max);
  double* Esm = reinterpret_cast<double*>(cambits + ((n_cams + 31) / 32 + 1) / 2 * 2 + 2);
  double* E = (n <= CC_SMEM_DIM) ? Esm : compL + T.comp_L_off[comp];
  __shared__ double s_gmax[NW];
  __shared__ int s_fail;

  for (int i = tid; i < n * n; i += CC_THREADS) E[i] = 0.0;
  for (int i = tid; i < (n_cams + 31) / 32; i += CC_THREADS) cambits[i] = 0u;
  if (tid == 0) s_fail = 0;
  __syncthreads();
  // observation part: diagonal 3x3 blocks and gradient
  for (int i = tid; i < m; i += CC_THREADS) {
    const int q = T.comp_pts[p0 + i];
    const double* v = V6 + 6 * (size_t)q;
    double* e = E + (size_t)(3 * i) * n + 3 * i;
    e[0] = v[0]; e[1] = v[1]; e[2] = v[2];
    e[n] = v[1]; e[n + 1] = v[3]; e[n + 2] = v[4];
    e[2 * n] = v[2]; e[2 * n + 1] = v[4]; e[2 * n + 2] = v[5];
    g[3 * i] = gp[3 * (size_t)q]; g[3 * i + 1] = gp[3 * (size_t)q + 1]; g[3 * i + 2] = gp[3 * (size_t)q + 2];
    // cameras that see this point
    for (int pos = pt_start[q]; pos < pt_start[q + 1]; ++pos) atomicOr(&cambits[pm_cam[pos] >> 5], 1u << (pm_cam[pos] & 31));
  }
  __syncthreads();
  // constraint part, one constraint at a time (ordered, deterministic): E += j^T j, g += j^T rs
  for (int cc = T.comp_c_start[comp]; cc < T.comp_c_start[comp + 1]; ++cc) {
    const int k = T.comp_cons[cc];
    const int nu = T.c_nu[k];
    const double rs = c_rs[k];
    const double d[3] = {c_dirw[3 * (size_t)k], c_dirw[3 * (size_t)k + 1], c_dirw[3 * (size_t)k + 2]};
    const int tot = nu * 3;
    for (int e = tid; e < tot * tot; e += CC_THREADS) {
      const int ra = e / tot, rb = e % tot;
      const int u = ra / 3, a = ra % 3, v = rb / 3, b = rb % 3;
      const double ju = T.c_coef[k * CC_MAX_UNIQ + u] * d[a], jv = T.c_coef[k * CC_MAX_UNIQ + v] * d[b];
      E[(size_t)(3 * T.c_lidx[k * CC_MAX_UNIQ + u] + a) * n + 3 * T.c_lidx[k * CC_MAX_UNIQ + v] + b] += ju * jv;
    }
    if (tid < tot) {
      const int u = tid / 3, a = tid % 3;
      g[3 * T.c_lidx[k * CC_MAX_UNIQ + u] + a] += T.c_coef[k * CC_MAX_UNIQ + u] * d[a] * rs;
    }
    __syncthreads();
  }
  // Marquardt scale (running max of the diagonal of J^T J incl. constraint rows), damping, total gradient
  double gm = 0.0;
  for (int i = tid; i < n; i += CC_THREADS) {
    const int q = T.comp_pts[p0 + i / 3];
    double* dq = Dp2 + 3 * (size_t)q + (i % 3);
    double D = fmax(*dq, E[(size_t)i * n + i]);  // running max; idempotent while the point is unchanged
    *dq = D;
    E[(size_t)i * n + i] += lam * (D > 0.0 ? D : 1.0);
    gpt[3 * (size_t)q + (i % 3)] = g[i];
    gm = fmax(gm, fabs(g[i]));
  }
  gm = warp_max(gm);
  if (lane == 0) s_gmax[wid] = gm;
  __syncthreads();
  if (tid == 0) {
    double mx = 0.0;
    for (int w = 0; w < NW; ++w) mx = fmax(mx, s_gmax[w]);
    if (mx > 0.0) atomicMax(gmax_bits, (unsigned long long)__double_as_longlong(mx));
  }
  // Cholesky, right-looking, in place (lower triangle)
  for (int k = 0; k < n; ++k) {
    if (tid == 0) {
      const double d = E[(size_t)k * n + k];
      if (!(d > 0.0)) s_fail = 1;
      E[(size_t)k * n + k] = sqrt(d > 0.0 ? d : 1.0);
    }
    __syncthreads();
    const double dk = E[(size_t)k * n + k];
    for (int i = k + 1 + tid; i < n; i += CC_THREADS) E[(size_t)i * n + k] /= dk;
    __syncthreads();
    const int rem = n - k - 1;
    for (int e = tid; e < rem * rem; e += CC_THREADS) {
      const int i = k + 1 + e / rem, j = k + 1 + e % rem;
      if (j <= i) E[(size_t)i * n + j] -= E[(size_t)i * n + k] * E[(size_t)j * n + k];
    }
    __syncthreads();
  }
  if (s_fail) {  // not positive definite: freeze the component (zero step)
    for (int i = tid; i < n * n; i += CC_THREADS) E[i] = ((i / n) == (i % n)) ? 1e150 : 0.0;
    __syncthreads();
  }
  // t = L^-1 g
  forward_solve_cols(E, n, g, 1, wid, lane, NW);
  __syncthreads();
  for (int i = tid; i < n; i += CC_THREADS) tvec[3 * (size_t)T.comp_pts[p0 + i / 3] + (i % 3)] = g[i];
  // Z rows per camera that sees the component: Wt (n x P) = sum over its observations, Y = L^-1 Wt
  for (int c = 0; c < n_cams; ++c) {
    if (!((cambits[c >> 5] >> (c & 31)) & 1u)) continue;  // block-uniform
    for (int i = tid; i < P * n; i += CC_THREADS) Wt[i] = 0.0;
    __syncthreads();
    for (int i = tid; i < m; i += CC_THREADS) {
      const int q = T.comp_pts[p0 + i];
      // W = sum_rows Jc^T Jp of the (camera, point) pair was left in the point's Zt rows by pt_pass_kernel (identity
      // factor for component points); entries of cameras that do not see the point hold last trial's transformed
      // values and count as zero
      for (int pos = pt_start[q]; pos < pt_start[q + 1]; ++pos) {
        if (pm_cam[pos] != c) continue;
        const double* w = Zt + (3 * (size_t)q) * LD + (size_t)c * P;
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
          for (int a = 0; a < 3; ++a) Wt[(size_t)p * n + 3 * i + a] = w[(size_t)a * LD + p];
        break;
      }
    }
    __syncthreads();
    forward_solve_cols(E, n, Wt, P, wid, lane, NW);
    __syncthreads();
    for (int e = tid; e < P * n; e += CC_THREADS) {
      const int p = e / n, i = e % n;
      Zt[(3 * (size_t)T.comp_pts[p0 + i / 3] + (i % 3)) * LD + (size_t)c * P + p] = Wt[(size_t)p * n + i];
    }
    __syncthreads();
  }
  // keep the factor for the back-substitution
  if (n <= CC_SMEM_DIM) {
    double* Lg = compL + T.comp_L_off[comp];
    for (int i = tid; i < n * n; i += CC_THREADS) Lg[i] = E[i];
  }
}

// One CTA per component: dp = -L^-T (t + Zt_rows dc), new points, predicted-reduction partial sums
__global__ void __launch_bounds__(CC_THREADS)
comp_backsub_kernel(const LmState* __restrict__ st, ConstraintTables T, int nP, const double* __restrict__ Zt, size_t LD,
                    const double* __restrict__ dc, const double* __restrict__ compL, const double* __restrict__ tvec,
                    const double* __restrict__ gpt, const double* __restrict__ Dp2, Ptr2 xp2,
                    double* __restrict__ dp_out, double* __restrict__ bpart, int bpart_stride, int bpart_off) {
  extern __shared__ __align__(16) double bsm[];
  if (st->done) return;
  const double lam = st->lam;
  const double* __restrict__ xp4 = xp2.p[st->cur];
  double* __restrict__ xp4_new = xp2.p[st->cur ^ 1];
  const int comp = blockIdx.x;
  const int p0 = T.comp_pt_start[comp], m = T.comp_pt_start[comp + 1] - p0, n = 3 * m;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = CC_THREADS / 32;
  double* dcs = bsm;        // nP
  double* v = dcs + nP;     // n_dim_max
  const double* L = compL + T.comp_L_off[comp];
  __shared__ double wsum[3][NW];
  for (int i = tid; i < nP; i += CC_THREADS) dcs[i] = dc[i];
  __syncthreads();
  for (int i = wid; i < n; i += NW) {
    const double* z = Zt + (3 * (size_t)T.comp_pts[p0 + i / 3] + (i % 3)) * LD;
    double u = 0.0;
    for (int k = lane; k < nP; k += 32) u = fma(z[k], dcs[k], u);
    u = warp_sum(u);
    if (lane == 0) v[i] = -(tvec[3 * (size_t)T.comp_pts[p0 + i / 3] + (i % 3)] + u);
  }
  __syncthreads();
  // back substitution L^T d = v (single warp; L row-major, column access)
  if (wid == 0) {
    for (int i = n - 1; i >= 0; --i) {
      double s = 0.0;
      for (int k = i + 1 + lane; k < n; k += 32) s = fma(L[(size_t)k * n + i], v[k], s);
      s = warp_sum(s);
      if (lane == 0) v[i] = (v[i] - s) / L[(size_t)i * n + i];
      __syncwarp();
    }
  }
  __syncthreads();
  double pred = 0.0, st2 = 0.0, x2 = 0.0;
  for (int i = tid; i < m; i += CC_THREADS) {
    const int q = T.comp_pts[p0 + i];
    const double* xo = xp4 + 4 * (size_t)q;
    double* xn = xp4_new + 4 * (size_t)q;
    const double d0 = v[3 * i], d1 = v[3 * i + 1], d2 = v[3 * i + 2];
    xn[0] = xo[0] + d0; xn[1] = xo[1] + d1; xn[2] = xo[2] + d2; xn[3] = 0.0;
    if (dp_out) { dp_out[3 * (size_t)q] = d0; dp_out[3 * (size_t)q + 1] = d1; dp_out[3 * (size_t)q + 2] = d2; }
    const double* D = Dp2 + 3 * (size_t)q;
    const double* gq = gpt + 3 * (size_t)q;
    const double e0 = D[0] > 0.0 ? D[0] : 1.0, e1 = D[1] > 0.0 ? D[1] : 1.0, e2 = D[2] > 0.0 ? D[2] : 1.0;
    pred += 0.5 * (d0 * (lam * e0 * d0 - gq[0]) + d1 * (lam * e1 * d1 - gq[1]) + d2 * (lam * e2 * d2 - gq[2]));
    st2 += d0 * d0 + d1 * d1 + d2 * d2;
    x2 += xo[0] * xo[0] + xo[1] * xo[1] + xo[2] * xo[2];
  }
  pred = warp_sum(pred); st2 = warp_sum(st2); x2 = warp_sum(x2);
  if (lane == 0) { wsum[0][wid] = pred; wsum[1][wid] = st2; wsum[2][wid] = x2; }
  __syncthreads();
  if (tid == 0) {
    double a = 0, b = 0, c = 0;
    for (int w = 0; w < NW; ++w) { a += wsum[0][w]; b += wsum[1][w]; c += wsum[2][w]; }
    bpart[bpart_off + comp] = a;
    bpart[bpart_stride + bpart_off + comp] = b;
    bpart[2 * (size_t)bpart_stride + bpart_off + comp] = c;
  }
}

}  // namespace cb
