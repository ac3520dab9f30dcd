// All-reduce over NVLink peer memory, fused with the kernels that produce the data.
//
// Every rank owns one cudaMalloc'd "symmetric" buffer, mapped into all other ranks with CUDA IPC:
//
//   [flags_big u64 x16][flags_small u64 x16][done u32, err i32][small slots 2 x 16 x 4 f64][data 0][data 1]
//
//   * schur_finalize_peer_kernel<P>: ONE launch does (1) the split-K finalize S = U - Z Z^T, b = g_c - Z t into
//     this rank's data[epoch & 1], (2) a release-store of the epoch into every peer's flag word for this rank once
//     the last block is done, (3) an acquire-spin on this rank's own flag words, (4) the sum over all ranks, read
//     straight from the peers' buffers over NVLink in rank order (so every rank gets the bitwise identical sum)
//     into the local reduced-system buffer.  No NCCL launch, no intermediate copy.
//   * peer_small_allreduce_kernel: the 4-double trial cost / predicted reduction / step norms, push model: each
//     rank stores its values into every peer's slot for it, then flag; the sum runs on local memory.
//
// Double buffering makes one barrier per reduction sufficient: a rank overwrites data[k & 1] for reduction k+2 only
// after it passed the barrier of k+1, which every peer signals after it finished reading reduction k.
// A spin that sees no signal for 20 s sets `err` and falls through; the host turns that into CB_E_CALLBACK.
#pragma once
#include "cb_lm.cuh"

namespace cb {

constexpr int PEER_MAXW = 16;
constexpr size_t PEER_OFF_FLAGS_BIG = 0, PEER_OFF_FLAGS_SMALL = 128, PEER_OFF_DONE = 256, PEER_OFF_ERR = 260,
                 PEER_OFF_SMALL = 512, PEER_OFF_DATA = 2048;
constexpr int PEER_SMALL_N = 4;

struct PeerTable {
  int rank, world;
  const double* data[PEER_MAXW][2];              // every rank's two data buffers (own entry = local pointer)
  unsigned long long* flags_big_of[PEER_MAXW];   // rank r's flags_big[16]; this rank writes entry [rank]
  unsigned long long* flags_small_of[PEER_MAXW];
  double* small_of[PEER_MAXW];                   // rank r's small slots [2][16][4]
  double* my_data[2];
  unsigned long long* my_flags_big;
  unsigned long long* my_flags_small;
  double* my_small;
  unsigned int* done;
  int* err;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// poll with relaxed loads (an acquire per iteration would fence every time), one acquire fence at the end
__device__ __forceinline__ void peer_wait(const unsigned long long* flag, unsigned long long epoch, int* err) {
  if (ld_acquire_sys(flag) >= epoch) return;
  const unsigned long long t0 = globaltimer_ns();
  unsigned spins = 0;
  while (ld_relaxed_sys(flag) < epoch) {
    if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > 20000000000ull) {  // 20 s: a peer died or never launched
      atomicExch(err, 1);
      return;
    }
  }
  __threadfence_system();
}

constexpr int PEER_THREADS = 512;

// Launched COOPERATIVELY (cudaLaunchCooperativeKernel): every block spins on the peers' flags, so the whole grid must
// be co-resident; the cooperative launch makes the runtime guarantee it (or fail the launch) instead of trusting an
// occupancy estimate.  The reduction's sequence number is read from the device-resident LM state (st->epoch_big + 1;
// reduced_prep_kernel advances it), so the launch is identical every trial and replays from a CUDA graph.
template <int P>
__global__ void __launch_bounds__(PEER_THREADS, 2)
schur_finalize_peer_kernel(const LmState* __restrict__ st, int nP, int n_blk, const int* __restrict__ tile_of,
                           const int* __restrict__ tile_slot_start, const int* __restrict__ tile_slots,
                           const double* __restrict__ part, const double* __restrict__ tpart, CPtr2 Upk2, CPtr2 gc2,
                           CPtr2 costsum2, const double* __restrict__ gmax, int red_slots, int rank_slot, PeerTable tab,
                           double* __restrict__ red) {
  if (st->done) return;  // identical on every rank: the decision is taken on all-reduced numbers
  const int cur = st->cur;
  const unsigned long long epoch = st->epoch_big + 1ull;
  const int parity = (int)(epoch & 1ull);
  double* mine = tab.my_data[parity];
  const size_t nn = (size_t)nP * nP, nfin = nn + nP + 1, slot0 = nn + 3 * (size_t)nP + 1, total = slot0 + red_slots;
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  // (1) this rank's partial reduced system
  for (size_t idx = t0; idx < nfin; idx += stride)
    finalize_elem<P>(idx, nP, n_blk, tile_of, tile_slot_start, tile_slots, part, tpart, Upk2.p[cur], gc2.p[cur],
                     costsum2.p[cur], mine);
  for (size_t s = t0; s < (size_t)red_slots; s += stride) mine[slot0 + s] = ((int)s == rank_slot) ? gmax[0] : 0.0;
  // (2) last block to finish publishes the epoch to every peer
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned prev = atomicAdd(tab.done, 1u);
    __threadfence_system();
    s_last = (prev == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0) *tab.done = 0u;
    if ((int)threadIdx.x < tab.world) st_release_sys(&tab.flags_big_of[threadIdx.x][tab.rank], epoch);
  }
  // (3) wait for every rank's epoch
  if ((int)threadIdx.x < tab.world) peer_wait(&tab.my_flags_big[threadIdx.x], epoch, tab.err);
  __syncthreads();
  // (4) each block sums its slice over the ranks in rank order, read over NVLink (cache-volatile: the buffers are
  //     rewritten every other epoch); all ranks therefore hold the bitwise identical reduced system
  for (size_t k = t0; k < total; k += stride) {
    double s = 0.0;
#pragma unroll 4
    for (int r = 0; r < tab.world; ++r) s += __ldcv(tab.data[r][parity] + k);
    red[k] = s;
  }
}

// n <= PEER_SMALL_N doubles, in place; one warp (stand-alone form, used outside the LM loop)
__global__ void peer_small_allreduce_kernel(double* buf, int n, PeerTable tab, unsigned long long epoch) {
  const int t = threadIdx.x;
  const int parity = (int)(epoch & 1ull);
  double v[PEER_SMALL_N];
#pragma unroll
  for (int k = 0; k < PEER_SMALL_N; ++k) v[k] = (k < n) ? buf[k] : 0.0;
  __syncwarp();
  if (t < tab.world) {
    double* dst = tab.small_of[t] + ((size_t)parity * PEER_MAXW + tab.rank) * PEER_SMALL_N;
#pragma unroll
    for (int k = 0; k < PEER_SMALL_N; ++k) __stcg(dst + k, v[k]);
    __threadfence_system();
    st_release_sys(&tab.flags_small_of[t][tab.rank], epoch);
    peer_wait(&tab.my_flags_small[t], epoch, tab.err);
  }
  __syncwarp();
  if (t < n) {
    double s = 0.0;
    for (int r = 0; r < tab.world; ++r) s += __ldcv(tab.my_small + ((size_t)parity * PEER_MAXW + r) * PEER_SMALL_N + t);
    buf[t] = s;
  }
}

// The accept / reject decision on several GPUs.  use_peer = 1: the 4 trial sums (cost, predicted reduction, step and
// x norms of this rank's points) are summed over the ranks right here (push model over peer memory, rank order =>
// identical on every rank); use_peer = 0: red2 was all-reduced by NCCL / the callback before this launch.  One warp.
__global__ void lm_decide_kernel(LmState* __restrict__ st, const double* __restrict__ sc, double* __restrict__ red2,
                                 LmLogRow* __restrict__ log, PeerTable tab, int use_peer) {
  if (st->done) return;
  const int t = threadIdx.x;
  if (use_peer) {
    const unsigned long long epoch = st->epoch_small + 1ull;
    const int parity = (int)(epoch & 1ull);
    if (t < tab.world) {
      double* dst = tab.small_of[t] + ((size_t)parity * PEER_MAXW + tab.rank) * PEER_SMALL_N;
#pragma unroll
      for (int k = 0; k < PEER_SMALL_N; ++k) __stcg(dst + k, red2[k]);
      __threadfence_system();
      st_release_sys(&tab.flags_small_of[t][tab.rank], epoch);
      peer_wait(&tab.my_flags_small[t], epoch, tab.err);
    }
    __syncwarp();
    if (t < PEER_SMALL_N) {
      double s = 0.0;
      for (int r = 0; r < tab.world; ++r) s += __ldcv(tab.my_small + ((size_t)parity * PEER_MAXW + r) * PEER_SMALL_N + t);
      red2[t] = s;
    }
    __syncwarp();
  }
  if (t == 0) {
    if (use_peer) st->epoch_small += 1ull;
    const double r2[4] = {red2[0], red2[1], red2[2], red2[3]};
    lm_decide(st, sc, r2, log);
  }
}

}  // namespace cb
