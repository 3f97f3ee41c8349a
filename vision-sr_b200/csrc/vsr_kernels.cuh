// vsr_kernels.cuh -- sm_100a kernels of the refinement engine.
//
//   fit_kernel<T,K,P>   one thread-block CLUSTER per (candidate, restart) run: thread 0 of the
//                       leader CTA advances the BFGS state machine (vsr_bfgs.h); whenever it
//                       asks for the objective all threads of all CTAs sweep their slice of
//                       the points -- TMA-staged once into distributed shared memory --
//                       through the interpreter (vsr_interp.h).  Replaces
//                       minimize(safe_loss, x0, 'BFGS') + the lambdified loss
//                       (reference bfgs.py:102-118).
//   eval_kernel<T,K,P>  batched loss (+ gradient) of (program, constants) pairs; grid.y
//                       splits the points.  Replaces bfgs.py:106-112 and :120-132.
//   eval_finalize       deterministic fixed-order sum of the split partials.
//
// Work mapping: lanes stride over points (coalesced column reads), P points per thread
// share one instruction decode; per-thread partial sums are fp64; reduction is
// __shfl_xor_sync inside the warp, shared memory across the CTA, fixed order throughout
// so results are reproducible run to run.
#ifndef VSR_KERNELS_CUH_
#define VSR_KERNELS_CUH_

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "vsr_bfgs.h"
#include "vsr_interp.h"
#include "vsr_isa.h"

namespace vsr {

struct ProgramTable {
  const vsr_insn_t* insns;
  const int32_t* insn_off;  // [C+1]
  const double* imms;
  const int32_t* imm_off;  // [C+1]
  const int32_t* k;        // [C]
};

struct Points {
  const void* X;  // column-major, column stride ldx elements
  const void* y;
  int64_t n;
  int64_t ldx;
};

struct FitArgs {
  ProgramTable pt;
  Points pts;
  const int32_t* run_prog;  // [n_runs] (this launch's group)
  const int32_t* run_slot;  // [n_runs]
  int32_t n_runs;
  int32_t kstride;
  const double* x0;
  double* out_consts;
  double* out_lastx;
  double* out_loss;
  int32_t* out_info;
  int32_t resident;      // 1: every CTA stages its slice of the points into shared memory
  int32_t tma_ok;        // 1: column starts and strides are 16-byte aligned (bulk copies)
  int32_t slice_stride;  // elements between columns of the staged slice
  long long* phase_cycles;  // optional [n_slots][8]: SM cycles per phase of the pass loop (leader thread 0)
  // budgeted rounds: a run that has not finished after `max_passes` sweeps of this launch saves
  // its optimiser image and exits; the next launch resumes it (resume = 1).  Finished runs are
  // flagged in run_done and their clusters exit at once.
  unsigned char* state;     // [n_slots][state_stride] bytes: FitState image | timer | workspace
  int32_t* run_done;        // [n_slots]
  int32_t state_stride;
  int32_t max_passes;       // <= 0: unlimited
  int32_t resume;
  FitOpts O;
};

struct EvalArgs {
  ProgramTable pt;
  Points pts;
  const int32_t* pair_prog;  // [n_pairs]
  const int32_t* pair_row;   // [n_pairs] row of `consts`
  const int32_t* pair_out;   // [n_pairs] row of the outputs
  int32_t n_pairs;
  int32_t kstride;
  const double* consts;
  double* partial;  // [n_pairs][nsplit][K+1]
  int32_t nsplit;
};

// ---- point source: coalesced global loads (read-only path) --------------------------
template <typename T, int P>
struct GlobalPoints {
  const T* X;
  int64_t ldx;
  int64_t idx[P];
  __device__ __forceinline__ T col(unsigned j, int p) const {
    return __ldg(X + (int64_t)j * ldx + idx[p]);
  }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// One pass over points [n0, n1): s += r^2, g[t] += r * d f/d c_t.
// Every thread runs the same number of iterations (tail lanes are masked), so the
// warp stays converged for the reduction that follows.
template <typename T, int K, int P>
__device__ __forceinline__ void sweep_points(const vsr_insn_t* prog, const double* imm,
                                             const T* cst, const T* __restrict__ X,
                                             const T* __restrict__ y, int64_t ldx, int64_t n0,
                                             int64_t n1, double& s, double (&g)[K > 0 ? K : 1]) {
  s = 0.0;
#pragma unroll
  for (int t = 0; t < (K > 0 ? K : 1); ++t) g[t] = 0.0;
  const int nt = blockDim.x;
  const int tid = threadIdx.x;
  Stack<T, K, P> stk;
  GlobalPoints<T, P> xs;
  xs.X = X;
  xs.ldx = ldx;
  for (int64_t base = n0; base < n1; base += (int64_t)nt * P) {
    bool valid[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const int64_t i = base + (int64_t)p * nt + tid;
      valid[p] = i < n1;
      xs.idx[p] = valid[p] ? i : (n1 - 1);
    }
    Dual<T, K> acc[P];
    eval_points<T, K, P>(prog, imm, cst, xs, acc, stk);
#pragma unroll
    for (int p = 0; p < P; ++p) {
      if (valid[p]) {
        const double r = (double)acc[p].v - (double)__ldg(y + xs.idx[p]);
        s += r * r;
#pragma unroll
        for (int t = 0; t < K; ++t) {
          const double gt = r * (double)acc[p].d[t];
          // a tangent that blew up where the value stayed finite (exp(-exp(x)) ...) counts 0
          g[t] += isfinite(gt) ? gt : 0.0;
        }
      }
    }
  }
}

// CTA-wide sum of (s, g[0..K)) into red[0..K]; red needs (nwarps)*(K+1) doubles.
// After the call thread 0 holds the totals in s / g.  Fixed order: lane butterfly, then
// warps 0..W-1.
template <int K>
__device__ __forceinline__ void block_sum(double& s, double (&g)[K > 0 ? K : 1], double* red) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  s = warp_sum(s);
#pragma unroll
  for (int t = 0; t < K; ++t) g[t] = warp_sum(g[t]);
  if (nw == 1) return;
  if (lane == 0) {
    red[warp * (K + 1)] = s;
#pragma unroll
    for (int t = 0; t < K; ++t) red[warp * (K + 1) + 1 + t] = g[t];
  }
  __syncthreads();
  if (warp == 0) {
    // component `lane` summed over the warps in order 0..nw-1 (fixed order), then handed
    // to thread 0
    double acc = 0.0;
    if (lane <= K)
      for (int w = 0; w < nw; ++w) acc += red[w * (K + 1) + lane];
    s = __shfl_sync(0xffffffffu, acc, 0);
#pragma unroll
    for (int t = 0; t < K; ++t) g[t] = __shfl_sync(0xffffffffu, acc, t + 1);
  }
}

// cooperative copy of one program into shared memory
__device__ __forceinline__ void load_program(const ProgramTable& pt, int prog, vsr_insn_t* s_insn,
                                             double* s_imm, int& n_insn, int& n_imm) {
  const int i0 = pt.insn_off[prog], i1 = pt.insn_off[prog + 1];
  const int m0 = pt.imm_off[prog], m1 = pt.imm_off[prog + 1];
  n_insn = i1 - i0;
  n_imm = m1 - m0;
  for (int i = threadIdx.x; i < n_insn; i += blockDim.x) s_insn[i] = pt.insns[i0 + i];
  for (int i = threadIdx.x; i < n_imm; i += blockDim.x) s_imm[i] = pt.imms[m0 + i];
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// ---- TMA (bulk async copy) + mbarrier primitives, sm_90+ PTX ------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "VSR_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra VSR_MBAR_DONE;\n"
      "bra VSR_MBAR_WAIT;\n"
      "VSR_MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> this CTA's shared memory, completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                              uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- point source: this CTA's slice, resident in shared memory ----------------------------------
// VAR operands of the shared-memory copy of the program are rewritten to column SLOTS of
// the slice, so col(j, p) is one LDS.
template <typename T, int P>
struct SlicePoints {
  const T* base;  // [n_slots][stride]
  int stride;
  int idx[P];
  __device__ __forceinline__ T col(unsigned j, int p) const { return base[j * stride + idx[p]]; }
};

// One pass over the resident slice (cnt points): s += r^2, g[t] += r * d f/d c_t.
template <typename T, int K, int P>
__device__ __forceinline__ void sweep_slice(const vsr_insn_t* prog, const double* imm, const T* cst,
                                            const T* xs, const T* ys, int stride, int cnt, double& s,
                                            double (&g)[K > 0 ? K : 1]) {
  s = 0.0;
#pragma unroll
  for (int t = 0; t < (K > 0 ? K : 1); ++t) g[t] = 0.0;
  const int nt = blockDim.x;
  const int tid = threadIdx.x;
  Stack<T, K, P> stk;
  SlicePoints<T, P> src;
  src.base = xs;
  src.stride = stride;
  for (int base = 0; base < cnt; base += nt * P) {
    bool valid[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const int i = base + p * nt + tid;
      valid[p] = i < cnt;
      src.idx[p] = valid[p] ? i : (cnt - 1);
    }
    Dual<T, K> acc[P];
    eval_points<T, K, P>(prog, imm, cst, src, acc, stk);
#pragma unroll
    for (int p = 0; p < P; ++p) {
      if (valid[p]) {
        const double r = (double)acc[p].v - (double)ys[src.idx[p]];
        s += r * r;
#pragma unroll
        for (int t = 0; t < K; ++t) {
          const double gt = r * (double)acc[p].d[t];
          g[t] += isfinite(gt) ? gt : 0.0;
        }
      }
    }
  }
}

// ---- fit kernel ------------------------------------------------------------------------------------
// One thread-block CLUSTER per (candidate, restart) run.  The points are split in
// contiguous slices, one per CTA of the cluster.  When a slice fits, its used columns and
// y are staged ONCE into the CTA's shared memory by TMA bulk copies and stay there for
// every pass of the run (hundreds to thousands): the whole data set lives in the
// cluster's distributed shared memory and HBM/L2 is read once per run.  CTA 0 (the
// leader) owns the optimiser state; per pass the other CTAs read the trial constants from
// the leader's shared memory and write their partial sums into it (DSMEM), with two
// cluster barriers.  Reduction order is fixed (lanes, warps, CTA rank).
//
// dynamic shared memory, in bytes (16-byte aligned sections):
//   ws[fit_workspace_doubles(k)] | cred[cs*(K+1)] | red[nwarps*(K+1)] | cst[k+1] | imm | insn |
//   slice: (n_slots + 1) * stride * sizeof(T)
__host__ __device__ inline size_t fit_smem_bytes(int kmax, int K, int nwarps, int cs, int n_insn,
                                                 int n_imm, int n_slots, int stride, int elem) {
  size_t dbl = (size_t)fit_workspace_doubles(kmax) + (size_t)cs * (K + 1) + (size_t)nwarps * (K + 1) +
               kmax + 1 + n_imm + n_insn + 2;
  dbl = (dbl + 1) & ~(size_t)1;  // 16-byte boundary
  return dbl * 8 + (size_t)(n_slots + 1) * stride * elem;
}

// Widest CTA the fit kernel is compiled for.  320 threads at <= 96 registers lets TWO CTAs
// (of different clusters, i.e. different runs) share an SM: while one run's cluster sits in
// its optimiser step or a cluster barrier the other one sweeps.
template <typename T, int K>
__host__ __device__ constexpr int fit_max_threads() {
  return (sizeof(T) == 8 && K > 8) ? 256 : 320;
}
template <typename T, int K>
__host__ __device__ constexpr int fit_min_ctas() {
  return (sizeof(T) == 8 && K > 8) ? 1 : 2;
}

// The optimiser step, out of line: its register and stack needs stay out of the sweep's
// allocation (the sweep is the hot loop; this runs on one warp between sweeps).
static __device__ __noinline__ int fit_step_call(FitState& S, const FitOpts& O) { return fit_step(S, O); }

template <typename T, int K, int P>
__global__ void __launch_bounds__((fit_max_threads<T, K>()), (fit_min_ctas<T, K>())) fit_kernel(const FitArgs a) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) double smem[];
  FitState S;  // private per lane; only warp 0 of the leader CTA runs the optimiser
  __shared__ double s_rf;
  __shared__ int s_action;
  __shared__ unsigned long long s_t0;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_slot_of[VSR_MAX_VARS];

  const int cs = (int)cluster.num_blocks();
  const int crank = (int)cluster.block_rank();
  const int run = blockIdx.x / cs;
  const int prog = a.run_prog[run];
  const int slot = a.run_slot[run];
  const int k = a.pt.k[prog];
  const int nw = (blockDim.x + 31) >> 5;
  const int tid = threadIdx.x;

  double* ws = smem;
  double* cred = ws + fit_workspace_doubles(k);
  double* red = cred + cs * (K + 1);
  T* cst = reinterpret_cast<T*>(red + nw * (K + 1));
  double* s_imm = red + nw * (K + 1) + k + 1;
  if (a.resume && a.run_done[slot]) return;  // finished in an earlier round; uniform over the cluster
  int n_insn, n_imm;
  const int m_imm = a.pt.imm_off[prog + 1] - a.pt.imm_off[prog];
  vsr_insn_t* s_insn = reinterpret_cast<vsr_insn_t*>(s_imm + m_imm);
  load_program(a.pt, prog, s_insn, s_imm, n_insn, n_imm);

  int32_t* info = a.out_info + (int64_t)slot * 4;
  if (k == 0) {  // nothing to optimise (reference bfgs.py:117-118); uniform over the cluster
    if (crank == 0 && tid == 0) {
      info[0] = VSR_FIT_NOT_RUN;
      info[1] = 0;
      info[2] = 0;
      info[3] = 0;
      a.out_loss[slot] = 0.0;
      if (a.run_done) a.run_done[slot] = 1;
    }
    return;
  }

  // ---- this CTA's slice of the points ----
  const int64_t N = a.pts.n;
  int64_t per = (N + cs - 1) / cs;
  per = (per + 31) & ~(int64_t)31;  // slices start on 32-point boundaries (TMA alignment)
  int64_t n0 = (int64_t)crank * per, n1 = n0 + per;
  n0 = n0 < N ? n0 : N;
  n1 = n1 < N ? n1 : N;
  const int cnt = (int)(n1 - n0);
  const T* X = static_cast<const T*>(a.pts.X);
  const T* y = static_cast<const T*>(a.pts.y);

  T* xs = nullptr;
  T* ys = nullptr;
  const int stride = a.slice_stride;
  if (a.resident) {
    size_t off = (size_t)((reinterpret_cast<char*>(s_insn + n_insn) - reinterpret_cast<char*>(smem)) + 15) &
                 ~(size_t)15;
    ys = reinterpret_cast<T*>(reinterpret_cast<char*>(smem) + off);
    xs = ys + stride;
    __syncthreads();  // program is in shared memory
    constexpr int kAlign = 16 / (int)sizeof(T);
    const int full = cnt & ~(kAlign - 1);  // elements per column that move as 16-byte units
    const bool use_tma = a.tma_ok && full > 0;
    if (tid == 0) {
      // rewrite VAR operands to slice column slots, in first-use order
      for (int j = 0; j < VSR_MAX_VARS; ++j) s_slot_of[j] = -1;
      int ns = 0;
      for (int i = 0; i < n_insn; ++i) {
        const vsr_insn_t w = s_insn[i];
        const unsigned op = VSR_OP(w);
        if (op >= VSR_LOAD && op <= VSR_RPOW && op != VSR_PUSH && VSR_SRC(w) == VSR_SRC_VAR) {
          const unsigned j = VSR_IDX(w);
          if (s_slot_of[j] < 0) s_slot_of[j] = ns++;
          s_insn[i] = (w & ~((vsr_insn_t)0xffff << 16)) | ((vsr_insn_t)s_slot_of[j] << 16);
        }
      }
      if (use_tma) {
        const uint32_t bytes = (uint32_t)full * (uint32_t)sizeof(T);
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&s_bar, bytes * (uint32_t)(ns + 1));
        tma_bulk_load(ys, y + n0, bytes, &s_bar);
        for (int j = 0; j < VSR_MAX_VARS; ++j)
          if (s_slot_of[j] >= 0)
            tma_bulk_load(xs + (size_t)s_slot_of[j] * stride, X + (int64_t)j * a.pts.ldx + n0, bytes, &s_bar);
      }
    }
    __syncthreads();  // slot table, rewritten program and the armed barrier are visible
    {
      // everything TMA does not move (all of it when the caller's memory is unaligned): coalesced loads
      const int first = use_tma ? full : 0;
      for (int i = first + tid; i < cnt; i += blockDim.x) ys[i] = y[n0 + i];
      for (int j = 0; j < VSR_MAX_VARS; ++j) {
        const int sj = s_slot_of[j];
        if (sj < 0) continue;
        for (int i = first + tid; i < cnt; i += blockDim.x)
          xs[(size_t)sj * stride + i] = X[(int64_t)j * a.pts.ldx + n0 + i];
      }
    }
    if (use_tma) mbar_wait(&s_bar, 0);  // every thread observes the completed transaction
    __syncthreads();
  }

  // handler ids over the opcode bytes (after the slot rewrite, which reads raw opcodes)
  __syncthreads();
  for (int i = tid; i < n_insn; i += blockDim.x) s_insn[i] = predecode(s_insn[i]);
  __syncthreads();

  const bool is_logic = crank == 0 && tid < 32;  // warp 0 of the leader: the optimiser
  unsigned char* image = a.state ? a.state + (int64_t)slot * a.state_stride : nullptr;
  constexpr int kImgHead = (int)((sizeof(FitState) + 8 + 15) / 16 * 16);  // FitState | s_t0, 16-aligned
  if (is_logic) {
    if (a.resume) {
      // every lane takes its own copy of the saved scalar state, the vectors go back to
      // shared memory, and the pending evaluation request (S.xe) is served first
      const FitState* saved = reinterpret_cast<const FitState*>(image);
      S = *saved;
      fit_rebase(S, k, ws);
      const double* wsg = reinterpret_cast<const double*>(image + kImgHead);
      for (int i = tid; i < fit_workspace_doubles(k); i += 32) ws[i] = wsg[i];
      if (tid == 0) s_t0 = *reinterpret_cast<const unsigned long long*>(image + sizeof(FitState));
      __syncwarp();
    } else {
      fit_init(S, k, ws, a.x0 + (int64_t)slot * a.kstride);
      if (tid == 0) s_t0 = 0ull;
    }
  }
  // leader state is initialised and every CTA of the cluster is running before any DSMEM access
  cluster.sync();

  const int* r_action = cluster.map_shared_rank(&s_action, 0);
  const double* r_xe = cluster.map_shared_rank(ws, 0);  // FitState.xe is the first k doubles of ws
  double* r_cred = cluster.map_shared_rank(cred, 0);
  const double inv_n = 1.0 / (double)N;

  // optional phase timing (measurement aid): cycles of the leader's thread 0 in
  // [0] optimiser logic  [1] first cluster barrier  [2] constant broadcast  [3] sweep
  // [4] CTA reduction  [5] second cluster barrier  [6] finalisation  [7] passes
  const bool timing = a.phase_cycles != nullptr && crank == 0 && tid == 0;
  long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tprev = timing ? clock64() : 0;
#define VSR_PHASE(i)                 \
  if (timing) {                      \
    const long long tn = clock64();  \
    ph[i] += tn - tprev;             \
    tprev = tn;                      \
  }

  int passes = 0;
  bool pending = a.resume != 0;  // a resumed run re-enters at its saved evaluation request
  for (;;) {
    if (is_logic) {
      int act = VSR_NEED_EVAL;
      if (!pending) act = fit_step_call(S, a.O);
      if (act == VSR_NEED_EVAL && a.max_passes > 0 && passes >= a.max_passes) {
        // budget of this round used up: save the image, stop here (the request stays pending)
        act = VSR_PAUSE;
        __syncwarp();
        if (tid == 0) {
          *reinterpret_cast<FitState*>(image) = S;
          *reinterpret_cast<unsigned long long*>(image + sizeof(FitState)) = s_t0;
        }
        double* wsg = reinterpret_cast<double*>(image + kImgHead);
        for (int i = tid; i < fit_workspace_doubles(k); i += 32) wsg[i] = ws[i];
      }
      if (tid == 0) s_action = act;
    }
    pending = false;
    ++passes;
    VSR_PHASE(0)
    cluster.sync();
    VSR_PHASE(1)
    if (*r_action != VSR_NEED_EVAL) break;
    // trial constants from the leader, in the arithmetic type of the sweep
    for (int i = tid; i < k; i += blockDim.x) cst[i] = (T)r_xe[i];
    __syncthreads();
    VSR_PHASE(2)
    double s, g[K > 0 ? K : 1];
    if (a.resident)
      sweep_slice<T, K, P>(s_insn, s_imm, cst, xs, ys, stride, cnt, s, g);
    else
      sweep_points<T, K, P>(s_insn, s_imm, cst, X, y, a.pts.ldx, n0, n1, s, g);
    VSR_PHASE(3)
    block_sum<K>(s, g, red);
    if (tid == 0) {
      r_cred[crank * (K + 1)] = s;
#pragma unroll
      for (int t = 0; t < K; ++t) r_cred[crank * (K + 1) + 1 + t] = g[t];
    }
    VSR_PHASE(4)
    cluster.sync();
    VSR_PHASE(5)
    if (is_logic) {
      // component `tid` of (sum r^2, sum r df/dc_t) over the CTAs of the cluster in rank
      // order; lane 0 applies the penalty rule, lanes 1..k scale the gradient
      double tot = 0.0;
      if (tid <= K)
        for (int r = 0; r < cs; ++r) tot += cred[r * (K + 1) + tid];
      const double f = a.O.loss_scale * (__shfl_sync(0xffffffffu, tot, 0) * inv_n);
      bool bad = !isfinite(f);
      if (a.O.stop_time < 1e8) {  // TimedFun (bfgs.py:29-33): the clock starts at the first call
        int late = 0;
        if (tid == 0) {
          const unsigned long long now = global_ns();
          if (s_t0 == 0ull)
            s_t0 = now;
          else if ((double)(now - s_t0) * 1e-9 >= a.O.stop_time)
            late = 1;
        }
        if (__shfl_sync(0xffffffffu, late, 0)) bad = true;
      }
      if (tid == 0) s_rf = bad ? a.O.penalty : f;
      if (tid >= 1 && tid <= K && tid - 1 < k) {
        const double gv = a.O.loss_scale * (2.0 * tot * inv_n);
        S.rg[tid - 1] = (bad || !isfinite(gv)) ? 0.0 : gv;
      }
      if (tid == 0) ph[7] += 1;
    }
    if (is_logic) {  // the response, to every lane's private optimiser state
      __syncwarp();
      S.rf = s_rf;
    }
    VSR_PHASE(6)
    // the leader's warp 0 goes straight back into fit_step; everyone else waits at the
    // cluster barrier above
  }
#undef VSR_PHASE
  if (timing)
    for (int i = 0; i < 8; ++i) a.phase_cycles[(int64_t)slot * 8 + i] += ph[i];

  if (crank == 0 && tid == 0 && s_action == VSR_DONE) {
    if (a.run_done) a.run_done[slot] = 1;
    double* oc = a.out_consts + (int64_t)slot * a.kstride;
    double* ol = a.out_lastx + (int64_t)slot * a.kstride;
    for (int i = 0; i < k; ++i) {
      oc[i] = S.xk[i];
      ol[i] = S.lastx[i];
    }
    a.out_loss[slot] = S.old_fval;
    info[0] = S.status;
    info[1] = S.it;
    info[2] = S.nfev;
    info[3] = 0;
  }
  // no CTA may exit while another can still read its shared memory
  cluster.sync();
}

// dynamic shared memory of eval_kernel, in doubles: red | cst[k] | imm | insn
template <typename T, int K, int P>
__global__ void __launch_bounds__(256) eval_kernel(const EvalArgs a) {
  extern __shared__ double smem[];
  const int pair = blockIdx.x;
  const int split = blockIdx.y;
  if (pair >= a.n_pairs) return;
  const int prog = a.pair_prog[pair];
  const int k = a.pt.k[prog];
  const int nw = (blockDim.x + 31) >> 5;
  double* red = smem;
  T* cst = reinterpret_cast<T*>(red + nw * (K + 1));
  double* s_imm = red + nw * (K + 1) + k + 1;
  const int m_imm = a.pt.imm_off[prog + 1] - a.pt.imm_off[prog];
  vsr_insn_t* s_insn = reinterpret_cast<vsr_insn_t*>(s_imm + m_imm);
  int n_insn, n_imm;
  load_program(a.pt, prog, s_insn, s_imm, n_insn, n_imm);
  const double* c = a.consts + (int64_t)a.pair_row[pair] * a.kstride;
  for (int i = threadIdx.x; i < k; i += blockDim.x) cst[i] = (T)c[i];
  __syncthreads();
  for (int i = threadIdx.x; i < n_insn; i += blockDim.x) s_insn[i] = predecode(s_insn[i]);
  __syncthreads();

  const int64_t N = a.pts.n;
  // split the points in chunks that are multiples of the CTA tile so every split but the
  // last is full
  const int64_t tile = (int64_t)blockDim.x * P;
  const int64_t tiles = (N + tile - 1) / tile;
  const int64_t per = (tiles + a.nsplit - 1) / a.nsplit;
  int64_t n0 = (int64_t)split * per * tile;
  int64_t n1 = (int64_t)(split + 1) * per * tile;
  n0 = n0 < N ? n0 : N;
  n1 = n1 < N ? n1 : N;

  double s, g[K > 0 ? K : 1];
  sweep_points<T, K, P>(s_insn, s_imm, cst, static_cast<const T*>(a.pts.X),
                        static_cast<const T*>(a.pts.y), a.pts.ldx, n0, n1, s, g);
  block_sum<K>(s, g, red);
  if (threadIdx.x == 0) {
    double* out = a.partial + ((int64_t)pair * a.nsplit + split) * (K + 1);
    out[0] = s;
#pragma unroll
    for (int t = 0; t < K; ++t) out[1 + t] = g[t];
  }
}

#if defined(VSR_API_TU)  // plain kernels: defined once, in the API translation unit
// out_loss[row] = sum_splits partial / N ; out_grad[row][t] = 2 * sum / N  (t < k, else 0)
__global__ void eval_finalize(const double* partial, const int32_t* pair_prog,
                              const int32_t* pair_out, const int32_t* prog_k, int n_pairs,
                              int nsplit, int K, int kstride, double inv_n, double* out_loss,
                              double* out_grad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int pair = idx / (K + 1);
  const int comp = idx - pair * (K + 1);
  if (pair >= n_pairs) return;
  double acc = 0.0;
  for (int sidx = 0; sidx < nsplit; ++sidx)
    acc += partial[((int64_t)pair * nsplit + sidx) * (K + 1) + comp];
  const int row = pair_out[pair];
  if (comp == 0) {
    out_loss[row] = acc * inv_n;
  } else if (out_grad != nullptr) {
    const int t = comp - 1;
    if (t < kstride) out_grad[(int64_t)row * kstride + t] = t < prog_k[pair_prog[pair]] ? 2.0 * acc * inv_n : 0.0;
  }
}

// fill rows of a [n][kstride] f64 array with nan (gradients of pairs too wide for duals)
__global__ void fill_nan_rows(const int32_t* rows, int n_rows, int kstride, double* out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rows * kstride) return;
  out[(int64_t)rows[idx / kstride] * kstride + idx % kstride] = __longlong_as_double(0x7ff8000000000000ll);
}

#endif  // VSR_API_TU

}  // namespace vsr

#endif  // VSR_KERNELS_CUH_
