// vsr_kernels.cuh -- sm_100a kernels of the refinement engine.
//
//   fit_kernel<T,K,P>   PERSISTENT thread-block clusters that keep several (candidate, restart)
//                       runs in flight each ("seats"): reserved warps of the leader CTA advance
//                       the BFGS state machines (vsr_bfgs.h) of one half of the seats while
//                       every other warp of the cluster sweeps its slice of the points --
//                       TMA-staged once into distributed shared memory -- through the
//                       interpreter (vsr_interp.h) for the other half.  Replaces
//                       minimize(safe_loss, x0, 'BFGS') + the lambdified loss
//                       (reference bfgs.py:102-118).
//   eval_kernel<T,K,P>  batched loss (+ gradient) of (program, constants) pairs; grid.y
//                       splits the points.  Replaces bfgs.py:106-112 and :120-132.
//   eval_finalize       deterministic fixed-order sum of the split partials.
//
// Work mapping: lanes stride over points (coalesced column reads), P points per thread
// share one instruction decode; per-thread partial sums are fp64 and parked in shared memory,
// one warp per component sums them (block_totals); fixed order throughout, so results are
// reproducible run to run.
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

constexpr int kMaxSeats = 8;  // runs a cluster can work on at a time

struct FitArgs {
  ProgramTable pt;
  Points pts;
  const int32_t* run_prog;  // [n_runs] (this launch's group), in queue order
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
  int32_t seats;         // runs in flight per cluster (<= warps per CTA, <= kMaxSeats)
  int32_t banks;         // 2: optimiser steps of one half of the seats overlap the sweeps of the other
  int32_t reserved;      // warps of the leader CTA that never sweep (0 or 2), see choose_geometry
  // seat layout in doubles, computed once on the host (fit_seat_layout): the kernel would
  // otherwise re-derive it from (kmax, K, warps, cluster size, ...) at every use
  int32_t seat_d, off_cred, off_cst, off_imm, off_insn;
  int32_t kmax, max_insn, max_imm;  // maxima over this launch's programs (size the seat areas)
  int32_t n_cols;                   // columns of X this launch's programs read
  int32_t col_of_var[VSR_MAX_VARS]; // slice column of variable j (-1: unused)
  int32_t* queue;           // [1] index of the next run to hand out (zeroed by the host)
  long long* phase_cycles;  // optional [n_slots][8]: cycles of the seat's optimiser lane 0:
                            // [0] optimiser logic, [3] everything else while seated, [7] passes; runs of seat 0 also
                            // carry the cluster's [1] barrier 1 [2] fetch [4] sweeps [5] barrier 2
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
  int32_t nan_to_num;  // 1: predictions pass through numpy's nan_to_num before the residual (vsr_score)
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

// keeps a value in a register: without it the compiler re-derives point indices and thread
// ranks from %tid inside every handler that needs them (it rematerialises rather than spend one
// of the 96 registers)
__device__ __forceinline__ void keep_in_register(int& v) { asm volatile("" : "+r"(v)); }

// One pass over points [n0, n1).  The thread's partial sums  sum r^2  and  sum r * d f/d c_t  go
// straight into its slots of the reduction scratch ([(K+1)][nt] doubles, see block_totals): no
// accumulator is live across the interpreter (they used to be spilled around it).
// Every thread runs the same number of iterations (tail lanes are masked), so the warp stays
// converged for the reduction that follows.
// NTN: the prediction passes through numpy's nan_to_num (nan -> 0, +-inf -> +-largest finite
// value of T) before the residual is taken: the drivers' scoring rule (Feynman_test.py:87).
template <typename T, int K, int P, bool NTN = false>
__device__ __forceinline__ void sweep_points(const vsr_insn_t* prog, const double* imm,
                                             const T* cst, const T* __restrict__ X,
                                             const T* __restrict__ y, int64_t ldx, int64_t n0,
                                             int64_t n1, double* scratch, int tid, int nt) {
  // tid / nt: index of this thread among the nt threads that take part in the sweep
  Stack<T, K, P> stk;
  GlobalPoints<T, P> xs;
  xs.X = X;
  xs.ldx = ldx;
  bool first = true;
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
    double ps = 0.0, pg[K > 0 ? K : 1];
#pragma unroll
    for (int t = 0; t < (K > 0 ? K : 1); ++t) pg[t] = 0.0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      if (valid[p]) {
        T pv = acc[p].v;
        if (NTN) {
          const T big = sizeof(T) == 8 ? (T)1.7976931348623157e308 : (T)3.4028234663852886e38;
          pv = pv != pv ? T(0) : (pv > big ? big : (pv < -big ? -big : pv));
        }
        const double r = (double)pv - (double)__ldg(y + xs.idx[p]);
        ps += r * r;
#pragma unroll
        for (int t = 0; t < K; ++t) {
          const double gt = r * (double)acc[p].d[t];
          // a tangent that blew up where the value stayed finite (exp(-exp(x)) ...) counts 0
          // (one compare and a predicated add; nan fails the compare)
          if (fabs(gt) <= 1.7976931348623157e308) pg[t] += gt;
        }
      }
    }
    if (first) {
      scratch[tid] = ps;
#pragma unroll
      for (int t = 0; t < K; ++t) scratch[(1 + t) * nt + tid] = pg[t];
      first = false;
    } else {
      scratch[tid] += ps;
#pragma unroll
      for (int t = 0; t < K; ++t) scratch[(1 + t) * nt + tid] += pg[t];
    }
  }
  if (first) {  // no points at all
    scratch[tid] = 0.0;
#pragma unroll
    for (int t = 0; t < K; ++t) scratch[(1 + t) * nt + tid] = 0.0;
  }
}

// named barrier over the `count` threads that take part in a sweep (count: multiple of 32)
__device__ __forceinline__ void sweep_barrier(int count) {
  asm volatile("bar.sync 1, %0;" ::"r"(count) : "memory");
}

// Totals of (sum r^2, sum r df/dc_t) over the `snt` threads that took part in a sweep, stored to
// out[0..K].  The sweep parked every thread's partial sums in shared memory (scratch:
// [(K+1)][snt] doubles, column per component); component c is summed by warp c % nw: each lane adds its strided share in
// index order, a lane butterfly finishes.  Fixed order, so results are reproducible and depend on
// snt only.  4x fewer instructions than a butterfly over K+1 values in every warp followed by a
// cross-warp stage (the shuffles were 7 % of the kernel's instructions).  Only the taking-part
// threads may call it (named barrier, so the optimiser warps of the leader CTA can stay out).
template <int K>
__device__ __forceinline__ void block_totals(double* scratch, int stid, int snt, double* out) {
  const int lane = stid & 31, warp = stid >> 5, nw = snt >> 5;
  sweep_barrier(snt);  // every thread's partial sums are parked (by the sweep)
  for (int c = warp; c <= K; c += nw) {
    const double* col = scratch + c * snt;
    double acc = 0.0;
    for (int j = lane; j < snt; j += 32) acc += col[j];
    acc = warp_sum(acc);
    if (lane == 0) out[c] = acc;
  }
  sweep_barrier(snt);  // the scratch may be rewritten
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

// One pass over the resident slice (cnt points); partial sums parked like sweep_points does.
template <typename T, int K, int P>
__device__ __forceinline__ void sweep_slice(const vsr_insn_t* prog, const double* imm, const T* cst,
                                            const T* xs, const T* ys, int stride, int cnt, double* scratch,
                                            int tid, int nt) {
  Stack<T, K, P> stk;
  SlicePoints<T, P> src;
  src.base = xs;
  src.stride = stride;
  bool first = true;
  for (int base = 0; base < cnt; base += nt * P) {
    bool valid[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const int i = base + p * nt + tid;
      valid[p] = i < cnt;
      src.idx[p] = valid[p] ? i : (cnt - 1);
      keep_in_register(src.idx[p]);
    }
    Dual<T, K> acc[P];
    eval_points<T, K, P>(prog, imm, cst, src, acc, stk);
    double ps = 0.0, pg[K > 0 ? K : 1];
#pragma unroll
    for (int t = 0; t < (K > 0 ? K : 1); ++t) pg[t] = 0.0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      if (src.idx[p] == base + p * nt + tid) {  // valid[p], from the kept index
        const double r = (double)acc[p].v - (double)ys[src.idx[p]];
        ps += r * r;
#pragma unroll
        for (int t = 0; t < K; ++t) {
          const double gt = r * (double)acc[p].d[t];
          if (fabs(gt) <= 1.7976931348623157e308) pg[t] += gt;
        }
      }
    }
    if (first) {
      scratch[tid] = ps;
#pragma unroll
      for (int t = 0; t < K; ++t) scratch[(1 + t) * nt + tid] = pg[t];
      first = false;
    } else {
      scratch[tid] += ps;
#pragma unroll
      for (int t = 0; t < K; ++t) scratch[(1 + t) * nt + tid] += pg[t];
    }
  }
  if (first) {  // no points in this CTA's slice
    scratch[tid] = 0.0;
#pragma unroll
    for (int t = 0; t < K; ++t) scratch[(1 + t) * nt + tid] = 0.0;
  }
}

// ---- fit kernel ------------------------------------------------------------------------------------
// PERSISTENT thread-block clusters, each working on `seats` (candidate, restart) runs at a time.
//
// A pass of one run is: optimiser step (one warp, 16-21 k cycles of dependent scalar work) ->
// sweep of all points through the interpreter (all warps of all CTAs) -> reduction.  With one run
// per cluster every warp but one idles through the optimiser step and the cluster barriers
// (ncu: two thirds of all warp samples sat in barrier waits).  Here a cluster holds G runs in
// "seats", split in two banks: while the optimiser turns of one bank run on reserved warps of the
// leader CTA, all other warps sweep the requests of the other bank back to back (the schedule is
// described at the loop below).  A seat whose run finishes takes the next run from the launch's
// queue (one atomic), so seats stay full until the queue drains; the last long runs then have
// their cluster to themselves and advance at single-run latency.
//
// The points are split in contiguous slices, one per CTA.  When a slice fits, the columns the
// launch's programs read and y are staged ONCE per cluster into shared memory by TMA bulk
// copies and stay there for every pass of every run the cluster handles.  Reduction order is
// fixed (lanes, warps, CTA rank), so a run's result does not depend on its seat or cluster.
//
// dynamic shared memory (doubles):  per seat [ FitState | ws | cred[cs*(K+1)] | cst | imm | insn ]
//   then the reduction scratch [(K+1)][threads], then the slice: (n_cols + 1) * stride * sizeof(T)
// The optimiser state of a seat lives in SHARED memory, one copy per seat that all 32 lanes of
// the seat's warp read (broadcast) and write (same value, uniform control flow).  As a per-lane
// local variable it was evicted from L1 by every sweep's spill and operand-stack traffic, and
// each pass re-read it from L2: 18-21 k cycles per optimiser step on an otherwise idle SM.
constexpr int kFitStateDoubles = (int)((sizeof(FitState) + 15) / 16 * 2);
// The code area of a seat is  cst[kSeatCstDoubles] | imm[VSR_MAX_IMMS] | insn[...]  with FIXED sizes
// in front of the instructions, so that the interpreter addresses constants, literals and
// instruction words at compile-time offsets from ONE register (see the sweep call sites).
constexpr int kSeatCstDoubles = VSR_MAX_CONSTS + 2;

__host__ __device__ inline size_t fit_seat_doubles(int kmax, int K, int nwarps, int cs, int max_insn,
                                                   int max_imm) {
  size_t d = kFitStateDoubles + (size_t)fit_workspace_doubles(kmax) + (size_t)cs * (K + 1) +
             kSeatCstDoubles + VSR_MAX_IMMS + max_insn + 1;  // + pad word after END
  (void)nwarps;
  return (d + 1) & ~(size_t)1;  // 16-byte multiple
}
// offsets (in doubles) of the parts of a seat, and the seat's size
struct SeatLayout {
  int seat_d, off_ws, off_cred, off_cst, off_imm, off_insn;
};
__host__ __device__ inline SeatLayout fit_seat_layout(int kmax, int K, int nwarps, int cs, int max_insn,
                                                      int max_imm) {
  SeatLayout L;
  L.off_ws = kFitStateDoubles;
  L.off_cred = L.off_ws + fit_workspace_doubles(kmax);
  L.off_cst = L.off_cred + cs * (K + 1);
  L.off_imm = L.off_cst + kSeatCstDoubles;
  L.off_insn = L.off_imm + VSR_MAX_IMMS;
  L.seat_d = (int)fit_seat_doubles(kmax, K, nwarps, cs, max_insn, max_imm);
  return L;
}

__host__ __device__ inline size_t fit_smem_bytes(int seats, int kmax, int K, int nwarps, int cs,
                                                 int max_insn, int max_imm, int n_cols, int stride,
                                                 int elem) {
  return (size_t)seats * fit_seat_doubles(kmax, K, nwarps, cs, max_insn, max_imm) * 8 +
         (size_t)(K + 1) * nwarps * 32 * 8 +  // reduction scratch of block_totals
         (n_cols >= 0 ? (size_t)(n_cols + 1) * stride * elem : 0);
}

// Widest CTA the fit kernel is compiled for: ONE CTA of 640 threads (20 warps at <= 96
// registers) per SM, so a cluster owns its SMs.  Measured on the BASELINE config-2 beams
// (tools/exp_schedules.py, us per 1000 sweeps): 8 CTAs x 640 threads 1034, 16 x 320 with two CTAs
// per SM 1106, 16 x 160 with four 1070: optimiser steps run uncontended and a sweep of
// N = 10 000 is one tile iteration.
#if !defined(VSR_FIT_THREADS)
#define VSR_FIT_THREADS 640
#define VSR_FIT_MINCTAS 1
#endif
template <typename T, int K>
__host__ __device__ constexpr int fit_max_threads() {
  return (sizeof(T) == 8 && K > 8) ? 256 : VSR_FIT_THREADS;
}
template <typename T, int K>
__host__ __device__ constexpr int fit_min_ctas() {
  return (sizeof(T) == 8 && K > 8) ? 1 : VSR_FIT_MINCTAS;
}

// The optimiser step, out of line: its register and stack needs stay out of the sweep's
// allocation (the sweep is the hot loop; this runs on one warp per seat between sweeps).
static __device__ __noinline__ int fit_step_call(FitState& S, const FitOpts& O) { return fit_step(S, O); }

struct SeatCtrl {
  int prog;     // program of the seated run, -1: seat empty
  int k;        // its number of constants
  int fresh;    // 1: the run was seated in its optimiser's last turn (every CTA must load its program)
  int drained;  // 1: the seat's optimiser found the launch's queue empty
};

// what the optimiser warp of a seat carries from one turn to the next (leader CTA only)
struct SeatBook {
  int slot;                     // output row of the seated run
  int pad;
  unsigned long long t0;        // TimedFun clock (bfgs.py:29-33)
  double rf;                    // scratch: lane 0 -> all lanes
  long long t_logic, t_seated, n_pass;  // phase_cycles bookkeeping
};

// One optimiser turn of seat `seat`, run by ONE warp of the leader CTA (any warp: all of the
// seat's state is in shared memory): take in the totals of the sweep that served the seat's last
// request, advance the run to its next request, and when it finishes write its results and seat
// the next run of the launch's queue.  Publishes the seat table entry the sweepers read in the
// next iteration.
template <typename T, int K>
__device__ __forceinline__ void seat_turn(const FitArgs& a, cooperative_groups::cluster_group& cluster, int seat,
                                          int lane, int cs, FitState& S, double* ws, const double* cred,
                                          SeatCtrl& ctrl, T* cst, SeatBook& book, int* s_drained) {
  const bool timing = a.phase_cycles != nullptr && lane == 0;
  long long ta = timing ? clock64() : 0;
  int my_prog = ctrl.prog, my_k = ctrl.k;
  if (my_prog >= 0) {
    // component `lane` of (sum r^2, sum r df/dc_t) over the CTAs of the cluster in rank order;
    // lane 0 applies the penalty rule, lanes 1..k scale the gradient
    const double inv_n = 1.0 / (double)a.pts.n;
    double tot = 0.0;
    if (lane <= K)
      for (int r = 0; r < cs; ++r) tot += cred[r * (K + 1) + lane];
    const double f = a.O.loss_scale * (__shfl_sync(0xffffffffu, tot, 0) * inv_n);
    bool bad = !isfinite(f);
    if (a.O.stop_time < 1e8) {  // TimedFun: the clock starts at the first call
      int late = 0;
      if (lane == 0) {
        const unsigned long long now = global_ns();
        if (book.t0 == 0ull)
          book.t0 = now;
        else if ((double)(now - book.t0) * 1e-9 >= a.O.stop_time)
          late = 1;
      }
      if (__shfl_sync(0xffffffffu, late, 0)) bad = true;
    }
    if (lane == 0) book.rf = bad ? a.O.penalty : f;
    if (lane >= 1 && lane <= K && lane - 1 < my_k) {
      const double gv = a.O.loss_scale * (2.0 * tot * inv_n);
      S.rg[lane - 1] = (bad || !isfinite(gv)) ? 0.0 : gv;
    }
    __syncwarp();
    S.rf = book.rf;
    if (lane == 0) book.n_pass += 1;
  }
  int fresh = 0;
  for (;;) {
    if (my_prog < 0) {  // empty seat: take the next run of the launch
      int r = -1;
      if (!*s_drained) {
        if (lane == 0) r = atomicAdd(a.queue, 1);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= a.n_runs) {
          *s_drained = 1;  // every lane, same value
          r = -1;
        }
      }
      if (r < 0) break;
      const int prog = a.run_prog[r];
      const int slot = a.run_slot[r];
      const int k = a.pt.k[prog];
      if (k == 0) {  // nothing to optimise (reference bfgs.py:117-118)
        if (lane == 0) {
          int32_t* info = a.out_info + (int64_t)slot * 4;
          info[0] = VSR_FIT_NOT_RUN;
          info[1] = 0;
          info[2] = 0;
          info[3] = 0;
          a.out_loss[slot] = 0.0;
        }
        continue;
      }
      my_prog = prog;
      my_k = k;
      fresh = 1;
      fit_init(S, k, ws, a.x0 + (int64_t)slot * a.kstride);
      if (lane == 0) {
        book.slot = slot;
        book.t0 = 0ull;
        if (timing) book.t_seated = clock64(), book.t_logic = 0, book.n_pass = 0;
      }
      __syncwarp();
    }
    const int act = fit_step_call(S, a.O);
    if (act == VSR_NEED_EVAL) break;
    // finished: results out, seat free, try to seat another run in this same turn
    __syncwarp();
    if (lane == 0) {
      const int slot = book.slot;
      double* oc = a.out_consts + (int64_t)slot * a.kstride;
      double* ol = a.out_lastx + (int64_t)slot * a.kstride;
      for (int i = 0; i < my_k; ++i) {
        oc[i] = S.xk[i];
        ol[i] = S.lastx[i];
      }
      a.out_loss[slot] = S.old_fval;
      int32_t* info = a.out_info + (int64_t)slot * 4;
      info[0] = S.status;
      info[1] = S.it;
      info[2] = S.nfev;
      info[3] = 0;
      if (timing) {
        const long long now = clock64();
        book.t_logic += now - ta;
        ta = now;
        long long* ph = a.phase_cycles + (int64_t)slot * 8;
        ph[0] += book.t_logic;
        ph[3] += (now - book.t_seated) - book.t_logic;
        ph[7] += book.n_pass;
      }
    }
    __syncwarp();
    my_prog = -1;
    fresh = 0;
  }
  // Publish: the seat table entry and the trial constants go to EVERY CTA of the cluster as
  // remote shared-memory stores (visible after the cluster barrier that ends the iteration), so
  // that the sweepers never read remote memory on their critical path.  Nobody reads these
  // locations during this iteration: the seat's bank is not the one being swept.
  __syncwarp();
  {
    SeatCtrl v;
    v.prog = my_prog;
    v.k = my_k;
    v.fresh = fresh;
    v.drained = *s_drained;
    for (int r = lane; r < cs; r += 32) *cluster.map_shared_rank(&ctrl, r) = v;
    if (my_prog >= 0) {
      for (int idx = lane; idx < cs * my_k; idx += 32) {
        const int r = idx / my_k, i = idx - r * my_k;
        cluster.map_shared_rank(cst, r)[i] = (T)S.xe[i];
      }
    }
  }
  if (timing && my_prog >= 0) book.t_logic += clock64() - ta;
}

template <typename T, int K, int P>
__global__ void __launch_bounds__((fit_max_threads<T, K>()), (fit_min_ctas<T, K>())) fit_kernel(const FitArgs a) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) double smem[];
  __shared__ SeatCtrl s_ctrl[kMaxSeats];
  __shared__ SeatBook s_book[kMaxSeats];
  __shared__ int s_drained;
  __shared__ __align__(8) uint64_t s_bar;

  const int cs = (int)cluster.num_blocks();
  const int crank = (int)cluster.block_rank();
  const int nw = (blockDim.x + 31) >> 5;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int G = a.seats;

  const int seat_d = a.seat_d;
  const int wsd = a.off_cred - kFitStateDoubles;
#define VSR_SEAT_STATE(g) (smem + (size_t)(g)*seat_d)
#define VSR_SEAT_WS(g) (VSR_SEAT_STATE(g) + kFitStateDoubles)
#define VSR_SEAT_CRED(g) (VSR_SEAT_STATE(g) + a.off_cred)
#define VSR_SEAT_CST(g) (reinterpret_cast<T*>(VSR_SEAT_STATE(g) + a.off_cst))
#define VSR_SEAT_IMM(g) (VSR_SEAT_STATE(g) + a.off_imm)
#define VSR_SEAT_INSN(g) (reinterpret_cast<vsr_insn_t*>(VSR_SEAT_STATE(g) + a.off_insn))

  // ---- this CTA's slice of the points ----
  // CTAs 1..cs-1 take `per` points each, the leader (rank 0) takes what is left at the end: in
  // the two-bank schedule two of its warps are busy with optimiser steps during every sweep, and
  // the host sizes `per` so that the leader's remainder fits its remaining warps.
  const int64_t N = a.pts.n;
  const int64_t per = a.slice_stride;
  int64_t n0 = crank == 0 ? (int64_t)(cs - 1) * per : (int64_t)(crank - 1) * per;
  int64_t n1 = crank == 0 ? N : n0 + per;
  n0 = n0 < N ? n0 : N;
  n1 = n1 < N ? n1 : N;
  const int cnt = (int)(n1 - n0);
  const T* X = static_cast<const T*>(a.pts.X);
  const T* y = static_cast<const T*>(a.pts.y);

  T* xs = nullptr;
  T* ys = nullptr;
  const int stride = a.slice_stride;
  if (a.resident) {
    ys = reinterpret_cast<T*>(smem + (size_t)G * seat_d + (size_t)(K + 1) * blockDim.x);
    xs = ys + stride;
    constexpr int kAlign = 16 / (int)sizeof(T);
    const int full = cnt & ~(kAlign - 1);  // elements per column that move as 16-byte units
    const bool use_tma = a.tma_ok && full > 0;
    if (tid == 0 && use_tma) {
      const uint32_t bytes = (uint32_t)full * (uint32_t)sizeof(T);
      mbar_init(&s_bar, 1);
      mbar_fence_init();
      mbar_expect_tx(&s_bar, bytes * (uint32_t)(a.n_cols + 1));
      tma_bulk_load(ys, y + n0, bytes, &s_bar);
      for (int j = 0; j < VSR_MAX_VARS; ++j)
        if (a.col_of_var[j] >= 0)
          tma_bulk_load(xs + (size_t)a.col_of_var[j] * stride, X + (int64_t)j * a.pts.ldx + n0, bytes, &s_bar);
    }
    __syncthreads();  // the armed barrier is visible
    {
      // everything TMA does not move (all of it when the caller's memory is unaligned): coalesced loads
      const int first = use_tma ? full : 0;
      for (int i = first + tid; i < cnt; i += blockDim.x) ys[i] = y[n0 + i];
      for (int j = 0; j < VSR_MAX_VARS; ++j) {
        const int sj = a.col_of_var[j];
        if (sj < 0) continue;
        for (int i = first + tid; i < cnt; i += blockDim.x)
          xs[(size_t)sj * stride + i] = X[(int64_t)j * a.pts.ldx + n0 + i];
      }
    }
    if (use_tma) mbar_wait(&s_bar, 0);  // every thread observes the completed transaction
  }
  if (tid < kMaxSeats) {
    s_ctrl[tid].prog = -1;
    s_ctrl[tid].k = 0;
    s_ctrl[tid].fresh = 0;
    s_ctrl[tid].drained = tid < G ? 0 : 1;
    s_book[tid].t0 = 0ull;
    s_book[tid].t_logic = s_book[tid].t_seated = s_book[tid].n_pass = 0;
  }
  if (tid == 0) s_drained = 0;
  __syncthreads();

  // every CTA of the cluster is running (and has its seat table initialised) before any DSMEM access
  cluster.sync();
  double* r_smem = cluster.map_shared_rank(smem, 0);

  // ---- the schedule ----
  // Seats are split in banks (seat g -> bank g & 1; one bank when the cluster is too small to
  // spare warps).  Iteration t SWEEPS the requests of bank t & 1, while the optimisers of the
  // OTHER bank take in their previous sweep's totals and advance their runs to the next
  // request.  One cluster barrier per iteration; a run advances one pass every two iterations,
  // and with two banks the optimiser turns (16-21 k cycles of one warp each) are hidden behind
  // the other bank's sweeps.
  //
  // Who does what in the leader CTA: when the slice layout reserves warps (a.reserved = 2, see
  // choose_geometry) warps 0..1 NEVER sweep and run the optimiser turns (two banks: warp j takes
  // seat lb + 2 j; one bank: warp j takes seats j, j + 2, ...), and the leader's slice is always
  // swept by warps 2..nw-1, so the partition of the points over threads -- and with it the
  // rounding of every sum -- does not depend on the number of seats or banks.  Without reserved
  // warps (small clusters, one bank) warp g takes seat g and every warp sweeps.
  const int n_banks = a.banks;  // 1 or 2
  const int reserved = crank == 0 ? a.reserved : 0;
  const bool may_logic = crank == 0 && (reserved > 0 ? warp < reserved : warp < G);
  const bool sweeper = warp >= reserved;
  const int swarp = warp - reserved, nsw = nw - reserved;
  int stid = swarp * 32 + lane, snt = nsw * 32;
  keep_in_register(stid);
  keep_in_register(snt);
  bool prev_empty = false;
  for (int t = 0;; ++t) {
    const int sb = t & 1;   // bank swept now
    const int lb = sb ^ 1;  // bank whose optimisers run now
    if (may_logic && (n_banks == 2 || lb == 0)) {
      // seats of bank lb this warp serves
      const int first = reserved > 0 ? (n_banks == 2 ? lb + 2 * warp : warp) : warp;
      const int step = reserved > 0 ? (n_banks == 2 ? 2 * reserved : reserved) : G;
      for (int g = first; g < G; g += step) {
        if (n_banks == 2 && (g & 1) != lb) continue;
        seat_turn<T, K>(a, cluster, g, lane, cs, *reinterpret_cast<FitState*>(VSR_SEAT_STATE(g)), VSR_SEAT_WS(g),
                        VSR_SEAT_CRED(g), s_ctrl[g], VSR_SEAT_CST(g), s_book[g], &s_drained);
      }
    }

    // ---- bank sb: seat table (this CTA's copy, published one iteration ago, stable now) ----
    int active = 0;
    bool empty = true;
    for (int g = n_banks == 2 ? sb : 0; g < G; g += n_banks) {
      if (n_banks == 1 && sb == 1) break;  // single bank: odd iterations only run the optimisers
      const SeatCtrl c = s_ctrl[g];
      if (c.prog >= 0) active |= 1 << g;
      if (c.prog >= 0 || !c.drained) empty = false;
    }

    if (sweeper && active) {
      for (int g = 0; g < G; ++g) {
        if (!((active >> g) & 1)) continue;
        const SeatCtrl c = s_ctrl[g];
        if (c.fresh) {
          // the run was seated one iteration ago: its program into this CTA's seat area, VAR
          // operands rewritten to slice columns, handler ids over the opcode bytes
          const int i0 = a.pt.insn_off[c.prog], ni = a.pt.insn_off[c.prog + 1] - i0;
          const int m0 = a.pt.imm_off[c.prog], nm = a.pt.imm_off[c.prog + 1] - m0;
          vsr_insn_t* s_insn = VSR_SEAT_INSN(g);
          double* s_imm = VSR_SEAT_IMM(g);
          for (int i = stid; i < ni; i += snt) {
            vsr_insn_t w = a.pt.insns[i0 + i];
            const unsigned op = VSR_OP(w);
            if (a.resident && op >= VSR_LOAD && op <= VSR_RPOW && op != VSR_PUSH && VSR_SRC(w) == VSR_SRC_VAR)
              w = (w & ~((vsr_insn_t)0xffff << 16)) | ((vsr_insn_t)a.col_of_var[VSR_IDX(w)] << 16);
            s_insn[i] = predecode(w);
          }
          if (stid == 0) s_insn[ni] = 0;  // pad word after END
          for (int i = stid; i < nm; i += snt) s_imm[i] = a.pt.imms[m0 + i];
        }
      }
      sweep_barrier(snt);
      for (int g = 0; g < G; ++g) {
        if (!((active >> g) & 1)) continue;
        // the seat's code area through ONE opaque byte offset: without the asm the compiler
        // re-derives the three pointers from (g, blockDim, kmax, ...) for every bytecode
        // instruction -- 17 of the 36 instructions of the dispatch sequence -- instead of
        // spending registers on them
        unsigned code_off = (unsigned)(reinterpret_cast<char*>(VSR_SEAT_INSN(g)) - reinterpret_cast<char*>(smem));
        asm volatile("" : "+r"(code_off));
        const vsr_insn_t* c_insn = reinterpret_cast<const vsr_insn_t*>(reinterpret_cast<char*>(smem) + code_off);
        const double* c_imm = reinterpret_cast<const double*>(c_insn) - VSR_MAX_IMMS;
        const T* c_cst = reinterpret_cast<const T*>(c_imm - kSeatCstDoubles);
        double* scratch = smem + (size_t)G * seat_d;
        if (a.resident)
          sweep_slice<T, K, P>(c_insn, c_imm, c_cst, xs, ys, stride, cnt, scratch, stid, snt);
        else
          sweep_points<T, K, P>(c_insn, c_imm, c_cst, X, y, a.pts.ldx, n0, n1, scratch, stid, snt);
        block_totals<K>(scratch, stid, snt,
                        r_smem + (size_t)g * seat_d + kFitStateDoubles + wsd + crank * (K + 1));
      }
    }
    cluster.sync();  // requests of bank lb and partial sums of bank sb are visible
    if (empty && prev_empty) break;  // both banks empty, queue drained: uniform over the cluster
    prev_empty = empty;
  }
#undef VSR_SEAT_STATE
#undef VSR_SEAT_WS
#undef VSR_SEAT_CRED
#undef VSR_SEAT_CST
#undef VSR_SEAT_IMM
#undef VSR_SEAT_INSN
  // no CTA may exit while another can still read its shared memory
  cluster.sync();
}

// dynamic shared memory of eval_kernel, in doubles: red[(K+1)*threads] | cst[k] | imm | insn
template <typename T, int K, int P>
__global__ void __launch_bounds__(256) eval_kernel(const EvalArgs a) {
  extern __shared__ double smem[];
  const int pair = blockIdx.x;
  const int split = blockIdx.y;
  if (pair >= a.n_pairs) return;
  const int prog = a.pair_prog[pair];
  const int k = a.pt.k[prog];
  const int nw = (blockDim.x + 31) >> 5;
  double* red = smem;  // reduction scratch of block_totals: [(K+1)][blockDim.x]
  const int n_red = (K + 1) * (int)blockDim.x;
  T* cst = reinterpret_cast<T*>(red + n_red);
  double* s_imm = red + n_red + k + 1;
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

  if (K == 0 && a.nan_to_num)  // scoring rule of the drivers, value only
    sweep_points<T, K, P, true>(s_insn, s_imm, cst, static_cast<const T*>(a.pts.X),
                                static_cast<const T*>(a.pts.y), a.pts.ldx, n0, n1, red, (int)threadIdx.x,
                                (int)blockDim.x);
  else
    sweep_points<T, K, P>(s_insn, s_imm, cst, static_cast<const T*>(a.pts.X),
                          static_cast<const T*>(a.pts.y), a.pts.ldx, n0, n1, red, (int)threadIdx.x,
                          (int)blockDim.x);
  block_totals<K>(red, (int)threadIdx.x, (int)blockDim.x,
                  a.partial + ((int64_t)pair * a.nsplit + split) * (K + 1));
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
