// vsr_kernels.cuh -- sm_100a kernels of the refinement engine.
//
//   fit_kernel<T,K,P>   one CTA per (candidate, restart) run: thread 0 advances the BFGS
//                       state machine (vsr_bfgs.h), all threads sweep the points through
//                       the interpreter (vsr_interp.h) whenever it asks for the objective.
//                       Replaces minimize(safe_loss, x0, 'BFGS') + the lambdified loss
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
  if (threadIdx.x == 0) {
    for (int w = 1; w < nw; ++w) {
      s += red[w * (K + 1)];
#pragma unroll
      for (int t = 0; t < K; ++t) g[t] += red[w * (K + 1) + 1 + t];
    }
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

// dynamic shared memory of fit_kernel, in doubles:
//   ws[fit_workspace_doubles(k)] | red[nwarps*(K+1)] | cst[kmax] | imm[n_imm] | insn[n_insn]
__host__ __device__ inline int fit_smem_doubles(int kmax, int K, int nwarps, int n_insn,
                                                int n_imm) {
  return fit_workspace_doubles(kmax) + nwarps * (K + 1) + kmax + n_imm + n_insn + 2;
}

template <typename T, int K, int P>
__global__ void __launch_bounds__(256) fit_kernel(const FitArgs a) {
  extern __shared__ double smem[];
  __shared__ FitState S;
  __shared__ int s_action;
  __shared__ unsigned long long s_t0;

  const int run = blockIdx.x;
  if (run >= a.n_runs) return;
  const int prog = a.run_prog[run];
  const int slot = a.run_slot[run];
  const int k = a.pt.k[prog];
  const int nw = (blockDim.x + 31) >> 5;

  double* ws = smem;
  double* red = ws + fit_workspace_doubles(k);
  T* cst = reinterpret_cast<T*>(red + nw * (K + 1));
  double* s_imm = red + nw * (K + 1) + k + 1;
  int n_insn, n_imm;
  const int m_imm = a.pt.imm_off[prog + 1] - a.pt.imm_off[prog];
  vsr_insn_t* s_insn = reinterpret_cast<vsr_insn_t*>(s_imm + m_imm);
  load_program(a.pt, prog, s_insn, s_imm, n_insn, n_imm);

  int32_t* info = a.out_info + (int64_t)slot * 4;
  if (k == 0) {  // nothing to optimise (reference bfgs.py:117-118)
    if (threadIdx.x == 0) {
      info[0] = VSR_FIT_NOT_RUN;
      info[1] = 0;
      info[2] = 0;
      info[3] = 0;
      a.out_loss[slot] = 0.0;
    }
    return;
  }
  if (threadIdx.x == 0) {
    fit_init(S, k, ws, a.x0 + (int64_t)slot * a.kstride);
    s_t0 = 0ull;
  }
  __syncthreads();

  const T* X = static_cast<const T*>(a.pts.X);
  const T* y = static_cast<const T*>(a.pts.y);
  const int64_t N = a.pts.n;
  const double inv_n = 1.0 / (double)N;

  for (;;) {
    if (threadIdx.x == 0) s_action = fit_step(S, a.O);
    __syncthreads();
    if (s_action == VSR_DONE) break;
    // constants of this evaluation, in the arithmetic type of the sweep
    for (int i = threadIdx.x; i < k; i += blockDim.x) cst[i] = (T)S.xe[i];
    __syncthreads();
    double s, g[K > 0 ? K : 1];
    sweep_points<T, K, P>(s_insn, s_imm, cst, X, y, a.pts.ldx, 0, N, s, g);
    block_sum<K>(s, g, red);
    if (threadIdx.x == 0) {
      double f = a.O.loss_scale * (s * inv_n);
      bool bad = !isfinite(f);
      if (a.O.stop_time < 1e8) {  // TimedFun (bfgs.py:29-33): the clock starts at the first call
        const unsigned long long now = global_ns();
        if (s_t0 == 0ull)
          s_t0 = now;
        else if ((double)(now - s_t0) * 1e-9 >= a.O.stop_time)
          bad = true;
      }
      if (bad) {
        S.rf = a.O.penalty;
#pragma unroll
        for (int t = 0; t < K; ++t)
          if (t < k) S.rg[t] = 0.0;
      } else {
        S.rf = f;
#pragma unroll
        for (int t = 0; t < K; ++t)
          if (t < k) {
            const double gv = a.O.loss_scale * (2.0 * g[t] * inv_n);
            S.rg[t] = isfinite(gv) ? gv : 0.0;
          }
      }
    }
    // thread 0 goes straight back into fit_step; the others wait at the barrier above
  }

  if (threadIdx.x == 0) {
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

}  // namespace vsr

#endif  // VSR_KERNELS_CUH_
