// vsr_bfgs.h -- one BFGS run as a resumable state machine.
//
// Restates what the reference gets from
//     scipy.optimize.minimize(safe_loss, x0, method='BFGS')
// (reference src/visymre/architectures/bfgs.py:115 and :179; scipy is an unpinned
// dependency of the reference, the algorithm below follows scipy 1.18.1):
//   _minimize_bfgs              scipy/optimize/_optimize.py:1345-1530
//   _line_search_wolfe12        scipy/optimize/_optimize.py:1156-1199
//   line_search_wolfe1 / scalar_search_wolfe1   scipy/optimize/_linesearch.py:37-190
//   DCSRCH / dcstep (MINPACK-2) scipy/optimize/_dcsrch.py
//   line_search_wolfe2 / scalar_search_wolfe2 / _zoom / _cubicmin / _quadmin
//                               scipy/optimize/_linesearch.py:192-680
//   ScalarFunction caching      scipy/optimize/_differentiable_functions.py:128-420
//   '2-point' forward difference scipy/optimize/_numdiff.py:580-700
//
// Why a state machine: on the GPU the optimiser of a run is ONE warp of its cluster's leader CTA
// (vsr_kernels.cuh: seat_turn).  It advances the run until the objective is needed at a new point,
// returns VSR_NEED_EVAL, the request goes out to the sweeping warps of the cluster, and the same
// warp resumes the run when the totals of the sweep have come back.  scipy's nested calls (BFGS ->
// line search -> phi/derphi -> ScalarFunction) are flattened into one protothread-style function;
// every variable that lives across an evaluation is a member of FitState -- which is also what
// travels when a run is handed over to another cluster.
//
// Execution model of fit_step on the device (a single GPU thread is ~30x slower than a
// CPU core on serial code, so the linear algebra must not be serial): all 32 lanes of the
// warp run the SAME control flow on identical scalar values.  The scalar state (FitState)
// lives in shared memory BETWEEN turns, one copy per run; a turn copies it into registers
// at entry (fit_step works on a private copy, so no lane ever does a read-modify-write on
// shared scalars and the compiler keeps them in registers across the vector stores) and
// writes it back when it yields.  The length-k vectors live in shared memory with element i
// handled by lane i % W (W = 8, 16 or 32 >= k: lanes beyond W repeat the work of the first W,
// so every reduction is a log2(W)-step butterfly that leaves the same value in all lanes):
// O(k) loops take one step, O(k^2) ones (H g, the BFGS update) k steps.  On the host `Lanes`
// degenerates to one lane and the loops are ordinary loops.
//
// The same header is compiled by g++ into oracle/hostsim (CPU tests check the logic
// against scipy itself); the product path only runs the CUDA build.
#ifndef VSR_BFGS_H_
#define VSR_BFGS_H_

#include "vsr_isa.h"

#if defined(__CUDACC__)
#define VSR_HDN __host__ __device__
#else
#include <cmath>
#define VSR_HDN
#endif

namespace vsr {

struct FitOpts {
  double gtol;        // 1e-5   (_optimize.py:1346)
  double c1;          // 1e-4
  double c2;          // 0.9
  double xrtol;       // 0
  double fd_eps;      // sqrt(eps) = 1.4901161193847656e-08 (_optimize.py:192)
  double penalty;     // 1e6    (bfgs.py:109)
  double loss_scale;  // 1 (MSE) or 1/mean(y) (NMSE, bfgs.py:85-90)
  double stop_time;   // seconds; the loss turns into `penalty` afterwards (bfgs.py:23-36)
  int maxiter_per_k;  // 200    (_optimize.py:1414)
  int grad_mode;      // VsrGradMode
};

enum { VSR_NEED_EVAL = 1, VSR_DONE = 0 };

// lane context of fit_step / fit_init.  W: lanes that own vector elements (k <= W)
template <int W>
struct LanesT {
#if defined(__CUDA_ARCH__)
  static __device__ __forceinline__ int first() { return (int)(threadIdx.x & (unsigned)(W - 1)); }
  static __device__ __forceinline__ int step() { return W; }
  static __device__ __forceinline__ void sync() { __syncwarp(); }
  // butterflies over the W lanes of a group: elements >= k are exact zeros, so the levels a
  // 32-lane butterfly would add on top contribute x + 0 = x: same value, fewer shuffles
  static __device__ __forceinline__ double sum(double v) {
#pragma unroll
    for (int m = W / 2; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
  }
  // three sums with their butterflies interleaved: the same additions in the same order as
  // three calls of sum(), one third of the latency (the optimiser warp is latency-bound)
  static __device__ __forceinline__ void sum3(double& a, double& b, double& c) {
#pragma unroll
    for (int m = W / 2; m > 0; m >>= 1) {
      const double ta = __shfl_xor_sync(0xffffffffu, a, m);
      const double tb = __shfl_xor_sync(0xffffffffu, b, m);
      const double tc = __shfl_xor_sync(0xffffffffu, c, m);
      a += ta;
      b += tb;
      c += tc;
    }
  }
  static __device__ __forceinline__ double maxv(double v) {
#pragma unroll
    for (int m = W / 2; m > 0; m >>= 1) {
      const double o = __shfl_xor_sync(0xffffffffu, v, m);
      v = o > v ? o : v;
    }
    return v;
  }
  static __device__ __forceinline__ bool any(bool b) { return __any_sync(0xffffffffu, b) != 0; }
  static __device__ __forceinline__ bool all(bool b) { return __all_sync(0xffffffffu, b) != 0; }
#else
  static int first() { return 0; }
  static int step() { return 1; }
  static void sync() {}
  static double sum(double v) { return v; }
  static void sum3(double&, double&, double&) {}
  static double maxv(double v) { return v; }
  static bool any(bool b) { return b; }
  static bool all(bool b) { return b; }
#endif
};
// element i of a length-k vector is handled by lane i (all of them by the host's one lane);
// `LN` is the LanesT<W> of the enclosing function
#if defined(__CUDA_ARCH__)
#define VSR_UNROLL4 _Pragma("unroll 4")
#define VSR_FI __forceinline__
#else
#define VSR_UNROLL4
#define VSR_FI inline
#endif
#if defined(__CUDA_ARCH__)
// k <= W on the device: at most ONE trip, spelled so that the compiler sees it (as a counted loop
// it unrolled every O(k) statement four times with a remainder loop)
#define VSR_FOR_K(i, k) for (int i = LN::first(); i < (k); i = (k))
#else
#define VSR_FOR_K(i, k) for (int i = LN::first(); i < (k); i += LN::step())
#endif

// DCSRCH task codes
enum { DC_START = 0, DC_FG = 1, DC_CONV = 2, DC_WARN = 3, DC_ERROR = 4 };

// Scalar state of a run plus ONE pointer to its vector workspace.  16-byte aligned and a
// multiple of 16 bytes so that a turn moves it between shared memory and registers in
// 128-bit pieces.
struct alignas(16) FitState {
  double* ws;  // [fit_workspace_doubles(k)], see the accessors below
  int pc;      // resume label of the protothread
  int k;       // number of constants
  // ---- request / response of one sweep over the points ----
  double rf;   // objective at xe() (scaled, penalty applied); its gradient goes to rg() (dual mode)
  // ---- ScalarFunction cache ----
  double cf;
  int f_ok, g_ok;
  int nfev, ngev;
  int fd_i;
  int have_gnew;
  double fd_dx;
  // ---- BFGS ----
  int it, maxiter, warnflag, status;
  double old_fval, old_old_fval, gnorm, alpha_k;
  // ---- line search (shared by wolfe1 and wolfe2) ----
  double phi0, derphi0, old_phi0;
  double stp, phi1, derphi1;
  int task, ls_i, ls_ok;
  // DCSRCH state (_dcsrch.py)
  int brackt, stage;
  // wolfe2 / zoom
  int zi, w2_i, zoom_ok, star_has_der, pad0_;
  double ls_fval, ls_oldfval;
  double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
  double alpha0, alpha1, phi_a0, phi_a1, derphi_a0, derphi_a1;
  double a_lo, a_hi, phi_lo, phi_hi, derphi_lo, a_rec, phi_rec, a_j, phi_aj, derphi_aj;
  double alpha_star, phi_star;
  // ---- the vectors, carved from ws (k doubles each, H k*k) ----
  VSR_HDN double* xe() const { return ws; }              // point the sweep must evaluate
  VSR_HDN double* rg() const { return ws + k; }          // gradient there (dual mode only)
  VSR_HDN double* cx() const { return ws + 2 * k; }      // current point of the ScalarFunction cache
  VSR_HDN double* cg() const { return ws + 3 * k; }      // gradient at cx
  VSR_HDN double* lastx() const { return ws + 4 * k; }   // last point evaluated (TimedFun.x, bfgs.py:35)
  VSR_HDN double* xk() const { return ws + 5 * k; }
  VSR_HDN double* gfk() const { return ws + 6 * k; }
  VSR_HDN double* pk() const { return ws + 7 * k; }
  VSR_HDN double* xt() const { return ws + 8 * k; }      // trial point xk + a*pk
  VSR_HDN double* gnew() const { return ws + 9 * k; }    // gradient returned by the line search
  VSR_HDN double* Hy() const { return ws + 10 * k; }     // scratch
  VSR_HDN double* H() const { return ws + 11 * k; }      // [k*k] inverse Hessian estimate
};

// number of doubles of workspace fit_init() carves for a run with k constants
VSR_HDN inline int fit_workspace_doubles(int k) { return 11 * k + k * k; }

template <int W = 32>
VSR_HDN inline void fit_init(FitState& S, int k, double* ws, const double* x0) {
  using LN = LanesT<W>;
  S.k = k;
  S.ws = ws;
  VSR_FOR_K(i, k) {
    ws[5 * k + i] = x0[i];  // xk
    ws[4 * k + i] = x0[i];  // lastx
  }
  S.pc = 0;
  S.f_ok = S.g_ok = 0;
  S.nfev = S.ngev = 0;
  S.it = 0;
  S.warnflag = 0;
  S.status = 0;
  S.rf = 0.0;
}

namespace detail {

VSR_HDN VSR_FI bool finite_d(double x) {
#if defined(__CUDA_ARCH__)
  return isfinite(x);
#else
  return std::isfinite(x);
#endif
}
VSR_HDN VSR_FI bool nan_d(double x) { return x != x; }
VSR_HDN VSR_FI double abs_d(double x) { return ::fabs(x); }
VSR_HDN VSR_FI double max_d(double a, double b) { return b > a ? b : a; }  // python max(a, b): a unless b > a
VSR_HDN VSR_FI double min_d(double a, double b) { return b < a ? b : a; }  // python min(a, b)
VSR_HDN VSR_FI double clip_d(double x, double lo, double hi) {
  // np.clip = minimum(maximum(x, lo), hi), nan propagates
  if (nan_d(x)) return x;
  double t = x < lo ? lo : x;
  return t > hi ? hi : t;
}
VSR_HDN VSR_FI double sign_d(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : (x == 0 ? 0.0 : x)); }

// dcstep (_dcsrch.py:502-728): safeguarded cubic/quadratic step of More-Thuente.
VSR_HDN VSR_FI void dcstep(double& stx, double& fx, double& dx, double& sty, double& fy,
                           double& dy, double& stp, double fp, double dp, int& brackt,
                           double stpmin, double stpmax) {
  const double sgnd = sign_d(dp) * sign_d(dx);
  double stpf, stpc, stpq, theta, s, gamma, p, q, r;
  if (fp > fx) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = max_d(max_d(abs_d(theta), abs_d(dx)), abs_d(dp));
    gamma = s * ::sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    p = (gamma - dx) + theta;
    q = ((gamma - dx) + gamma) + dp;
    r = p / q;
    stpc = stx + r * (stp - stx);
    stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    if (abs_d(stpc - stx) <= abs_d(stpq - stx))
      stpf = stpc;
    else
      stpf = stpc + (stpq - stpc) / 2.0;
    brackt = 1;
  } else if (sgnd < 0.0) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = max_d(max_d(abs_d(theta), abs_d(dx)), abs_d(dp));
    gamma = s * ::sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = ((gamma - dp) + gamma) + dx;
    r = p / q;
    stpc = stp + r * (stx - stp);
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (abs_d(stpc - stp) > abs_d(stpq - stp))
      stpf = stpc;
    else
      stpf = stpq;
    brackt = 1;
  } else if (abs_d(dp) < abs_d(dx)) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = max_d(max_d(abs_d(theta), abs_d(dx)), abs_d(dp));
    double rad = (theta / s) * (theta / s) - (dx / s) * (dp / s);
    rad = max_d(0.0, rad);  // python max(0, nan) == 0
    gamma = s * ::sqrt(rad);
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = (gamma + (dx - dp)) + gamma;
    r = p / q;
    if (r < 0 && gamma != 0)
      stpc = stp + r * (stx - stp);
    else if (stp > stx)
      stpc = stpmax;
    else
      stpc = stpmin;
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      if (abs_d(stpc - stp) < abs_d(stpq - stp))
        stpf = stpc;
      else
        stpf = stpq;
      if (stp > stx)
        stpf = min_d(stp + 0.66 * (sty - stp), stpf);
      else
        stpf = max_d(stp + 0.66 * (sty - stp), stpf);
    } else {
      if (abs_d(stpc - stp) > abs_d(stpq - stp))
        stpf = stpc;
      else
        stpf = stpq;
      stpf = clip_d(stpf, stpmin, stpmax);
    }
  } else {
    if (brackt) {
      theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      s = max_d(max_d(abs_d(theta), abs_d(dy)), abs_d(dp));
      gamma = s * ::sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      p = (gamma - dp) + theta;
      q = ((gamma - dp) + gamma) + dy;
      r = p / q;
      stpc = stp + r * (sty - stp);
      stpf = stpc;
    } else if (stp > stx) {
      stpf = stpmax;
    } else {
      stpf = stpmin;
    }
  }
  if (fp > fx) {
    sty = stp;
    fy = fp;
    dy = dp;
  } else {
    if (sgnd < 0) {
      sty = stx;
      fy = fx;
      dy = dx;
    }
    stx = stp;
    fx = fp;
    dx = dp;
  }
  stp = stpf;
}

// DCSRCH._iterate (_dcsrch.py:244-500).  ftol = c1, gtol = c2.
VSR_HDN VSR_FI void dcsrch_iterate(FitState& S, double& stp, double f, double g, int& task,
                                   double ftol, double gtol, double xtol, double stpmin,
                                   double stpmax) {
  const double p5 = 0.5, p66 = 0.66, xtrapl = 1.1, xtrapu = 4.0;
  if (task == DC_START) {
    if (stp < stpmin) task = DC_ERROR;
    if (stp > stpmax) task = DC_ERROR;
    if (g >= 0) task = DC_ERROR;
    if (task == DC_ERROR) return;
    S.brackt = 0;
    S.stage = 1;
    S.finit = f;
    S.ginit = g;
    S.gtest = ftol * S.ginit;
    S.width = stpmax - stpmin;
    S.width1 = S.width / p5;
    S.stx = 0.0;
    S.fx = S.finit;
    S.gx = S.ginit;
    S.sty = 0.0;
    S.fy = S.finit;
    S.gy = S.ginit;
    S.stmin = 0.0;
    S.stmax = stp + xtrapu * stp;
    task = DC_FG;
    return;
  }
  const double ftest = S.finit + stp * S.gtest;
  if (S.stage == 1 && f <= ftest && g >= 0) S.stage = 2;
  // warnings, then convergence (a later test overrides an earlier one)
  if (S.brackt && (stp <= S.stmin || stp >= S.stmax)) task = DC_WARN;
  if (S.brackt && S.stmax - S.stmin <= xtol * S.stmax) task = DC_WARN;
  if (stp == stpmax && f <= ftest && g <= S.gtest) task = DC_WARN;
  if (stp == stpmin && (f > ftest || g >= S.gtest)) task = DC_WARN;
  if (f <= ftest && abs_d(g) <= gtol * -S.ginit) task = DC_CONV;
  if (task == DC_WARN || task == DC_CONV) return;

  {
    // stage 1 works on the modified function psi (f - stp*gtest); one dcstep call site for both
    // stages (the routine is inlined: two sites doubled the optimiser's code)
    const bool mod = S.stage == 1 && f <= S.fx && f > ftest;
    double fm = f, gm = g, fxm = S.fx, fym = S.fy, gxm = S.gx, gym = S.gy;
    if (mod) {
      fm = f - stp * S.gtest;
      fxm = S.fx - S.stx * S.gtest;
      fym = S.fy - S.sty * S.gtest;
      gm = g - S.gtest;
      gxm = S.gx - S.gtest;
      gym = S.gy - S.gtest;
    }
    dcstep(S.stx, fxm, gxm, S.sty, fym, gym, stp, fm, gm, S.brackt, S.stmin, S.stmax);
    if (mod) {
      S.fx = fxm + S.stx * S.gtest;
      S.fy = fym + S.sty * S.gtest;
      S.gx = gxm + S.gtest;
      S.gy = gym + S.gtest;
    } else {
      S.fx = fxm;
      S.fy = fym;
      S.gx = gxm;
      S.gy = gym;
    }
  }
  if (S.brackt) {
    if (abs_d(S.sty - S.stx) >= p66 * S.width1) stp = S.stx + p5 * (S.sty - S.stx);
    S.width1 = S.width;
    S.width = abs_d(S.sty - S.stx);
  }
  if (S.brackt) {
    S.stmin = min_d(S.stx, S.sty);
    S.stmax = max_d(S.stx, S.sty);
  } else {
    S.stmin = stp + xtrapl * (stp - S.stx);
    S.stmax = stp + xtrapu * (stp - S.stx);
  }
  stp = clip_d(stp, stpmin, stpmax);
  if ((S.brackt && (stp <= S.stmin || stp >= S.stmax)) ||
      (S.brackt && S.stmax - S.stmin <= xtol * S.stmax))
    stp = S.stx;
  task = DC_FG;
}

// _cubicmin / _quadmin (_linesearch.py): a non-finite result means "None".
VSR_HDN VSR_FI bool cubicmin(double a, double fa, double fpa, double b, double fb, double c,
                             double fc, double& xmin) {
  const double C = fpa;
  const double db = b - a, dc = c - a;
  const double denom = (db * dc) * (db * dc) * (db - dc);
  const double r0 = fb - fa - C * db, r1 = fc - fa - C * dc;
  double A = (dc * dc) * r0 + (-(db * db)) * r1;
  double B = (-(dc * dc * dc)) * r0 + (db * db * db) * r1;
  if (denom == 0.0) return false;
  A /= denom;
  B /= denom;
  const double radical = B * B - 3 * A * C;
  if (!(radical >= 0.0) || A == 0.0) return false;
  xmin = a + (-B + ::sqrt(radical)) / (3 * A);
  return finite_d(xmin);
}
VSR_HDN VSR_FI bool quadmin(double a, double fa, double fpa, double b, double fb, double& xmin) {
  const double D = fa, C = fpa;
  const double db = b - a * 1.0;
  if (db * db == 0.0) return false;
  const double B = (fb - D - C * db) / (db * db);
  if (B == 0.0 || !finite_d(B)) return false;
  xmin = a - C / (2.0 * B);
  return finite_d(xmin);
}

}  // namespace detail

// ---- protothread plumbing -----------------------------------------------------------
#define VSR_CO_YIELD_(S, n)   \
  do {                        \
    (S).pc = (n) + 1;         \
    goto vsr_yield_;          \
    case (n) + 1:;            \
  } while (0)
#define VSR_CO_YIELD(S) VSR_CO_YIELD_(S, __COUNTER__)

// ScalarFunction._update_x via fun()/grad(): np.array_equal(x, self.x)
#define VSR_OBJ_SETX(S, xptr)                                      \
  do {                                                             \
    bool same_ = true;                                             \
    VSR_FOR_K(i_, (S).k)                                           \
    if (!((xptr)[i_] == v_cx[i_])) same_ = false;                \
    same_ = LN::all(same_);                                     \
    if (!same_) {                                                  \
      VSR_FOR_K(i_, (S).k)v_cx[i_] = (xptr)[i_];                 \
      (S).f_ok = 0;                                                \
      (S).g_ok = 0;                                                \
    }                                                              \
  } while (0)

// ScalarFunction._update_fun.  In dual mode one sweep returns f and its gradient.
#define VSR_OBJ_UPDATE_FUN(S, O)                                   \
  do {                                                             \
    if (!(S).f_ok) {                                               \
      VSR_FOR_K(i_, (S).k)v_xe[i_] = v_cx[i_];                 \
      VSR_CO_YIELD(S);                                             \
      (S).cf = (S).rf;                                             \
      (S).nfev += 1;                                               \
      (S).f_ok = 1;                                                \
      VSR_FOR_K(i_, (S).k)v_lastx[i_] = v_xe[i_];              \
      if ((O).grad_mode == VSR_GRAD_DUAL) {                        \
        VSR_FOR_K(i_, (S).k)v_cg[i_] = v_rg[i_];               \
        (S).g_ok = 1;                                              \
        (S).ngev += 1;                                             \
      }                                                            \
    }                                                              \
  } while (0)

// ScalarFunction._update_grad; FD branch = approx_derivative('2-point', abs_step=eps)
#define VSR_OBJ_UPDATE_GRAD(S, O)                                                    \
  do {                                                                               \
    if (!(S).g_ok) {                                                                 \
      VSR_OBJ_UPDATE_FUN(S, O);                                                      \
      if (!(S).g_ok) {                                                               \
        for ((S).fd_i = 0; (S).fd_i < (S).k; ++(S).fd_i) {                           \
          {                                                                          \
            LN::sync(); /* cx[fd_i] was written by its owner lane */              \
            const double x_ = v_cx[(S).fd_i];                                      \
            double h_ = (O).fd_eps;                                                  \
            if ((x_ + h_) - x_ == 0.0) {                                             \
              const double a_ = ::fabs(x_) > 1.0 ? ::fabs(x_) : 1.0;                 \
              h_ = 1.4901161193847656e-08 * (x_ >= 0 ? 1.0 : -1.0) * a_;             \
            }                                                                        \
            (S).fd_dx = (x_ + h_) - x_;                                              \
            VSR_FOR_K(i_, (S).k)v_xe[i_] = (i_ == (S).fd_i) ? x_ + h_ : v_cx[i_]; \
          }                                                                          \
          VSR_CO_YIELD(S);                                                           \
          VSR_FOR_K(i_, (S).k) {                                                     \
            if (i_ == (S).fd_i) v_cg[i_] = ((S).rf - (S).cf) / (S).fd_dx;          \
            v_lastx[i_] = v_xe[i_];                                              \
          }                                                                          \
          (S).nfev += 1;                                                             \
        }                                                                            \
        (S).g_ok = 1;                                                                \
        (S).ngev += 1;                                                               \
      }                                                                              \
    }                                                                                \
  } while (0)

// phi(a) = f(xk + a*pk) ; derphi(a) = grad(xk + a*pk) . pk   (closures of
// line_search_wolfe1/2).  The trial point is rebuilt with the same arithmetic scipy
// uses so the ScalarFunction cache hits exactly when scipy's does.
#define VSR_LS_PHI(S, O, a, out)                                                  \
  do {                                                                            \
    VSR_FOR_K(i_, (S).k)v_xt[i_] = v_xk[i_] + (a) * v_pk[i_];               \
    VSR_OBJ_SETX(S, v_xt);                                                      \
    VSR_OBJ_UPDATE_FUN(S, O);                                                     \
    (out) = (S).cf;                                                               \
  } while (0)
#define VSR_LS_DERPHI(S, O, a, out)                                               \
  do {                                                                            \
    VSR_FOR_K(i_, (S).k)v_xt[i_] = v_xk[i_] + (a) * v_pk[i_];               \
    VSR_OBJ_SETX(S, v_xt);                                                      \
    VSR_OBJ_UPDATE_GRAD(S, O);                                                    \
    {                                                                             \
      double d_ = 0.0;                                                            \
      VSR_FOR_K(i_, (S).k) {                                                      \
        v_gnew[i_] = v_cg[i_];                                                \
        d_ += v_cg[i_] * v_pk[i_];                                            \
      }                                                                           \
      (S).have_gnew = 1;                                                          \
      (out) = LN::sum(d_);                                                     \
    }                                                                             \
  } while (0)

// Advance the run.  Returns VSR_NEED_EVAL when the caller must evaluate the objective
// at S.xe() (store it in S.rf, and its gradient in S.rg() in dual mode), VSR_DONE when
// finished (result: S.xk(), S.old_fval, S.status, S.it, S.nfev, S.lastx()).
template <int W = 32>
VSR_HDN VSR_FI int fit_step(FitState& G_, const FitOpts& O) {
  using namespace detail;
  using LN = LanesT<W>;
#if defined(__CUDA_ARCH__)
  // the state and the workspace live in shared memory: LDS/STS instead of generic loads and stores
  __builtin_assume(__isShared(&G_));
#endif
  FitState S = G_;  // private copy: registers (and an L1-resident stack) for the length of the turn
  const int k = S.k;
  double* const v_xe = S.ws;
#if defined(__CUDA_ARCH__)
  __builtin_assume(__isShared(v_xe));
#endif
  double* const v_rg = S.ws + k;
  double* const v_cx = S.ws + 2 * k;
  double* const v_cg = S.ws + 3 * k;
  double* const v_lastx = S.ws + 4 * k;
  double* const v_xk = S.ws + 5 * k;
  double* const v_gfk = S.ws + 6 * k;
  double* const v_pk = S.ws + 7 * k;
  double* const v_xt = S.ws + 8 * k;
  double* const v_gnew = S.ws + 9 * k;
  double* const v_Hy = S.ws + 10 * k;
  double* const v_H = S.ws + 11 * k;
  switch (S.pc) {
    case 0:
      // ScalarFunction.__init__: f and grad at x0
      VSR_FOR_K(i, k) v_cx[i] = v_xk[i];
      S.f_ok = S.g_ok = 0;
      VSR_OBJ_UPDATE_FUN(S, O);
      VSR_OBJ_UPDATE_GRAD(S, O);
      S.old_fval = S.cf;
      S.it = 0;
      S.maxiter = k * O.maxiter_per_k;
      {
        double n2 = 0.0, gm = 0.0;
        bool gnan = false;
        VSR_FOR_K(i, k) {
          const double gi = v_cg[i];
          v_gfk[i] = gi;
          for (int j = 0; j < k; ++j) v_H[i * k + j] = (i == j) ? 1.0 : 0.0;
          n2 += gi * gi;
          if (nan_d(gi)) gnan = true;
          if (abs_d(gi) > gm) gm = abs_d(gi);
        }
        n2 = LN::sum(n2);
        gm = LN::maxv(gm);
        gnan = LN::any(gnan);
        S.old_old_fval = S.old_fval + ::sqrt(n2) / 2;
        S.gnorm = gnan ? ::nan("") : gm;  // vecnorm(gfk, inf) = amax(|gfk|), nan propagates
      }
      S.warnflag = 0;

      while (S.gnorm > O.gtol && S.it < S.maxiter) {
        // pk = -Hk . gfk   (row i on lane i)
        LN::sync();
        {
          double d = 0.0;
          VSR_FOR_K(i, k) {
            double acc = 0.0;
            VSR_UNROLL4
            for (int j = 0; j < k; ++j) acc += v_H[i * k + j] * v_gfk[j];
            v_pk[i] = -acc;
            d += v_gfk[i] * -acc;
          }
          S.derphi0 = LN::sum(d);
        }
        // ---------------- line_search_wolfe1 ----------------
        S.phi0 = S.old_fval;
        S.old_phi0 = S.old_old_fval;
        S.have_gnew = 0;  // gval = [gfk]
        if (S.derphi0 != 0) {
          S.alpha1 = min_d(1.0, 1.01 * 2 * (S.phi0 - S.old_phi0) / S.derphi0);
          if (S.alpha1 < 0) S.alpha1 = 1.0;
        } else {
          S.alpha1 = 1.0;
        }
        S.phi1 = S.phi0;
        S.derphi1 = S.derphi0;
        S.task = DC_START;
        S.ls_ok = 0;
        S.stp = S.alpha1;
        for (S.ls_i = 0; S.ls_i < 100; ++S.ls_i) {
          S.stp = S.alpha1;
          dcsrch_iterate(S, S.stp, S.phi1, S.derphi1, S.task, O.c1, O.c2, 1e-14, 1e-100, 1e100);
          if (!finite_d(S.stp)) {
            S.task = DC_WARN;
            break;
          }
          if (S.task == DC_FG) {
            S.alpha1 = S.stp;
            VSR_LS_PHI(S, O, S.stp, S.phi1);
            VSR_LS_DERPHI(S, O, S.stp, S.derphi1);
          } else {
            break;
          }
        }
        if (S.ls_i >= 100) S.task = DC_WARN;  // for-else: did not converge
        S.ls_ok = (S.task == DC_CONV);
        if (S.ls_ok) {
          S.alpha_k = S.stp;
          S.ls_fval = S.phi1;
          S.ls_oldfval = S.phi0;
          // gval[0]: gradient of the last derphi call (gfk when there was none)
          if (!S.have_gnew) VSR_FOR_K(i, k) v_gnew[i] = v_gfk[i];
          S.have_gnew = 1;
        } else {
          // ---------------- line_search_wolfe2 (fallback) ----------------
          S.have_gnew = 0;
          S.alpha0 = 0.0;
          if (S.derphi0 != 0)
            S.alpha1 = min_d(1.0, 1.01 * 2 * (S.phi0 - S.old_phi0) / S.derphi0);
          else
            S.alpha1 = 1.0;
          if (S.alpha1 < 0) S.alpha1 = 1.0;
          S.alpha1 = min_d(S.alpha1, 1e100);
          VSR_LS_PHI(S, O, S.alpha1, S.phi_a1);
          S.phi_a0 = S.phi0;
          S.derphi_a0 = S.derphi0;
          S.ls_ok = 0;        // alpha_star is not None
          S.star_has_der = 0; // derphi_star is not None
          S.zoom_ok = -1;     // -1: no zoom requested, 0/1: zoom arguments ready
          for (S.w2_i = 0; S.w2_i < 10; ++S.w2_i) {
            if (S.alpha1 == 0 || S.alpha0 > 1e100) {
              S.ls_ok = 0;
              S.star_has_der = 0;
              break;
            }
            if ((S.phi_a1 > S.phi0 + O.c1 * S.alpha1 * S.derphi0) ||
                ((S.phi_a1 >= S.phi_a0) && S.w2_i > 0)) {
              S.a_lo = S.alpha0;
              S.a_hi = S.alpha1;
              S.phi_lo = S.phi_a0;
              S.phi_hi = S.phi_a1;
              S.derphi_lo = S.derphi_a0;
              S.zoom_ok = 0;
              break;
            }
            VSR_LS_DERPHI(S, O, S.alpha1, S.derphi_a1);
            if (abs_d(S.derphi_a1) <= -O.c2 * S.derphi0) {
              S.alpha_star = S.alpha1;
              S.phi_star = S.phi_a1;
              S.ls_ok = 1;
              S.star_has_der = 1;
              break;
            }
            if (S.derphi_a1 >= 0) {
              S.a_lo = S.alpha1;
              S.a_hi = S.alpha0;
              S.phi_lo = S.phi_a1;
              S.phi_hi = S.phi_a0;
              S.derphi_lo = S.derphi_a1;
              S.zoom_ok = 0;
              break;
            }
            {
              const double alpha2 = min_d(2 * S.alpha1, 1e100);
              S.alpha0 = S.alpha1;
              S.alpha1 = alpha2;
            }
            S.phi_a0 = S.phi_a1;
            VSR_LS_PHI(S, O, S.alpha1, S.phi_a1);
            S.derphi_a0 = S.derphi_a1;
          }
          if (S.w2_i >= 10 && S.zoom_ok < 0 && !S.ls_ok) {
            // for-else: maxiter reached; alpha_star = alpha1, derphi_star = None
            S.alpha_star = S.alpha1;
            S.phi_star = S.phi_a1;
            S.ls_ok = 1;
            S.star_has_der = 0;
          }
          if (S.zoom_ok == 0) {
            // ---------------- _zoom ----------------
            S.zi = 0;
            S.phi_rec = S.phi0;
            S.a_rec = 0.0;
            for (;;) {
              {
                const double dalpha = S.a_hi - S.a_lo;
                double a, b;
                if (dalpha < 0) {
                  a = S.a_hi;
                  b = S.a_lo;
                } else {
                  a = S.a_lo;
                  b = S.a_hi;
                }
                bool have = false;
                double aj = 0.0;
                if (S.zi > 0) {
                  const double cchk = 0.2 * dalpha;
                  have = cubicmin(S.a_lo, S.phi_lo, S.derphi_lo, S.a_hi, S.phi_hi, S.a_rec,
                                  S.phi_rec, aj);
                  if (have && ((aj > b - cchk) || (aj < a + cchk))) have = false;
                }
                if (!have) {
                  const double qchk = 0.1 * dalpha;
                  have = quadmin(S.a_lo, S.phi_lo, S.derphi_lo, S.a_hi, S.phi_hi, aj);
                  if (!have || (aj > b - qchk) || (aj < a + qchk)) aj = S.a_lo + 0.5 * dalpha;
                }
                S.a_j = aj;
              }
              VSR_LS_PHI(S, O, S.a_j, S.phi_aj);
              if ((S.phi_aj > S.phi0 + O.c1 * S.a_j * S.derphi0) || (S.phi_aj >= S.phi_lo)) {
                S.phi_rec = S.phi_hi;
                S.a_rec = S.a_hi;
                S.a_hi = S.a_j;
                S.phi_hi = S.phi_aj;
              } else {
                VSR_LS_DERPHI(S, O, S.a_j, S.derphi_aj);
                if (abs_d(S.derphi_aj) <= -O.c2 * S.derphi0) {
                  S.alpha_star = S.a_j;
                  S.phi_star = S.phi_aj;
                  S.ls_ok = 1;
                  S.star_has_der = 1;
                  break;
                }
                if (S.derphi_aj * (S.a_hi - S.a_lo) >= 0) {
                  S.phi_rec = S.phi_hi;
                  S.a_rec = S.a_hi;
                  S.a_hi = S.a_lo;
                  S.phi_hi = S.phi_lo;
                } else {
                  S.phi_rec = S.phi_lo;
                  S.a_rec = S.a_lo;
                }
                S.a_lo = S.a_j;
                S.phi_lo = S.phi_aj;
                S.derphi_lo = S.derphi_aj;
              }
              S.zi += 1;
              if (S.zi > 10) {
                S.ls_ok = 0;
                S.star_has_der = 0;
                break;
              }
            }
          }
          if (S.ls_ok) {
            S.alpha_k = S.alpha_star;
            S.ls_fval = S.phi_star;
            S.ls_oldfval = S.phi0;
            // derphi_star None -> gfkp1 None (recomputed below); else gval[0]
            if (!S.star_has_der) S.have_gnew = 0;
          }
        }
        if (!S.ls_ok) {
          S.warnflag = 2;  // _LineSearchError
          break;
        }
        S.old_fval = S.ls_fval;
        S.old_old_fval = S.ls_oldfval;
#ifdef VSR_TRACE
        VSR_TRACE(S);
#endif
        // xkp1 = xk + alpha_k*pk
        VSR_FOR_K(i, k) {
          v_Hy[i] = S.alpha_k * v_pk[i];  // sk (kept in Hy until the update below)
          v_xt[i] = v_xk[i] + v_Hy[i];
        }
        if (!S.have_gnew) {
          VSR_OBJ_SETX(S, v_xt);
          VSR_OBJ_UPDATE_GRAD(S, O);
          VSR_FOR_K(i, k) v_gnew[i] = v_cg[i];
        }
        {
          double gm = 0.0, pn2 = 0.0, xn2 = 0.0, ys = 0.0;
          bool gnan = false;
          VSR_FOR_K(i, k) {
            pn2 += v_pk[i] * v_pk[i];
            const double yi = v_gnew[i] - v_gfk[i];
            v_pk[i] = yi;  // yk (pk is free until the next iteration)
            v_gfk[i] = v_gnew[i];
            v_xk[i] = v_xt[i];
            xn2 += v_xk[i] * v_xk[i];
            ys += yi * v_Hy[i];
            if (nan_d(v_gfk[i])) gnan = true;
            if (abs_d(v_gfk[i]) > gm) gm = abs_d(v_gfk[i]);
          }
          LN::sum3(pn2, xn2, ys);
          const double rhok_inv = ys;
          gm = LN::maxv(gm);
          gnan = LN::any(gnan);
          S.it += 1;
          S.gnorm = gnan ? ::nan("") : gm;
          if (S.gnorm <= O.gtol) break;
          if (S.alpha_k * ::sqrt(pn2) <= O.xrtol * (O.xrtol + ::sqrt(xn2))) break;
          if (!finite_d(S.old_fval)) {
            S.warnflag = 2;
            break;
          }
          // BFGS update of the inverse Hessian (_optimize.py:1496-1499)
          //   H <- A1 (H A2) + rho s s^T,  A1 = I - rho s y^T,  A2 = I - rho y s^T
          // evaluated as the two products scipy forms, using their rank-one structure:
          //   T = H A2      = H - rho (H y) s^T        (row i on lane i)
          //   A1 T          = T - rho s (y^T T)        (column sums on lane j, rows on lane i)
          const double* sk = v_Hy;
          const double* yk = v_pk;
          const double rhok = (rhok_inv == 0.0) ? 1000.0 : 1.0 / rhok_inv;
          double* w = v_xt;  // free until the next line search
          LN::sync();     // yk, sk complete
          VSR_FOR_K(i, k) {
            double ui = 0.0;
            VSR_UNROLL4
            for (int j = 0; j < k; ++j) ui += v_H[i * k + j] * yk[j];
            VSR_UNROLL4
            for (int j = 0; j < k; ++j) v_H[i * k + j] -= rhok * ui * sk[j];
          }
          LN::sync();  // T complete
          VSR_FOR_K(j, k) {
            double a2 = 0.0;
            VSR_UNROLL4
            for (int l = 0; l < k; ++l) a2 += yk[l] * v_H[l * k + j];
            w[j] = a2;
          }
          LN::sync();  // w complete
          VSR_FOR_K(i, k) {
            VSR_UNROLL4
            for (int j = 0; j < k; ++j) v_H[i * k + j] += rhok * sk[i] * sk[j] - rhok * sk[i] * w[j];
          }
        }
      }
      // ---- termination message (_optimize.py:1503-1513) ----
      {
        bool xnan = false;
        VSR_FOR_K(i, k) if (nan_d(v_xk[i])) xnan = true;
        xnan = LN::any(xnan);
        if (S.warnflag == 2)
          S.status = VSR_FIT_PRECLOSS;
        else if (S.it >= S.maxiter)
          S.status = VSR_FIT_MAXITER;
        else if (nan_d(S.gnorm) || nan_d(S.old_fval) || xnan)
          S.status = VSR_FIT_NAN;
        else
          S.status = VSR_FIT_SUCCESS;
      }
      S.pc = -1;
      G_ = S;
      return VSR_DONE;
    default:
      return VSR_DONE;
  }
vsr_yield_:  // the one exit of every VSR_CO_YIELD: the private copy goes back to shared memory
  G_ = S;
  return VSR_NEED_EVAL;
}

}  // namespace vsr

#endif  // VSR_BFGS_H_
