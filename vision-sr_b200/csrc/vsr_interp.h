// vsr_interp.h -- the skeleton interpreter: evaluates one compiled skeleton at P
// points at once as forward-mode dual numbers with K tangents.
//
// Replaces the lambdified scalar loss of the reference
// (src/visymre/architectures/bfgs.py:104-112: one Python term per data point) and
// the per-restart lambdify over variables (bfgs.py:120-132).  Numeric semantics are
// numpy's (bfgs.py:38-40): a domain violation gives nan, overflow gives inf; the
// caller turns a non-finite loss into the 1e6 penalty.
//
// The same code is compiled by nvcc into the kernels and by g++ into the host
// simulator the CPU tests use to check the ISA semantics (oracle/hostsim); the
// product path only ever runs the CUDA build.
//
// Template parameters
//   T  float or double (arithmetic type of the point evaluation)
//   K  number of tangents carried (0 = value only)
//   P  points evaluated per call; the decode of each instruction is shared by the P
//      points, which amortises the interpreter overhead and gives the FP pipes P
//      independent dependency chains.
#ifndef VSR_INTERP_H_
#define VSR_INTERP_H_

#include "vsr_isa.h"

#if defined(__CUDACC__)
#define VSR_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define VSR_HD inline
#endif

namespace vsr {

// ---- scalar math, overloaded on float/double ------------------------------------
#define VSR_M1(name, fd, ff)                         \
  VSR_HD double name(double x) { return fd(x); }     \
  VSR_HD float name(float x) { return ff(x); }
VSR_M1(m_sqrt, ::sqrt, ::sqrtf)
VSR_M1(m_exp, ::exp, ::expf)
VSR_M1(m_log, ::log, ::logf)
VSR_M1(m_sin, ::sin, ::sinf)
VSR_M1(m_cos, ::cos, ::cosf)
VSR_M1(m_tan, ::tan, ::tanf)
VSR_M1(m_asin, ::asin, ::asinf)
VSR_M1(m_acos, ::acos, ::acosf)
VSR_M1(m_atan, ::atan, ::atanf)
VSR_M1(m_sinh, ::sinh, ::sinhf)
VSR_M1(m_cosh, ::cosh, ::coshf)
VSR_M1(m_tanh, ::tanh, ::tanhf)
VSR_M1(m_abs, ::fabs, ::fabsf)
#undef VSR_M1
VSR_HD double m_pow(double a, double b) { return ::pow(a, b); }
VSR_HD float m_pow(float a, float b) { return ::powf(a, b); }
VSR_HD void m_sincos(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  ::sincos(x, s, c);
#else
  *s = ::sin(x);
  *c = ::cos(x);
#endif
}
VSR_HD void m_sincos(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
  ::sincosf(x, s, c);
#else
  *s = ::sinf(x);
  *c = ::cosf(x);
#endif
}
template <typename T>
VSR_HD T m_sign(T x) {
  return x > T(0) ? T(1) : (x < T(0) ? T(-1) : (x == T(0) ? T(0) : x));
}
// x**n for integer n >= 0 by binary powering (sympy Integer exponents)
template <typename T>
VSR_HD T m_powi(T x, int n) {
  T r = T(1);
  T b = x;
  while (n > 0) {
    if (n & 1) r *= b;
    n >>= 1;
    if (n) b *= b;
  }
  return r;
}
// multiply a tangent by a derivative that may be infinite while the tangent is an
// exact zero (sqrt at 0, asin at +-1): 0*inf must stay 0.
template <typename T>
VSR_HD T m_scale(T d, T fp) {
  return d == T(0) ? T(0) : d * fp;
}

// ---- dual numbers -----------------------------------------------------------------
template <typename T, int K>
struct Dual {
  T v;
  T d[K > 0 ? K : 1];
};

// Operand stack.  Only binary nodes whose two children are both non-leaf push, and
// the compiler orders children Sethi-Ullman style, so depth stays tiny; the array is
// dynamically indexed and therefore lives in (L1-resident) local memory on the GPU.
template <typename T, int K, int P>
struct Stack {
  T s[VSR_MAX_STACK][P][K + 1];
};

// One unary node applied to one dual number.  OP is a compile-time constant, so the
// switch folds away.  `n` is the POWI exponent.
template <int OP, typename T, int K>
VSR_HD void unary_apply(Dual<T, K>& x, unsigned am, int n) {
  const T a = x.v;
  const bool live = (K > 0) && am != 0u;
  T f = a, fp = T(1);
  bool guard = false;  // derivative may be infinite where the value is finite
  switch (OP) {
    case VSR_NEG:
      f = -a;
      fp = T(-1);
      break;
    case VSR_ABS:
      f = m_abs(a);
      fp = m_sign(a);
      break;
    case VSR_SIGN:
      f = m_sign(a);
      fp = T(0);
      break;
    case VSR_INV:
      f = T(1) / a;
      fp = -f * f;
      break;
    case VSR_SQRT:
      f = m_sqrt(a);
      fp = T(0.5) / f;
      guard = true;
      break;
    case VSR_EXP:
      f = m_exp(a);
      fp = f;
      break;
    case VSR_LOG:
      f = m_log(a);
      fp = T(1) / a;
      break;
    case VSR_SIN:
      if (live) {
        m_sincos(a, &f, &fp);
      } else {
        f = m_sin(a);
      }
      break;
    case VSR_COS:
      if (live) {
        T s;
        m_sincos(a, &s, &f);
        fp = -s;
      } else {
        f = m_cos(a);
      }
      break;
    case VSR_TAN:
      f = m_tan(a);
      fp = T(1) + f * f;
      break;
    case VSR_ASIN:
      f = m_asin(a);
      if (live) fp = T(1) / m_sqrt(T(1) - a * a);
      guard = true;
      break;
    case VSR_ACOS:
      f = m_acos(a);
      if (live) fp = T(-1) / m_sqrt(T(1) - a * a);
      guard = true;
      break;
    case VSR_ATAN:
      f = m_atan(a);
      fp = T(1) / (T(1) + a * a);
      break;
    case VSR_SINH:
      f = m_sinh(a);
      if (live) fp = m_cosh(a);
      break;
    case VSR_COSH:
      f = m_cosh(a);
      if (live) fp = m_sinh(a);
      break;
    case VSR_TANH:
      f = m_tanh(a);
      fp = T(1) - f * f;
      break;
    case VSR_POWI: {
      const int m = n < 0 ? -n : n;
      const T pw = m_powi(a, m - 1);  // a^(|n|-1)
      const T full = pw * a;          // a^|n|
      if (n > 0) {
        f = full;
        fp = T(n) * pw;
      } else {
        f = T(1) / full;
        fp = T(n) * f / a;
      }
      break;
    }
    default:
      break;
  }
  x.v = f;
  if (K > 0) {
    if (guard) {
#pragma unroll
      for (int i = 0; i < K; ++i)
        if ((am >> i) & 1u) x.d[i] = m_scale(x.d[i], fp);
    } else {
#pragma unroll
      for (int i = 0; i < K; ++i)
        if ((am >> i) & 1u) x.d[i] *= fp;
    }
  }
}

// Where the points come from: col(j, p) returns x_{j+1} of the p-th point of this call.
// (a functor so the kernels can read registers/shared memory and the host simulator a
// plain array.)

template <typename T, int K, int P, typename XSrc>
VSR_HD void eval_points(const vsr_insn_t* __restrict__ prog, const double* __restrict__ imm,
                        const T* __restrict__ cst, const XSrc& xs, Dual<T, K> (&acc)[P],
                        Stack<T, K, P>& stk) {
  int sp = 0;
#pragma unroll
  for (int p = 0; p < P; ++p) {
    acc[p].v = T(0);
#pragma unroll
    for (int i = 0; i < (K > 0 ? K : 1); ++i) acc[p].d[i] = T(0);
  }
  for (int pc = 0;; ++pc) {
    const vsr_insn_t w = prog[pc];
    const unsigned op = VSR_OP(w);
    if (op == VSR_END) break;
    const unsigned am = VSR_AMASK(w);
    const unsigned bm = VSR_BMASK(w);
    const unsigned lm = am | bm;
    (void)lm;

    if (op == VSR_PUSH) {
#pragma unroll
      for (int p = 0; p < P; ++p) {
        stk.s[sp][p][0] = acc[p].v;
#pragma unroll
        for (int i = 0; i < K; ++i)
          if ((am >> i) & 1u) stk.s[sp][p][i + 1] = acc[p].d[i];
      }
      ++sp;
      continue;
    }

    if (op <= VSR_RPOW) {  // LOAD and the binary ops: fetch the operand first
      const unsigned src = VSR_SRC(w);
      const unsigned idx = VSR_IDX(w);
      T bv[P];
      T bd[P][K > 0 ? K : 1];
#pragma unroll
      for (int p = 0; p < P; ++p)
#pragma unroll
        for (int i = 0; i < (K > 0 ? K : 1); ++i) bd[p][i] = T(0);
      if (src == VSR_SRC_STACK) {
        --sp;
#pragma unroll
        for (int p = 0; p < P; ++p) {
          bv[p] = stk.s[sp][p][0];
#pragma unroll
          for (int i = 0; i < K; ++i)
            if ((bm >> i) & 1u) bd[p][i] = stk.s[sp][p][i + 1];
        }
      } else if (src == VSR_SRC_VAR) {
#pragma unroll
        for (int p = 0; p < P; ++p) bv[p] = xs.col(idx, p);
      } else if (src == VSR_SRC_CONST) {
        const T c = cst[idx];
#pragma unroll
        for (int p = 0; p < P; ++p) {
          bv[p] = c;
#pragma unroll
          for (int i = 0; i < K; ++i) bd[p][i] = (i == (int)idx) ? T(1) : T(0);
        }
      } else {
        const T c = (T)imm[idx];
#pragma unroll
        for (int p = 0; p < P; ++p) bv[p] = c;
      }

      switch (op) {
        case VSR_LOAD:
#pragma unroll
          for (int p = 0; p < P; ++p) {
            acc[p].v = bv[p];
#pragma unroll
            for (int i = 0; i < K; ++i) acc[p].d[i] = bd[p][i];
          }
          break;
        case VSR_ADD:
#pragma unroll
          for (int p = 0; p < P; ++p) {
            acc[p].v += bv[p];
#pragma unroll
            for (int i = 0; i < K; ++i)
              if ((bm >> i) & 1u) acc[p].d[i] += bd[p][i];
          }
          break;
        case VSR_SUB:
#pragma unroll
          for (int p = 0; p < P; ++p) {
            acc[p].v -= bv[p];
#pragma unroll
            for (int i = 0; i < K; ++i)
              if ((bm >> i) & 1u) acc[p].d[i] -= bd[p][i];
          }
          break;
        case VSR_RSUB:
#pragma unroll
          for (int p = 0; p < P; ++p) {
            acc[p].v = bv[p] - acc[p].v;
#pragma unroll
            for (int i = 0; i < K; ++i)
              if ((lm >> i) & 1u) acc[p].d[i] = bd[p][i] - acc[p].d[i];
          }
          break;
        case VSR_MUL:
#pragma unroll
          for (int p = 0; p < P; ++p) {
            const T a = acc[p].v;
#pragma unroll
            for (int i = 0; i < K; ++i)
              if ((lm >> i) & 1u) acc[p].d[i] = acc[p].d[i] * bv[p] + a * bd[p][i];
            acc[p].v = a * bv[p];
          }
          break;
        case VSR_DIV:  // acc / b
#pragma unroll
          for (int p = 0; p < P; ++p) {
            const T q = acc[p].v / bv[p];
#pragma unroll
            for (int i = 0; i < K; ++i)
              if ((lm >> i) & 1u) acc[p].d[i] = (acc[p].d[i] - q * bd[p][i]) / bv[p];
            acc[p].v = q;
          }
          break;
        case VSR_RDIV:  // b / acc
#pragma unroll
          for (int p = 0; p < P; ++p) {
            const T a = acc[p].v;
            const T q = bv[p] / a;
#pragma unroll
            for (int i = 0; i < K; ++i)
              if ((lm >> i) & 1u) acc[p].d[i] = (bd[p][i] - q * acc[p].d[i]) / a;
            acc[p].v = q;
          }
          break;
        case VSR_POW:  // acc ** b
#pragma unroll
          for (int p = 0; p < P; ++p) {
            const T a = acc[p].v;
            const T f = m_pow(a, bv[p]);
            if (K > 0 && lm) {
              // d/da = b a^(b-1);  d/db = f ln a  (0 where a == 0 and the power vanishes)
              const T fa = am ? bv[p] * m_pow(a, bv[p] - T(1)) : T(0);
              T fb = T(0);
              if (bm) fb = (a == T(0) && f == T(0)) ? T(0) : f * m_log(a);
#pragma unroll
              for (int i = 0; i < K; ++i)
                if ((lm >> i) & 1u)
                  acc[p].d[i] = m_scale(acc[p].d[i], fa) + m_scale(bd[p][i], fb);
            }
            acc[p].v = f;
          }
          break;
        case VSR_RPOW:  // b ** acc
#pragma unroll
          for (int p = 0; p < P; ++p) {
            const T a = acc[p].v;
            const T f = m_pow(bv[p], a);
            if (K > 0 && lm) {
              T fa = T(0);
              if (am) fa = (bv[p] == T(0) && f == T(0)) ? T(0) : f * m_log(bv[p]);
              const T fb = bm ? a * m_pow(bv[p], a - T(1)) : T(0);
#pragma unroll
              for (int i = 0; i < K; ++i)
                if ((lm >> i) & 1u)
                  acc[p].d[i] = m_scale(acc[p].d[i], fa) + m_scale(bd[p][i], fb);
            }
            acc[p].v = f;
          }
          break;
        default:
          break;
      }
      continue;
    }

    // unary ops on acc: v = f(v), d *= f'(v).  The dispatch is outside the loop over
    // the P points so one decode serves all of them.
    switch (op) {
#define VSR_UCASE(OPC)                                                  \
  case OPC:                                                             \
    _Pragma("unroll") for (int p = 0; p < P; ++p)                       \
        unary_apply<OPC, T, K>(acc[p], am, (int)(int16_t)VSR_IDX(w));   \
    break;
      VSR_UCASE(VSR_NEG)
      VSR_UCASE(VSR_ABS)
      VSR_UCASE(VSR_SIGN)
      VSR_UCASE(VSR_INV)
      VSR_UCASE(VSR_SQRT)
      VSR_UCASE(VSR_EXP)
      VSR_UCASE(VSR_LOG)
      VSR_UCASE(VSR_SIN)
      VSR_UCASE(VSR_COS)
      VSR_UCASE(VSR_TAN)
      VSR_UCASE(VSR_ASIN)
      VSR_UCASE(VSR_ACOS)
      VSR_UCASE(VSR_ATAN)
      VSR_UCASE(VSR_SINH)
      VSR_UCASE(VSR_COSH)
      VSR_UCASE(VSR_TANH)
      VSR_UCASE(VSR_POWI)
#undef VSR_UCASE
      default:
        break;
    }
  }
}

}  // namespace vsr

#endif  // VSR_INTERP_H_
