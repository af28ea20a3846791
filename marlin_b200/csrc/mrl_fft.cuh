// marlin_b200 - FFT building blocks for sm_100a (also compiled by NVRTC and, for logic tests
// only, by the host emulation in tests/emu).
//
// Replaces the reference's torch::fft::rfftn / irfftn dispatch
// (src/actions/DomainAction.C:854-867, :1054-1066).  Layout contract kept bit-exact:
// C-order [x][y][z], half spectrum on the LAST axis, forward unnormalised, inverse 1/N.
//
// Design: every 1-D transform is a FORWARD complex Stockham autosort FFT; inverses use
// conj(FFT(conj(.))).  Real transforms pack two real sequences into one complex pencil
// (z = a + i b) and separate / merge the two Hermitian half spectra in the pass epilogue /
// prologue, so any length (even or odd) runs the same code.
//   * RegFFT    - power-of-two and other smooth sizes: E = N/TP points per thread live in
//                 registers, radix-R butterflies (R in {2,3,4,5,8}) in registers, one
//                 shared-memory exchange between stages.
//   * smem_fft  - any N: runtime mixed-radix Stockham in shared memory with a generic O(R^2)
//                 butterfly for prime factors without a specialised kernel.
#pragma once

#if defined(MRL_EMU)
#include "cuda_emu.h"
#define MRL_DI inline
#define MRL_HD inline
#define MRL_HDC constexpr
#define MRL_UNROLL
#define MRL_DYN_SMEM(name) unsigned char *name = emu::g_dyn_smem
#else
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif
#define MRL_DI __device__ __forceinline__
#define MRL_HD __host__ __device__ __forceinline__
#define MRL_HDC __host__ __device__ constexpr
#define MRL_UNROLL _Pragma("unroll")
#define MRL_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace mrl {

// ------------------------------------------------------------------ complex value type
template <class T> struct alignas(2 * sizeof(T)) cx { T x, y; };

template <class T> MRL_HD cx<T> mk(T x, T y) { cx<T> r; r.x = x; r.y = y; return r; }
template <class T> MRL_HD cx<T> operator+(cx<T> a, cx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <class T> MRL_HD cx<T> operator-(cx<T> a, cx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <class T> MRL_HD cx<T> operator*(cx<T> a, cx<T> b) {
  return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <class T> MRL_HD cx<T> operator*(cx<T> a, T s) { return mk<T>(a.x * s, a.y * s); }
template <class T> MRL_HD cx<T> conj(cx<T> a) { return mk<T>(a.x, -a.y); }
template <class T> MRL_HD cx<T> mul_mi(cx<T> a) { return mk<T>(a.y, -a.x); }  // a * (-i)
template <class T> MRL_HD cx<T> mul_pi(cx<T> a) { return mk<T>(-a.y, a.x); }  // a * (+i)

// ------------------------------------------------------------------ radix butterflies
// In-place forward DFT of R points: a[k] <- sum_r a[r] exp(-2 pi i r k / R).
template <class T, int R> struct Butterfly;

template <class T> struct Butterfly<T, 2> {
  static MRL_DI void run(cx<T> *a) {
    cx<T> t = a[0] - a[1];
    a[0] = a[0] + a[1];
    a[1] = t;
  }
};
template <class T> struct Butterfly<T, 3> {
  static MRL_DI void run(cx<T> *a) {
    const T h = T(0.86602540378443864676372317075294);  // sqrt(3)/2
    cx<T> s = a[1] + a[2], d = a[1] - a[2];
    cx<T> m = mk<T>(a[0].x - T(0.5) * s.x, a[0].y - T(0.5) * s.y);
    cx<T> w = mul_mi(d) * h;
    a[0] = a[0] + s;
    a[1] = m + w;
    a[2] = m - w;
  }
};
template <class T> struct Butterfly<T, 4> {
  static MRL_DI void run(cx<T> *a) {
    cx<T> t0 = a[0] + a[2], t1 = a[0] - a[2], t2 = a[1] + a[3], t3 = mul_mi(a[1] - a[3]);
    a[0] = t0 + t2;
    a[1] = t1 + t3;
    a[2] = t0 - t2;
    a[3] = t1 - t3;
  }
};
template <class T> struct Butterfly<T, 5> {
  static MRL_DI void run(cx<T> *a) {
    const T c1 = T(0.30901699437494742410229341718282);   // cos(2pi/5)
    const T c2 = T(-0.80901699437494742410229341718282);  // cos(4pi/5)
    const T s1 = T(0.95105651629515357211643933337938);   // sin(2pi/5)
    const T s2 = T(0.58778525229247312916870595463907);   // sin(4pi/5)
    cx<T> p1 = a[1] + a[4], p2 = a[2] + a[3], d1 = a[1] - a[4], d2 = a[2] - a[3];
    cx<T> m1 = mk<T>(a[0].x + c1 * p1.x + c2 * p2.x, a[0].y + c1 * p1.y + c2 * p2.y);
    cx<T> m2 = mk<T>(a[0].x + c2 * p1.x + c1 * p2.x, a[0].y + c2 * p1.y + c1 * p2.y);
    cx<T> n1 = mul_mi(mk<T>(s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y));
    cx<T> n2 = mul_mi(mk<T>(s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y));
    a[0] = a[0] + p1 + p2;
    a[1] = m1 + n1;
    a[4] = m1 - n1;
    a[2] = m2 + n2;
    a[3] = m2 - n2;
  }
};
template <class T> struct Butterfly<T, 8> {
  static MRL_DI void run(cx<T> *a) {
    const T c = T(0.70710678118654752440084436210485);
    cx<T> e[4] = {a[0], a[2], a[4], a[6]};
    cx<T> o[4] = {a[1], a[3], a[5], a[7]};
    Butterfly<T, 4>::run(e);
    Butterfly<T, 4>::run(o);
    cx<T> o1 = mk<T>(c * (o[1].x + o[1].y), c * (o[1].y - o[1].x));   // * W8^1
    cx<T> o2 = mul_mi(o[2]);                                          // * W8^2
    cx<T> o3 = mk<T>(c * (o[3].y - o[3].x), -c * (o[3].x + o[3].y));  // * W8^3
    a[0] = e[0] + o[0];
    a[4] = e[0] - o[0];
    a[1] = e[1] + o1;
    a[5] = e[1] - o1;
    a[2] = e[2] + o2;
    a[6] = e[2] - o2;
    a[3] = e[3] + o3;
    a[7] = e[3] - o3;
  }
};

// ------------------------------------------------------------------ shared-memory views
// Strided tile: element idx of column col at base[idx*TK + col] (conflict-free when a
// quarter-warp spans the TK*sizeof(cx) = 128 B of one idx row).
template <class T, int TK> struct SmTile {
  cx<T> *base;
  int col;
  MRL_DI void st(int idx, cx<T> v) const { base[idx * TK + col] = v; }
  MRL_DI cx<T> ld(int idx) const { return base[idx * TK + col]; }
};
// Contiguous pencil with one pad element every 8 (keeps the radix-8 stage scatter
// idx = 8 b + k conflict-free for 16-byte elements).
template <class T> struct SmPencil {
  cx<T> *base;
  static MRL_HD int padded(int n) { return n + (n >> 3) + 1; }
  MRL_DI void st(int idx, cx<T> v) const { base[idx + (idx >> 3)] = v; }
  MRL_DI cx<T> ld(int idx) const { return base[idx + (idx >> 3)]; }
};

// ------------------------------------------------------------------ register FFT
// N = R0*R1*R2*R3 points, TP threads per pencil, E = N/TP points per thread.  Thread t
// owns x[t + TP*e], e in [0,E), before AND after the transform (natural order out).
template <int N_, int TP_, int R0_, int R1_ = 1, int R2_ = 1, int R3_ = 1> struct FFTCfg {
  static constexpr int N = N_, TP = TP_, E = N_ / TP_;
  static constexpr int R0 = R0_, R1 = R1_, R2 = R2_, R3 = R3_;
  static constexpr int NS = (R1_ == 1) ? 1 : (R2_ == 1) ? 2 : (R3_ == 1) ? 3 : 4;
  static_assert(R0_ * R1_ * R2_ * R3_ == N_, "radices must multiply to N");
  static_assert(N_ % TP_ == 0, "TP must divide N");
  static_assert(E % R0_ == 0 && E % R1_ == 0 && E % R2_ == 0 && E % R3_ == 0,
                "every radix must divide the points per thread");
};

// Barrier policy of a RegFFT exchange: the whole CTA (default) or one named-barrier group.
struct CtaSync {
  MRL_DI void sync() const { __syncthreads(); }
  MRL_DI void sync_release() const { __syncthreads(); }  // barrier after which the buffer may be re-armed
};
struct NoHook {
  MRL_DI void operator()() const {}
};

// Twiddle sources for the inter-stage factors W_N^{S p k}, k in [1,R).
//  TwTable: looked up in a table of exp(-2 pi i k / N) (shared or global memory).
//  TwRegs : for power-of-two radices the factors of one butterfly are powers of one root
//           w = W_N^{S p}; a thread's (stage, butterfly) pairs never change, so it keeps
//           w, w^2 (, w^4) in registers for the whole kernel and forms the other powers with a
//           few multiplications - no shared-memory traffic for twiddles at all.
template <class T> struct TwTable {
  const cx<T> *tw;
  template <int ST, int R, int S, int EB> MRL_DI void apply(cx<T> (&a)[R], int u, int b) const {
    const int p = b / S;
    MRL_UNROLL
    for (int k = 1; k < R; ++k) a[k] = a[k] * tw[S * p * k];
  }
};

// LEAN = true keeps only w per butterfly and squares it on the fly (fewer registers, 2 more
// complex multiplications per radix-8 butterfly).
template <class T, class C, bool LEAN = false> struct TwRegs {
  static constexpr int E = C::E, TP = C::TP;
  static MRL_HDC int lg(int r) { return r == 8 ? 3 : r == 4 ? 2 : r == 2 ? 1 : 0; }
  static MRL_HDC int radix(int st) { return st == 0 ? C::R0 : st == 1 ? C::R1 : st == 2 ? C::R2 : C::R3; }
  static MRL_HDC int span(int st) { return st == 0 ? 1 : st == 1 ? C::R0 : st == 2 ? C::R0 * C::R1 : C::R0 * C::R1 * C::R2; }
  static MRL_HDC int nb(int r) { return LEAN ? 1 : lg(r); }  // bases kept per butterfly
  static MRL_HDC int count(int st) { return st >= C::NS - 1 ? 0 : (E / radix(st)) * nb(radix(st)); }
  static MRL_HDC int offset(int st) { return st == 0 ? 0 : offset(st - 1) + count(st - 1); }
  static constexpr int TOTAL = offset(C::NS - 1) > 0 ? offset(C::NS - 1) : 1;
  static_assert((C::R0 == 8 || C::R0 == 4 || C::R0 == 2) && (C::R1 == 8 || C::R1 == 4 || C::R1 == 2 || C::R1 == 1) &&
                    (C::R2 == 8 || C::R2 == 4 || C::R2 == 2 || C::R2 == 1) && (C::R3 == 8 || C::R3 == 4 || C::R3 == 2 || C::R3 == 1),
                "TwRegs needs power-of-two radices");
  cx<T> w[TOTAL];

  template <int ST> MRL_DI void init_stage(const cx<T> *tw, int t) {
    if constexpr (ST < C::NS - 1) {
      constexpr int R = radix(ST), S = span(ST), EB = E / R, L = nb(R);
      MRL_UNROLL
      for (int u = 0; u < EB; ++u) {
        const int p = (t + TP * u) / S;
        MRL_UNROLL
        for (int l = 0; l < L; ++l) w[offset(ST) + u * L + l] = tw[S * p * (1 << l)];
      }
      init_stage<ST + 1>(tw, t);
    }
  }
  // tw: table of exp(-2 pi i k / N) (any memory space); t: this thread's index within its pencil
  MRL_DI void init(const cx<T> *tw, int t) { init_stage<0>(tw, t); }

  template <int ST, int R, int S, int EB> MRL_DI void apply(cx<T> (&a)[R], int u, int) const {
    constexpr int L = nb(R);
    const cx<T> *b = w + offset(ST) + u * L;
    const cx<T> w1 = b[0];
    if constexpr (R == 2) {
      a[1] = a[1] * w1;
    } else {
      const cx<T> w2 = LEAN ? w1 * w1 : b[L > 1 ? 1 : 0];
      const cx<T> w3 = w1 * w2;
      a[1] = a[1] * w1;
      a[2] = a[2] * w2;
      a[3] = a[3] * w3;
      if constexpr (R == 8) {
        const cx<T> w4 = LEAN ? w2 * w2 : b[L > 2 ? 2 : 0];
        a[4] = a[4] * w4;
        a[5] = a[5] * (w1 * w4);
        a[6] = a[6] * (w2 * w4);
        a[7] = a[7] * (w3 * w4);
      }
    }
  }
};

template <class T, class C> struct RegFFT {
  static constexpr int E = C::E, TP = C::TP;

  // HOOK runs once, right after the last shared-memory exchange has been read back (from then
  // on the exchange buffer is no longer touched by this transform).
  template <int ST, int R, int S, class SM, class TW, class BAR, class HOOK>
  static MRL_DI void stage(cx<T> (&v)[C::E], int t, const SM &sm, const TW &tw, const BAR &bar, const HOOK &hook) {
    constexpr int EB = E / R;
    constexpr bool last = (ST == C::NS - 1);
    MRL_UNROLL
    for (int u = 0; u < EB; ++u) {
      cx<T> a[R];
      MRL_UNROLL
      for (int r = 0; r < R; ++r) a[r] = v[u + r * EB];
      Butterfly<T, R>::run(a);
      if constexpr (last) {
        MRL_UNROLL
        for (int k = 0; k < R; ++k) v[u + k * EB] = a[k];
      } else {
        const int b = t + TP * u;
        const int q = b % S, p = b / S;
        tw.template apply<ST, R, S, EB>(a, u, b);
        MRL_UNROLL
        for (int k = 0; k < R; ++k) sm.st(q + S * (R * p + k), a[k]);
      }
    }
    if constexpr (!last) {
      bar.sync();
      MRL_UNROLL
      for (int e = 0; e < E; ++e) v[e] = sm.ld(t + TP * e);
      if constexpr (ST == C::NS - 2) {
        bar.sync_release();
        hook();
      } else {
        bar.sync();
      }
    }
  }

  template <class SM, class TW, class BAR, class HOOK>
  static MRL_DI void run_tw(cx<T> (&v)[C::E], int t, const SM &sm, const TW &tw, const BAR &bar, const HOOK &hook) {
    stage<0, C::R0, 1>(v, t, sm, tw, bar, hook);
    if constexpr (C::NS > 1) stage<1, C::R1, C::R0>(v, t, sm, tw, bar, hook);
    if constexpr (C::NS > 2) stage<2, C::R2, C::R0 * C::R1>(v, t, sm, tw, bar, hook);
    if constexpr (C::NS > 3) stage<3, C::R3, C::R0 * C::R1 * C::R2>(v, t, sm, tw, bar, hook);
    if constexpr (C::NS == 1) hook();
  }

  // tw: table of exp(-2 pi i k / N), k in [0,N)
  template <class SM, class BAR = CtaSync, class HOOK = NoHook>
  static MRL_DI void run(cx<T> (&v)[C::E], int t, const SM &sm, const cx<T> *tw, const BAR &bar = BAR(),
                         const HOOK &hook = HOOK()) {
    run_tw(v, t, sm, TwTable<T>{tw}, bar, hook);
  }
};

// ------------------------------------------------------------------ generic smem FFT
struct FFTPlanDev {
  int n;
  int nstages;
  int radix[20];
};

// One butterfly (index b, column col) of a Stockham stage: reads src, writes dst.
template <class T, int R, int TK>
MRL_DI void smem_butterfly(const cx<T> *src, cx<T> *dst, const cx<T> *tw, int n, int s, int b, int col) {
  const int m = n / R;  // input stride
  cx<T> a[R];
  MRL_UNROLL
  for (int r = 0; r < R; ++r) a[r] = src[(b + r * m) * TK + col];
  Butterfly<T, R>::run(a);
  const int q = b % s, p = b / s;
  dst[(q + s * (R * p)) * TK + col] = a[0];
  MRL_UNROLL
  for (int k = 1; k < R; ++k) dst[(q + s * (R * p + k)) * TK + col] = a[k] * tw[(s * p * k) % n];
}

template <class T, int TK>
MRL_DI void smem_butterfly_any(const cx<T> *src, cx<T> *dst, const cx<T> *tw, int n, int R, int s, int b,
                               int col) {
  const int m = n / R;
  const int q = b % s, p = b / s;
  for (int k = 0; k < R; ++k) {
    cx<T> acc = mk<T>(T(0), T(0));
    for (int r = 0; r < R; ++r) {
      // W_R^{r k} = W_n^{m r k}; keep the index in range with 64-bit arithmetic
      const int wi = (int)(((long long)m * r * k) % n);
      acc = acc + src[(b + r * m) * TK + col] * tw[wi];
    }
    dst[(q + s * (R * p + k)) * TK + col] = acc * tw[(int)(((long long)s * p * k) % n)];
  }
}

// Forward FFT of TK interleaved pencils of length plan.n held in A ([idx*TK+col]); B is
// scratch of the same size.  Returns the buffer holding the result.  All threads of the CTA
// must call it (tid in [0,nthreads)).
template <class T, int TK>
MRL_DI cx<T> *smem_fft(cx<T> *A, cx<T> *B, const cx<T> *tw, const FFTPlanDev &plan, int tid, int nthreads) {
  const int n = plan.n;
  int s = 1;
  cx<T> *src = A, *dst = B;
  for (int st = 0; st < plan.nstages; ++st) {
    const int R = plan.radix[st];
    const int nb = (n / R) * TK;
    for (int w = tid; w < nb; w += nthreads) {
      const int b = w / TK, col = w % TK;
      switch (R) {
        case 2: smem_butterfly<T, 2, TK>(src, dst, tw, n, s, b, col); break;
        case 3: smem_butterfly<T, 3, TK>(src, dst, tw, n, s, b, col); break;
        case 4: smem_butterfly<T, 4, TK>(src, dst, tw, n, s, b, col); break;
        case 5: smem_butterfly<T, 5, TK>(src, dst, tw, n, s, b, col); break;
        case 8: smem_butterfly<T, 8, TK>(src, dst, tw, n, s, b, col); break;
        default: smem_butterfly_any<T, TK>(src, dst, tw, n, R, s, b, col); break;
      }
    }
    __syncthreads();
    s *= R;
    cx<T> *tmp = src;
    src = dst;
    dst = tmp;
  }
  return src;
}

}  // namespace mrl
