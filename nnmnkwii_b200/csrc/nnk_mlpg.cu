// nnk_mlpg.cu -- batched MLPG (maximum likelihood parameter generation) on sm_100a.
//
// Replaces, for a whole batch of utterances and every static dimension at once:
//   paramgen.mlpg       paramgen/_mlpg.py:92-199  (build_poe :53-89 -> _bandmat/tensor.pyx:20-64,
//                       82-174; bla.solveh -> _bandmat/linalg.pyx:36-104, 106-176, 290-304)
//   paramgen.mlpg_grad  paramgen/_mlpg.py:202-281 (closed form  tau_w * (W_w P^-1 o), O(T))
//
// Mapping: one "chain" = one static dimension of one stream of one utterance = one symmetric
// banded T x T system  P y = b,  P = sum_w W_w^T diag(tau_w) W_w,  b = sum_w W_w^T (tau_w * mu_w).
// One warp owns 32 chains of one utterance (lane = feature dimension, so every global access is a
// coalesced row segment of the (T, D) frame matrix) and walks time:
//   forward sweep : assemble row t of P and b in registers from a sliding window of frames,
//                   eliminate (L D L^T, band depth S = L+U), forward-substitute; the S+1 numbers a
//                   frame needs later (z/d, l_1..l_S) go to a float64 scratch laid out
//                   [t][j][lane] so each access is one 256 B line per warp;
//   backward sweep: y[t] = zs[t] - sum_j l_j[t] y[t+j], written to the output in the input dtype.
// All arithmetic is float64 (the reference's bandmat is float64-only); P is never materialised in
// HBM.  L D L^T instead of the reference's L L^T: same pivots d[t] (so the same not-positive-
// definite test, linalg.pyx:79), no sqrt on the loop-carried dependency chain, results agree to
// rounding (1e-15 relative).
#include <stdlib.h>

#include "nnk_mlpg.cuh"
#include "nnk_mlpg_tma.cuh"
#include "nnk_mlpg_as.cuh"

// assembler warps per chain group and TMA stages per assembler (A/B builds: -DNNK_AS_NA=2 -DNNK_AS_NSA=2)
#ifndef NNK_AS_NA
#define NNK_AS_NA 3
#define NNK_AS_NSA 1
#endif
// depth of the band-row (PB) ring in tiles; must be >= the producers' tile stride (see nnk_mlpg_as.cuh)
#ifndef NNK_AS_ND
#define NNK_AS_ND (NNK_AS_PAIRS ? 6 : 4)
#endif

namespace nnk {

__device__ __forceinline__ double load_go(const void* go, int is_f64, int64_t idx) {
  return is_f64 ? ld_stream(reinterpret_cast<const double*>(go) + idx)
                : (double)ld_stream(reinterpret_cast<const float*>(go) + idx);
}

template <typename Tin, int NW, int L, int U, int MODE, int PF>
__global__ void __launch_bounds__(32) mlpg_kernel(const __grid_constant__ MlpgParams<Tin, NW, L, U> p) {
  constexpr int S = L + U;
  constexpr int NT = S + 1;
  const int lane = threadIdx.x;
  const int item = blockIdx.x;
  const int urank = p.urank0 + item / p.n_groups;
  const int grp = item % p.n_groups;
  const int utt = p.order ? p.order[urank] : urank;
  const int64_t row0 = p.utt_off[utt];
  const int T = p.utt_len ? p.utt_len[utt] : (int)(p.utt_off[utt + 1] - row0);
  if (T <= 0) return;
  const int64_t orow0 = p.out_off ? p.out_off[utt] : row0;
  const int chain = grp * 32 + lane;
  const bool active = chain < p.n_chain;
  nnk_chain_t ch;
  ch.in_col = 0; ch.win_stride = 0; ch.out_col = 0; ch.flags = 1;
  if (active) ch = p.chains[chain];
  const bool solve = active && !(ch.flags & 1);
  const int nw = p.win.nw;
  const int m_edge = p.win.m_edge;
  const bool var_global = (p.var_ld == 0);

  const Tin* mptr = p.means + row0 * p.in_ld + ch.in_col;
  const Tin* vptr = p.vars + (var_global ? 0 : row0 * p.var_ld) + ch.in_col;
  double* ws = p.ws + (size_t)item * ((size_t)p.max_T * NT * 32) + lane;

  // ---- pass-through chains (flags & 1): plain copy (fwd) / gradient of a copy (grad) -----------
  if (active && (ch.flags & 1)) {
    if (MODE == MODE_FWD) {
      Tin* o = reinterpret_cast<Tin*>(p.out) + orow0 * p.out_ld + ch.out_col;
      for (int t = 0; t < T; ++t) o[(int64_t)t * p.out_ld] = mptr[(int64_t)t * p.in_ld];
    } else if (MODE == MODE_GRAD) {
      float* o = reinterpret_cast<float*>(p.out) + orow0 * p.out_ld + ch.in_col;
      for (int t = 0; t < T; ++t) o[(int64_t)t * p.out_ld] = (float)load_go(p.go, p.go_f64, (row0 + t) * p.go_ld + chain);
    }
  }

  Tin gv[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w) gv[w] = (solve && var_global && w < nw) ? vptr[w * ch.win_stride] : Tin(1);

  // raw (un-converted) frame loads; t outside [0, T) or idle lanes give zeros
  auto load_raw = [&](int t, Tin(&m)[NW], Tin(&v)[NW], double& g) {
#pragma unroll
    for (int w = 0; w < NW; ++w) { m[w] = Tin(0); v[w] = Tin(1); }
    g = 0.0;
    if (t >= 0 && t < T && solve) {
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        if (w < nw) {
          if (MODE == MODE_FWD) m[w] = ld_stream(mptr + (int64_t)t * p.in_ld + w * ch.win_stride);
          v[w] = var_global ? gv[w] : ld_stream(vptr + (int64_t)t * p.var_ld + w * ch.win_stride);
        }
      }
      if (MODE != MODE_FWD) g = load_go(p.go, p.go_f64, (row0 + t) * p.go_ld + chain);
    }
  };
  // tau_w[t] with the reference's edge rule; tm = tau * mu (paramgen/_mlpg.py:188-195)
  auto to_frame = [&](int t, const Tin(&m)[NW], const Tin(&v)[NW], double(&tau)[NW], double(&tm)[NW]) {
    const bool in = (t >= 0 && t < T);
    const bool edge = (m_edge == 0) || (t < m_edge) || (t >= T - m_edge);
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const bool on = in && (w < nw) && !(w > 0 && edge);
      const double tw = on ? recip_in_dtype<Tin>::f(v[w]) : 0.0;
      tau[w] = tw;
      tm[w] = tw * (double)m[w];
    }
  };

  // ---- forward sweep -----------------------------------------------------------------------------
  double wt[NT][NW], wm[NT][NW];  // wt[i] = tau of frame (t + L - i)
  double wg[NT];                  // grad/solve mode: right-hand side of frame (t + L - i)
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    wg[i] = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) { wt[i][w] = 0.0; wm[i][w] = 0.0; }
  }
  // window before step 0 (after the shift at step 0, wt[i] = frame L - i): preload frames 0..L-1
#pragma unroll
  for (int i = 0; i < L; ++i) {  // frame f = L-1-i goes to slot i (it will be shifted to i+1)
    Tin m[NW], v[NW]; double g;
    load_raw(L - 1 - i, m, v, g);
    to_frame(L - 1 - i, m, v, wt[i], wm[i]);
    wg[i] = g;
  }
  Tin rm[PF][NW], rv[PF][NW];
  double rg[PF];
#pragma unroll
  for (int j = 0; j < PF; ++j) load_raw(L + j, rm[j], rv[j], rg[j]);

  double vcol[S + 1][S + 1], lcol[S + 1][S + 1], zz[S + 1];
#pragma unroll
  for (int k = 0; k <= S; ++k) {
    zz[k] = 0.0;
#pragma unroll
    for (int j = 0; j <= S; ++j) { vcol[k][j] = 0.0; lcol[k][j] = 0.0; }
  }
  double iv1 = 0.0;
  bool reported = false;

  for (int t0 = 0; t0 < T; t0 += PF) {
#pragma unroll
    for (int jj = 0; jj < PF; ++jj) {
      const int t = t0 + jj;
      if (t < T) {
        // slide the frame window and take frame t+L from the prefetch ring; refill the ring
#pragma unroll
        for (int i = NT - 1; i > 0; --i) {
          wg[i] = wg[i - 1];
#pragma unroll
          for (int w = 0; w < NW; ++w) { wt[i][w] = wt[i - 1][w]; wm[i][w] = wm[i - 1][w]; }
        }
        to_frame(t + L, rm[jj], rv[jj], wt[0], wm[0]);
        wg[0] = rg[jj];
        load_raw(t + L + PF, rm[jj], rv[jj], rg[jj]);

        // row t of P (acc[m] = P[t][t+m]) and of b
        double acc[S + 1];
#pragma unroll
        for (int m = 0; m <= S; ++m) {
          double a = 0.0;
#pragma unroll
          for (int w = 0; w < NW; ++w)
#pragma unroll
            for (int i = 0; i + m < NT; ++i) a = fma(p.win.q[w][m][i], wt[i][w], a);
          acc[m] = a;
        }
        double bb;
        if (MODE == MODE_FWD) {
          bb = 0.0;
#pragma unroll
          for (int w = 0; w < NW; ++w)
#pragma unroll
            for (int i = 0; i < NT; ++i) bb = fma(p.win.c[w][i], wm[i][w], bb);
        } else {
          bb = wg[L];
        }
        // eliminate: older columns first (their normalised entries are ready) ...
#pragma unroll
        for (int k = 2; k <= S; ++k) {
#pragma unroll
          for (int m = 0; m + k <= S; ++m) acc[m] = fma(-vcol[k][k + m], lcol[k][k], acc[m]);
          bb = fma(-lcol[k][k], zz[k], bb);
        }
        // ... the newest column last: only this part waits on the previous pivot's reciprocal
        if (S >= 1) {
#pragma unroll
          for (int m = 0; m + 1 <= S; ++m) acc[m] = fma(-(vcol[1][1 + m] * vcol[1][1]), iv1, acc[m]);
          bb = fma(-(vcol[1][1] * zz[1]), iv1, bb);
        }
        const double d = acc[0];
        if (!(d > 0.0) && solve && !reported) {  // linalg.pyx:79-82
          reported = true;
          report_not_pd(p.status, utt, chain, t + 1);
        }
        const double ivd = __drcp_rn(d);
        double* wsp = ws + (size_t)t * (NT * 32);
        wsp[0] = bb * ivd;
#pragma unroll
        for (int k = S; k >= 2; --k) {
          zz[k] = zz[k - 1];
#pragma unroll
          for (int j = 0; j <= S; ++j) { vcol[k][j] = vcol[k - 1][j]; lcol[k][j] = lcol[k - 1][j]; }
        }
        if (S >= 1) {
          zz[1] = bb;
#pragma unroll
          for (int j = 1; j <= S; ++j) {
            vcol[1][j] = acc[j];
            const double lj = acc[j] * ivd;
            lcol[1][j] = lj;
            wsp[j * 32] = lj;
          }
          iv1 = ivd;
        }
      }
    }
  }

  // ---- backward sweep ----------------------------------------------------------------------------
  double yw[S + 1];  // yw[j] = y[t + j]
#pragma unroll
  for (int j = 0; j <= S; ++j) yw[j] = 0.0;
  double rz[PF], rl[PF][S + 1];
  auto load_ws = [&](int t, double& z, double(&l)[S + 1]) {
    z = 0.0;
#pragma unroll
    for (int j = 0; j <= S; ++j) l[j] = 0.0;
    if (t >= 0) {
      const double* wsp = ws + (size_t)t * (NT * 32);
      z = wsp[0];
#pragma unroll
      for (int j = 1; j <= S; ++j) l[j] = wsp[j * 32];
    }
  };
#pragma unroll
  for (int j = 0; j < PF; ++j) load_ws(T - 1 - j, rz[j], rl[j]);

  const int t_end = (MODE == MODE_GRAD) ? -L : 0;
  for (int t0 = T - 1; t0 >= t_end; t0 -= PF) {
#pragma unroll
    for (int jj = 0; jj < PF; ++jj) {
      const int t = t0 - jj;
      if (t >= t_end) {
#pragma unroll
        for (int j = S; j > 0; --j) yw[j] = yw[j - 1];
        double y = rz[jj];
#pragma unroll
        for (int j = 1; j <= S; ++j) y = fma(-rl[jj][j], yw[j], y);
        if (t < 0) y = 0.0;
        yw[0] = y;
        load_ws(t - PF, rz[jj], rl[jj]);
        if (MODE != MODE_GRAD) {
          if (solve) st_stream(reinterpret_cast<Tin*>(p.out) + (orow0 + t) * p.out_ld + ch.out_col, (Tin)y);
        } else {
          // row r = t + L of the gradient: tau_w[r] * sum_k c[w][L+k] x[r+k],  x[r+k] = yw[L+k]
          const int r = t + L;
          if (r < T && solve) {
            Tin m[NW], v[NW];
            double tau[NW], tm[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) { m[w] = Tin(0); v[w] = Tin(1); }
#pragma unroll
            for (int w = 0; w < NW; ++w)
              if (w < nw) v[w] = var_global ? gv[w] : vptr[(int64_t)r * p.var_ld + w * ch.win_stride];
            to_frame(r, m, v, tau, tm);
#pragma unroll
            for (int w = 0; w < NW; ++w) {
              if (w < nw) {
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < NT; ++i) s = fma(p.win.c[w][i], yw[i], s);  // c[w][L+k], k = i-L
                reinterpret_cast<float*>(p.out)[(orow0 + r) * p.out_ld + ch.in_col + w * ch.win_stride] =
                    (float)(tau[w] * s);
              }
            }
          }
        }
      }
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------
template <int NW, int L, int U>
static bool fill_wintab(const nnk_windows_t& w, WinTab<NW, L, U>& tab) {
  constexpr int NT = L + U + 1, S = L + U;
  if (w.nw < 1 || w.nw > NW) return false;
  int m_edge = 0;
  for (int i = 0; i < NW; ++i)
    for (int j = 0; j < NT; ++j) tab.c[i][j] = 0.0;
  for (int i = 0; i < w.nw; ++i) {
    if (w.l[i] < 0 || w.u[i] < 0 || w.l[i] > L || w.u[i] > U) return false;
    for (int k = -w.l[i]; k <= w.u[i]; ++k) tab.c[i][L + k] = w.coef[i][w.l[i] + k];
    m_edge = w.l[i] > m_edge ? w.l[i] : m_edge;
    m_edge = w.u[i] > m_edge ? w.u[i] : m_edge;
  }
  for (int i = 0; i < NW; ++i)
    for (int m = 0; m <= S; ++m)
      for (int j = 0; j < NT; ++j) tab.q[i][m][j] = (j + m < NT) ? tab.c[i][j] * tab.c[i][j + m] : 0.0;
  tab.nw = w.nw;
  tab.m_edge = m_edge;
  return true;
}

static void win_extent(const nnk_windows_t& w, int& L, int& U) {
  L = 0; U = 0;
  for (int i = 0; i < w.nw && i < NNK_MAX_WIN; ++i) {
    L = w.l[i] > L ? w.l[i] : L;
    U = w.u[i] > U ? w.u[i] : U;
  }
}

// which template instance serves a window set: returns S (band depth of the instance) or -1
static int pick_instance(const nnk_windows_t& w, int& inst) {
  if (w.nw < 1 || w.nw > NNK_MAX_WIN) return -1;
  int L, U;
  win_extent(w, L, U);
  if (L > NNK_MAX_HALF || U > NNK_MAX_HALF) return -1;
  if (w.nw == 1 && L == 0 && U == 0) { inst = 0; return 0; }
  if (w.nw <= 3 && L <= 1 && U <= 1) { inst = 1; return 2; }
  if (w.nw <= 3 && L <= 2 && U <= 2) { inst = 2; return 4; }
  inst = 3;
  return 2 * NNK_MAX_HALF;
}

// NNK_MLPG_DIRECT=1 selects the register-prefetch kernel for A/B measurements (both are CUDA paths)
static bool force_direct_loads() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NNK_MLPG_DIRECT"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

// NNK_MLPG_SINGLE=1 selects the single-warp TMA kernel instead of the assembler/solver pair
static bool force_single_warp() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NNK_MLPG_SINGLE"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

static bool is_std_windows(const nnk_windows_t& w) {
  if (w.nw != 3 || w.l[0] != 0 || w.u[0] != 0 || w.l[1] != 1 || w.u[1] != 1 || w.l[2] != 1 || w.u[2] != 1) return false;
  return w.coef[0][0] == 1.0 && w.coef[1][0] == -0.5 && w.coef[1][1] == 0.0 && w.coef[1][2] == 0.5 &&
         w.coef[2][0] == 1.0 && w.coef[2][1] == -2.0 && w.coef[2][2] == 1.0;
}

template <typename Tin, int NW, int L, int U, int MODE>
static int launch_mlpg(const nnk_mlpg_args_t& a, cudaStream_t st) {
  constexpr int NT = L + U + 1;
  constexpr int PF = (L + U <= 2) ? 4 : 2;
  MlpgParams<Tin, NW, L, U> p;
  if (!fill_wintab<NW, L, U>(a.win, p.win)) { set_error("window set does not fit kernel instance"); return NNK_ERR_UNSUPPORTED; }
  p.means = (const Tin*)a.means; p.vars = (const Tin*)a.vars; p.go = a.grad_out; p.go_f64 = a.go_f64; p.out = a.out;
  p.in_ld = a.in_ld; p.var_ld = a.var_ld; p.go_ld = a.go_ld; p.out_ld = a.out_ld;
  p.utt_off = a.utt_off; p.out_off = a.out_off; p.utt_len = a.utt_len; p.order = a.order; p.chains = a.chains;
  p.n_utt = a.n_utt; p.n_chain = a.n_chain; p.n_groups = (a.n_chain + 31) / 32; p.max_T = a.max_T;
  p.ws = (double*)a.workspace; p.status = (unsigned long long*)a.status_word;
  const size_t per_item = (size_t)a.max_T * NT * 32 * sizeof(double);
  size_t items_cap = per_item ? a.workspace_bytes / per_item : 0;
  int utt_per_launch = (int)(items_cap / (size_t)p.n_groups);
  if (utt_per_launch < 1) { set_error("workspace too small: need >= %zu bytes", per_item * p.n_groups); return NNK_ERR_WORKSPACE; }
  // forward solves go through the TMA-staged kernel unless the rows are too wide for its ring
  TmaGeom geom;
  size_t smem_bytes = 0;
  constexpr int ES = (int)sizeof(Tin);
  constexpr bool GRAD = (MODE == MODE_GRAD);
  constexpr bool GRAD_MODE = GRAD;
  constexpr int TT = 4, NS = 4, TTB = 4, NA = NNK_AS_NA, NSA = NNK_AS_NSA, ND = NNK_AS_ND;
  // backward-sweep scratch ring of the paired kernel: 64 frames in flight (32 when the variance rows ride along)
  constexpr int TTB_AS = 8, NSB_AS = GRAD_MODE ? 4 : 8;
  AsGeom as_geom;
  size_t as_smem = 0;
  // the gradient takes the staged path when grad_out is float32 (what autograd hands over)
  const bool paired = (MODE == MODE_FWD || (GRAD && !a.go_f64)) && !force_single_warp() && (NT <= 5) &&
                      as_geometry<TT, NA, NSA, ND, TTB_AS, NSB_AS>(GRAD ? a.go_ld * 4 : a.in_ld * ES, a.var_ld * ES, GRAD, L,
                                                                   NT, as_geom, as_smem);
  const bool staged = !force_direct_loads() && (a.win.nw == NW) &&
                      ((MODE == MODE_FWD && tma_geometry<TT, NS, TTB>(a.in_ld, a.var_ld, ES, NT, geom, smem_bytes)) ||
                       (GRAD && paired));
  for (int u0 = 0; u0 < a.n_utt; u0 += utt_per_launch) {
    const int nu = (a.n_utt - u0 < utt_per_launch) ? a.n_utt - u0 : utt_per_launch;
    p.urank0 = u0;
    if (staged) {
      constexpr bool CAN_STD = (NW == 3 && L == 1 && U == 1);
      const bool stdw = CAN_STD && is_std_windows(a.win);
      const bool varg = (a.var_ld == 0);
      const int grid = nu * p.n_groups;
      constexpr int AS_MODE = GRAD ? MODE_GRAD : MODE_FWD;  // MODE_SOLVE never gets here
      if (paired) {
#define NNK_LAUNCH_AS(STDV, VARGV)                                                                                   \
  do {                                                                                                              \
    auto kern = mlpg_fwd_as_kernel<Tin, NW, L, U, STDV, VARGV, AS_MODE, TT, NA, NSA, ND, TTB_AS, NSB_AS>;                  \
    NNK_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)as_smem));         \
    kern<<<grid, 32 * (NA + 1), as_smem, st>>>(p, as_geom);                                                         \
  } while (0)
        if (stdw && !varg) NNK_LAUNCH_AS(CAN_STD, false);
        else if (stdw && varg) NNK_LAUNCH_AS(CAN_STD, true);
        else if (!varg) NNK_LAUNCH_AS(false, false);
        else NNK_LAUNCH_AS(false, true);
#undef NNK_LAUNCH_AS
      } else {
#define NNK_LAUNCH_TMA(STDV, VARGV) \
  mlpg_fwd_tma_kernel<Tin, NW, L, U, STDV, VARGV, TT, NS, TTB><<<grid, 32, smem_bytes, st>>>(p, geom)
        if (stdw && !varg) NNK_LAUNCH_TMA(CAN_STD, false);
        else if (stdw && varg) NNK_LAUNCH_TMA(CAN_STD, true);
        else if (!varg) NNK_LAUNCH_TMA(false, false);
        else NNK_LAUNCH_TMA(false, true);
#undef NNK_LAUNCH_TMA
      }
    }
    else
      mlpg_kernel<Tin, NW, L, U, MODE, PF><<<nu * p.n_groups, 32, 0, st>>>(p);
    count_launch();
    NNK_CUDA_CHECK(cudaGetLastError());
  }
  return NNK_OK;
}

template <typename Tin, int MODE>
static int dispatch_inst(const nnk_mlpg_args_t& a, cudaStream_t st) {
  int inst = -1;
  if (pick_instance(a.win, inst) < 0) {
    set_error("unsupported window set: nw=%d (max %d) or half-width > %d", a.win.nw, NNK_MAX_WIN, NNK_MAX_HALF);
    return NNK_ERR_UNSUPPORTED;
  }
  switch (inst) {
    case 0: return launch_mlpg<Tin, 1, 0, 0, MODE>(a, st);
    case 1: return launch_mlpg<Tin, 3, 1, 1, MODE>(a, st);
    case 2: return launch_mlpg<Tin, 3, 2, 2, MODE>(a, st);
    default: return launch_mlpg<Tin, NNK_MAX_WIN, NNK_MAX_HALF, NNK_MAX_HALF, MODE>(a, st);
  }
}

static int check_args(const nnk_mlpg_args_t* a, bool grad) {
  NNK_REQUIRE(a != nullptr, NNK_ERR_ARG, "args is NULL");
  NNK_REQUIRE(a->dtype == NNK_F32 || a->dtype == NNK_F64, NNK_ERR_ARG, "dtype must be NNK_F32 or NNK_F64");
  NNK_REQUIRE(a->n_utt >= 0 && a->n_chain >= 0 && a->max_T >= 0, NNK_ERR_ARG, "negative size");
  NNK_REQUIRE(a->n_chain < (1 << 21) && a->max_T < (1 << 21) - 1 && a->n_utt < (1 << 22), NNK_ERR_ARG, "size exceeds status key range");
  if (a->n_utt == 0 || a->n_chain == 0 || a->max_T == 0) return 1;  // nothing to do
  NNK_REQUIRE(a->vars && a->out && a->utt_off && a->chains && a->status_word, NNK_ERR_ARG, "NULL device pointer");
  NNK_REQUIRE(grad ? a->grad_out != nullptr : a->means != nullptr, NNK_ERR_ARG, "NULL input pointer");
  NNK_REQUIRE(a->workspace != nullptr, NNK_ERR_WORKSPACE, "NULL workspace");
  return 0;
}

}  // namespace nnk

using namespace nnk;

extern "C" size_t nnk_mlpg_workspace_bytes(int32_t n_utt, int32_t n_chain, int32_t max_T, const nnk_windows_t* win) {
  int inst = -1;
  if (!win) return 0;
  const int S = pick_instance(*win, inst);
  if (S < 0) return 0;
  const size_t groups = (size_t)((n_chain + 31) / 32);
  return (size_t)n_utt * groups * (size_t)max_T * (size_t)(S + 1) * 32 * sizeof(double);
}

#ifdef NNK_AS_PROF
// debug builds only: read and clear the phase counters of mlpg_fwd_as_kernel (synchronises)
extern "C" int nnk_as_prof_read(unsigned long long* out16) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, nnk::g_as_prof, sizeof(unsigned long long) * 16);
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(nnk::g_as_prof, z, sizeof(z));
  return 0;
}
#endif

extern "C" int nnk_mlpg_fwd(const nnk_mlpg_args_t* a, void* stream) {
  int r = check_args(a, false);
  if (r < 0) return r;
  if (r > 0) return NNK_OK;
  DeviceGuard guard(a->out);
  cudaStream_t st = (cudaStream_t)stream;
  return a->dtype == NNK_F32 ? dispatch_inst<float, MODE_FWD>(*a, st) : dispatch_inst<double, MODE_FWD>(*a, st);
}

extern "C" int nnk_mlpg_solve(const nnk_mlpg_args_t* a, void* stream) {
  int r = check_args(a, true);
  if (r < 0) return r;
  if (r > 0) return NNK_OK;
  DeviceGuard guard(a->out);
  cudaStream_t st = (cudaStream_t)stream;
  return a->dtype == NNK_F32 ? dispatch_inst<float, MODE_SOLVE>(*a, st) : dispatch_inst<double, MODE_SOLVE>(*a, st);
}

extern "C" int nnk_mlpg_grad(const nnk_mlpg_args_t* a, void* stream) {
  int r = check_args(a, true);
  if (r < 0) return r;
  if (r > 0) return NNK_OK;
  DeviceGuard guard(a->out);
  cudaStream_t st = (cudaStream_t)stream;
  return a->dtype == NNK_F32 ? dispatch_inst<float, MODE_GRAD>(*a, st) : dispatch_inst<double, MODE_GRAD>(*a, st);
}
