// nnk_mlpg_tma.cuh -- the production MLPG forward kernel: same arithmetic as mlpg_kernel<MODE_FWD>
// (nnk_mlpg.cu) but every byte the sweeps consume is staged through shared memory by the TMA engine
// (cp.async.bulk, 1-D bulk copies completing on mbarriers), so the memory latency is hidden by a
// ring of tiles instead of being exposed on the scoreboard of a single warp.
//
//   forward  sweep: ring of NS tiles of TT frames of the (T, D) means / variances rows (the column
//                   span [cmin, cmax] the warp's 32 chains touch), refilled by lane 0;
//   backward sweep: ring of NS tiles of TTB frames of the warp's own float64 factor scratch.
// ncu on the register-prefetch version showed 64 % of issue slots stalled on long_scoreboard at the
// first use of every loaded frame (profiles/r01_mlpg_v1_*.txt): the hardware scoreboard has six
// counting slots per warp, so "prefetch 4 frames ahead into registers" degenerates to waiting for
// the newest load.  Bulk copies are tracked by mbarrier transaction counts, not by the scoreboard.
//
// Alignment: cp.async.bulk needs 16-byte aligned source, destination and size.  Rows of a (T, 187)
// float32 matrix are 748 bytes, so a tile generally starts 0/4/8/12 bytes past a 16-byte boundary:
// the copy is widened to the enclosing aligned range (at most 15 bytes before / after, inside the
// same cudaMalloc allocation, whose extent is 256-byte granular) and the reader adds the offset.
#pragma once
#include "nnk_mlpg.cuh"

namespace nnk {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded spin: a lost transaction must become an error, never a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (unsigned spin = 0; spin < (1u << 28); ++spin)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct TmaGeom {
  int TT, NS, TTB;       // frames per input tile, ring depth, frames per scratch tile
  uint32_t sb_in;        // bytes of one input stage (one array)
  uint32_t sb_ws;        // bytes of one scratch stage
};

template <typename Tin, int NW, int L, int U>
__global__ void __launch_bounds__(32) mlpg_fwd_tma_kernel(const __grid_constant__ MlpgParams<Tin, NW, L, U> p,
                                                          const TmaGeom g) {
  constexpr int S = L + U;
  constexpr int NT = S + 1;
  constexpr int ES = (int)sizeof(Tin);
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // [2 * NS]: forward ring, backward ring
  unsigned char* ring = smem + 128;

  const int lane = threadIdx.x;
  const int item = blockIdx.x;
  const int urank = p.urank0 + item / p.n_groups;
  const int grp = item % p.n_groups;
  const int utt = p.order ? p.order[urank] : urank;
  const int64_t row0 = p.utt_off[utt];
  const int T = p.utt_len ? p.utt_len[utt] : (int)(p.utt_off[utt + 1] - row0);
  if (T <= 0) return;
  const int chain = grp * 32 + lane;
  const bool active = chain < p.n_chain;
  nnk_chain_t ch;
  ch.in_col = 0; ch.win_stride = 0; ch.out_col = 0; ch.flags = 1;
  if (active) ch = p.chains[chain];
  const bool copy_lane = active && (ch.flags & 1);
  const bool solve = active && !(ch.flags & 1);
  const int nw = p.win.nw;
  const int m_edge = p.win.m_edge;
  const bool var_global = (p.var_ld == 0);
  const int NS = g.NS, TT = g.TT, TTB = g.TTB;

  // column span of this warp
  int lo_c = active ? ch.in_col : INT_MAX;
  int hi_c = active ? ch.in_col + (solve ? (nw - 1) * ch.win_stride : 0) : -1;
  const int cmin = __reduce_min_sync(0xffffffffu, lo_c);
  const int cmax = __reduce_max_sync(0xffffffffu, hi_c);

  if (lane == 0) {
    for (int s = 0; s < 2 * NS; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  const unsigned char* mbase = reinterpret_cast<const unsigned char*>(p.means);
  const unsigned char* vbase = reinterpret_cast<const unsigned char*>(p.vars);
  const int ntile = (T + TT - 1) / TT;

  // byte range of input tile k in `base` (row stride ld elements), widened to 16-byte alignment
  auto tile_range = [&](const unsigned char* base, int64_t ld, int k, uint64_t& a0, uint32_t& bytes, uint32_t& mis) {
    const int t0 = k * TT, t1 = min(T, t0 + TT);
    const uint64_t A0 = (uint64_t)base + (uint64_t)(((row0 + t0) * ld + cmin) * ES);
    const uint64_t A1 = (uint64_t)base + (uint64_t)(((row0 + t1 - 1) * ld + cmax + 1) * ES);
    a0 = A0 & ~(uint64_t)15;
    bytes = (uint32_t)(((A1 + 15) & ~(uint64_t)15) - a0);
    mis = (uint32_t)(A0 - a0);
  };
  auto issue_in = [&](int k) {  // lane 0 only
    const int s = k % NS;
    uint64_t a0, b0 = 0; uint32_t nb, nb2 = 0, mis;
    tile_range(mbase, p.in_ld, k, a0, nb, mis);
    if (!var_global) tile_range(vbase, p.var_ld, k, b0, nb2, mis);
    mbar_expect_tx(bars + s, nb + nb2);
    bulk_g2s(ring + (size_t)s * 2 * g.sb_in, reinterpret_cast<const void*>(a0), nb, bars + s);
    if (!var_global) bulk_g2s(ring + (size_t)s * 2 * g.sb_in + g.sb_in, reinterpret_cast<const void*>(b0), nb2, bars + s);
  };
  if (lane == 0)
    for (int k = 0; k < NS && k < ntile; ++k) issue_in(k);

  Tin gv[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w)
    gv[w] = (solve && var_global && w < nw) ? p.vars[ch.in_col + w * ch.win_stride] : Tin(1);

  double* ws = p.ws + (size_t)item * ((size_t)p.max_T * NT * 32);

  // ---- forward sweep ---------------------------------------------------------------------------
  double wt[NT][NW], wm[NT][NW];  // wt[i] = tau of frame (t + L - i)
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int w = 0; w < NW; ++w) { wt[i][w] = 0.0; wm[i][w] = 0.0; }
  double vcol[S + 1][S + 1], lcol[S + 1][S + 1], zz[S + 1];
#pragma unroll
  for (int k = 0; k <= S; ++k) {
    zz[k] = 0.0;
#pragma unroll
    for (int j = 0; j <= S; ++j) { vcol[k][j] = 0.0; lcol[k][j] = 0.0; }
  }
  double iv1 = 0.0;
  bool reported = false;
  Tin* outp = reinterpret_cast<Tin*>(p.out) + row0 * p.out_ld + ch.out_col;

  // one elimination step for row t, with the window already holding frames t-U .. t+L
  auto step = [&](int t) {
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
    double bb = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
      for (int i = 0; i < NT; ++i) bb = fma(p.win.c[w][i], wm[i][w], bb);
#pragma unroll
    for (int k = 2; k <= S; ++k) {
#pragma unroll
      for (int m = 0; m + k <= S; ++m) acc[m] = fma(-vcol[k][k + m], lcol[k][k], acc[m]);
      bb = fma(-lcol[k][k], zz[k], bb);
    }
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
    double* wsp = ws + (size_t)t * (NT * 32) + lane;
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
  };
  auto push_frame = [&](int f, const Tin(&m)[NW], const Tin(&v)[NW]) {
#pragma unroll
    for (int i = NT - 1; i > 0; --i)
#pragma unroll
      for (int w = 0; w < NW; ++w) { wt[i][w] = wt[i - 1][w]; wm[i][w] = wm[i - 1][w]; }
    const bool in = (f < T);
    const bool edge = (m_edge == 0) || (f < m_edge) || (f >= T - m_edge);
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const bool on = in && solve && (w < nw) && !(w > 0 && edge);
      const double tw = on ? recip_in_dtype<Tin>::f(v[w]) : 0.0;
      wt[0][w] = tw;
      wm[0][w] = tw * (double)m[w];
    }
  };

  for (int k = 0; k < ntile; ++k) {
    const int s = k % NS;
    mbar_wait(bars + s, (uint32_t)((k / NS) & 1));
    const int t0 = k * TT, t1 = min(T, t0 + TT);
    const unsigned char* sm_m = ring + (size_t)s * 2 * g.sb_in;
    const unsigned char* sm_v = sm_m + g.sb_in;
    const uint32_t mis_m = (uint32_t)(((uint64_t)mbase + (uint64_t)(((row0 + t0) * p.in_ld + cmin) * ES)) & 15);
    const uint32_t mis_v = var_global ? 0u : (uint32_t)(((uint64_t)vbase + (uint64_t)(((row0 + t0) * p.var_ld + cmin) * ES)) & 15);
    for (int f = t0; f < t1; ++f) {
      Tin m[NW], v[NW];
#pragma unroll
      for (int w = 0; w < NW; ++w) { m[w] = Tin(0); v[w] = Tin(1); }
      if (active) {
        const int64_t rm = (int64_t)(f - t0) * p.in_ld - cmin + ch.in_col;
        const int64_t rv = (int64_t)(f - t0) * p.var_ld - cmin + ch.in_col;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          if (w < nw && (solve || w == 0)) {
            m[w] = *reinterpret_cast<const Tin*>(sm_m + mis_m + (rm + w * ch.win_stride) * ES);
            if (solve) v[w] = var_global ? gv[w] : *reinterpret_cast<const Tin*>(sm_v + mis_v + (rv + w * ch.win_stride) * ES);
          }
        }
      }
      if (copy_lane) st_stream(outp + (int64_t)f * p.out_ld, m[0]);  // pass-through column
      push_frame(f, m, v);
      if (f >= L) step(f - L);
    }
    __syncwarp();
    if (lane == 0 && k + NS < ntile) issue_in(k + NS);
  }
  {  // drain: the last L rows see zero frames beyond the end
    Tin m[NW], v[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) { m[w] = Tin(0); v[w] = Tin(1); }
#pragma unroll
    for (int e = 0; e < L; ++e) {
      push_frame(T + e, m, v);
      if (T + e >= L) step(T + e - L);
    }
  }

  // ---- backward sweep: y[t] = zs[t] - sum_j l_j[t] y[t+j] -----------------------------------------
  // the factor scratch was written with ordinary stores; order them before the async-proxy reads
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncwarp();
  uint64_t* bbar = bars + NS;
  const int nbt = (T + TTB - 1) / TTB;
  auto issue_ws = [&](int kb) {  // lane 0 only; tiles are consumed from the last to the first
    const int idx = nbt - 1 - kb;
    const int s = kb % NS;
    const int t0 = idx * TTB, t1 = min(T, t0 + TTB);
    const uint32_t nb = (uint32_t)(t1 - t0) * NT * 32 * 8;
    mbar_expect_tx(bbar + s, nb);
    bulk_g2s(ring + (size_t)s * g.sb_ws, ws + (size_t)t0 * (NT * 32), nb, bbar + s);
  };
  if (lane == 0)
    for (int kb = 0; kb < NS && kb < nbt; ++kb) issue_ws(kb);
  double yw[S + 1];
#pragma unroll
  for (int j = 0; j <= S; ++j) yw[j] = 0.0;
  for (int kb = 0; kb < nbt; ++kb) {
    const int s = kb % NS;
    mbar_wait(bbar + s, (uint32_t)((kb / NS) & 1));
    const int idx = nbt - 1 - kb;
    const int t0 = idx * TTB, t1 = min(T, t0 + TTB);
    const double* smw = reinterpret_cast<const double*>(ring + (size_t)s * g.sb_ws) + lane;
    for (int t = t1 - 1; t >= t0; --t) {
      const double* fr = smw + (size_t)(t - t0) * (NT * 32);
#pragma unroll
      for (int j = S; j > 0; --j) yw[j] = yw[j - 1];
      double y = fr[0];
#pragma unroll
      for (int j = 1; j <= S; ++j) y = fma(-fr[j * 32], yw[j], y);
      yw[0] = y;
      if (solve) st_stream(outp + (int64_t)t * p.out_ld, (Tin)y);
    }
    __syncwarp();
    if (lane == 0 && kb + NS < nbt) issue_ws(kb + NS);
  }
}

// ring geometry for a given row stride; returns false if the rows are too wide for the staged kernel
static inline bool tma_geometry(int64_t in_ld, int64_t var_ld, int es, int nt, TmaGeom& g, size_t& smem_bytes) {
  const int64_t ld = in_ld > var_ld ? in_ld : var_ld;
  g.NS = 4;
  g.TTB = 4;
  for (int TT = 8; TT >= 2; TT /= 2) {
    g.TT = TT;
    const size_t sb_in = ((size_t)TT * (size_t)ld * es + 32 + 15) / 16 * 16;
    const size_t sb_ws = (size_t)g.TTB * nt * 32 * 8;
    const size_t fwd = (size_t)g.NS * 2 * sb_in, bwd = (size_t)g.NS * sb_ws;
    const size_t tot = 128 + (fwd > bwd ? fwd : bwd);
    if (tot <= (size_t)40 * 1024) {
      g.sb_in = (uint32_t)sb_in;
      g.sb_ws = (uint32_t)sb_ws;
      smem_bytes = tot;
      return true;
    }
  }
  return false;
}

}  // namespace nnk
