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
// try_wait with a suspend-time hint: a warp that expects to wait long (a producer blocked on a full
// ring) parks instead of polling and stealing issue slots from the warp it is waiting for
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (unsigned spin = 0; spin < (1u << 24); ++spin) {
    if (mbar_try_wait_hint(bar, parity, 2000u)) return;
    __nanosleep(200);
  }
  __trap();
}
// bounded spin: a lost transaction must become an error, never a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (unsigned spin = 0; spin < (1u << 28); ++spin)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// L2 prefetch of a 16-byte aligned range (no shared memory, no completion tracking): the later bulk
// copy of the same range then pays L2 instead of HBM latency
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

struct TmaGeom {
  uint32_t sb_in;  // bytes of one input stage (one array)
  uint32_t sb_ws;  // bytes of one scratch stage
};

// reciprocal of a positive, normal double: hardware seed (MUFU.RCP64H, relative error e ~ 2^-20) and
// one cubic correction x (1 + e + e^2): error ~ e^3 < 2^-53, three dependent FMAs on the loop-carried
// chain of the elimination instead of the four of two Newton steps.  Not correctly rounded (<= 1 ulp);
// the pivots it inverts are only used inside the factorisation.
__device__ __forceinline__ double rcp_pos(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  const double e = fma(-d, x, 1.0);
  const double t = fma(e, e, e);
  return fma(x, t, x);
}

// 1 / v in the INPUT dtype like the reference (paramgen/_mlpg.py:188).  float: MUFU.RCP + one
// Newton step in FMA -- the in-range path of the IEEE-rounded __frcp_rn, without its special-case
// branch (variances are finite, normal, non-zero numbers); double: IEEE division.
template <bool B> struct FullTile { static constexpr bool value = B; };

template <typename T> struct recip_fast;
template <> struct recip_fast<float> {
  static __device__ __forceinline__ double f(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    const float e = fmaf(-v, r, 1.0f);
    r = fmaf(r, e, r);
    return (double)r;
  }
};
template <> struct recip_fast<double> {
  static __device__ __forceinline__ double f(double v) { return __drcp_rn(v); }
};

// STD: the window set is exactly static / [-0.5, 0, 0.5] / [1, -2, 1] (HTS, Merlin, the reference's
// docs and tests): the band rows are assembled from closed-form expressions instead of the generic
// coefficient tables.  VARG: global (D,) variances (var_ld == 0).  The kernel requires nw == NW.
//
// Every lane executes the same instruction stream: idle lanes (chain >= n_chain) and pass-through
// lanes read clamped in-tile addresses and solve a dummy chain whose results are never stored, so
// the hot loop has no divergent branches.  Interior tiles (TT real frames, no edge frames, no
// skipped rows) run the branch-free FULL path; the first / last tiles and the drain run the same
// code with the per-frame predicates enabled.
template <typename Tin, int NW, int L, int U, bool STD, bool VARG, int TT, int NS, int TTB>
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
  const int m_edge = p.win.m_edge;

  // column span of this warp; idle lanes are clamped onto it
  const int lo_c = active ? ch.in_col : INT_MAX;
  const int hi_c = active ? ch.in_col + (solve ? (NW - 1) * ch.win_stride : 0) : -1;
  const int cmin = __reduce_min_sync(0xffffffffu, lo_c);
  const int cmax = __reduce_max_sync(0xffffffffu, hi_c);
  const int my_col = active ? ch.in_col : cmin;
  const int my_stride = solve ? ch.win_stride : 0;

  if (lane == 0) {
    for (int s = 0; s < 2 * NS; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  const int ntile = (T + TT - 1) / TT;
  const int ldb_m = (int)(p.in_ld * ES), ldb_v = (int)(p.var_ld * ES);  // row strides in bytes
  // global byte address of (frame 0, column cmin) of this utterance
  const uint64_t g_m = (uint64_t)p.means + (uint64_t)((row0 * p.in_ld + cmin) * ES);
  const uint64_t g_v = (uint64_t)p.vars + (uint64_t)((VARG ? 0 : row0 * p.var_ld + cmin) * ES);
  const uint32_t span_b = (uint32_t)(cmax - cmin + 1) * ES;

  auto issue_in = [&](int k, int s) {  // lane 0 only: tile k -> stage s
    const int nfr = min(TT, T - k * TT);
    const uint64_t A0 = g_m + (uint64_t)((int64_t)k * TT * ldb_m);
    const uint64_t a0 = A0 & ~(uint64_t)15;
    const uint32_t nb = (uint32_t)(((A0 + (uint64_t)((nfr - 1) * (int64_t)ldb_m) + span_b + 15) & ~(uint64_t)15) - a0);
    uint32_t nb2 = 0;
    uint64_t b0 = 0;
    if (!VARG) {
      const uint64_t B0 = g_v + (uint64_t)((int64_t)k * TT * ldb_v);
      b0 = B0 & ~(uint64_t)15;
      nb2 = (uint32_t)(((B0 + (uint64_t)((nfr - 1) * (int64_t)ldb_v) + span_b + 15) & ~(uint64_t)15) - b0);
    }
    mbar_expect_tx(bars + s, nb + nb2);
    bulk_g2s(ring + (size_t)s * 2 * g.sb_in, reinterpret_cast<const void*>(a0), nb, bars + s);
    if (!VARG) bulk_g2s(ring + (size_t)s * 2 * g.sb_in + g.sb_in, reinterpret_cast<const void*>(b0), nb2, bars + s);
  };
  if (lane == 0)
    for (int k = 0; k < NS && k < ntile; ++k) issue_in(k, k);

  // global variances: tau is constant over time (up to the edge rule)
  double gtau[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w)
    gtau[w] = VARG ? recip_in_dtype<Tin>::f(p.vars[my_col + w * my_stride]) : 0.0;

  double* const ws0 = p.ws + (size_t)item * ((size_t)p.max_T * NT * 32);
  double* wsp = ws0 + lane;
  const int64_t orow0 = p.out_off ? p.out_off[utt] : row0;
  Tin* const outp = reinterpret_cast<Tin*>(p.out) + orow0 * p.out_ld + ch.out_col;

  // ---- forward sweep ---------------------------------------------------------------------------
  double vcol[S + 1][S + 1], lcol[S + 1][S + 1], zz[S + 1];
#pragma unroll
  for (int k = 0; k <= S; ++k) {
    zz[k] = 0.0;
#pragma unroll
    for (int j = 0; j <= S; ++j) { vcol[k][j] = 0.0; lcol[k][j] = 0.0; }
  }
  double iv1 = 0.0;
  int bad = 0;  // 1-based frame of the first non-positive pivot of this chain

  // eliminate row t given its assembled band row acc[m] = P[t][t+m] and right-hand side bb
  auto eliminate = [&](int t, double(&acc)[S + 1], double bb) {
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
    bad = (bad == 0 && !(d > 0.0)) ? t + 1 : bad;  // linalg.pyx:79-82, reported after the sweep
    const double ivd = rcp_pos(d);
    wsp[0] = bb * ivd;
#pragma unroll
    for (int k = S; k >= 2; --k) {  // only the entries later rows still need are carried
      zz[k] = zz[k - 1];
#pragma unroll
      for (int j = k; j <= S; ++j) { vcol[k][j] = vcol[k - 1][j]; lcol[k][j] = lcol[k - 1][j]; }
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
    wsp += NT * 32;
  };

  // carry[i] = frame (t0 - (NT-1) + i) of the previous tile (oldest first)
  constexpr int NC = NT - 1 > 0 ? NT - 1 : 1;
  double cft[NC][NW], cfm[NC][NW];
#pragma unroll
  for (int i = 0; i < NC; ++i)
#pragma unroll
    for (int w = 0; w < NW; ++w) { cft[i][w] = 0.0; cfm[i][w] = 0.0; }

  // byte offsets of this lane's NW columns inside a staged row
  int colb[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w) colb[w] = (my_col - cmin + w * my_stride) * ES;

  // A tile is processed in three phases so that the independent work of TT frames (loads,
  // reciprocals, assembly of the band rows) is issued back to back and only the short elimination
  // recurrence is serial: (1) frames -> (tau, tau*mu); (2) band rows of P and b; (3) eliminate.
  // FULL: TT real interior frames (no edge rule, no skipped rows); otherwise nfr real frames are
  // staged and nproc frames are consumed (zeros past the end).
  auto do_tile = [&](auto full_tag, int t0, int nfr, int nproc, const unsigned char* sm_m, const unsigned char* sm_v) {
    constexpr bool FULL = decltype(full_tag)::value;
    double ft[TT + NT - 1][NW], fm[TT + NT - 1][NW];
#pragma unroll
    for (int i = 0; i < NT - 1; ++i)
#pragma unroll
      for (int w = 0; w < NW; ++w) { ft[i][w] = cft[i][w]; fm[i][w] = cfm[i][w]; }
    // phase 1
#pragma unroll
    for (int j = 0; j < TT; ++j) {
      const int f = t0 + j;
      const bool real = FULL || (j < nfr);
      const bool edge = !FULL && ((m_edge == 0) || (f < m_edge) || (f >= T - m_edge));
      const int jj = real ? j : 0;  // keep the address inside the stage for frames past the end
      Tin mraw[NW];
#pragma unroll
      for (int w = 0; w < NW; ++w) mraw[w] = *reinterpret_cast<const Tin*>(sm_m + jj * ldb_m + colb[w]);
      if (copy_lane && real) st_stream(outp + (int64_t)f * p.out_ld, mraw[0]);  // pass-through column
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        double tw;
        if (VARG) tw = gtau[w];
        else tw = recip_fast<Tin>::f(*reinterpret_cast<const Tin*>(sm_v + jj * ldb_v + colb[w]));
        if (!FULL) tw = (!real || (w > 0 && edge)) ? 0.0 : tw;
        ft[NT - 1 + j][w] = tw;
        fm[NT - 1 + j][w] = tw * (double)mraw[w];
      }
    }
    // phase 2: row (t0 + j - L) sees frame (t0 + j - i) in window slot i  ->  ft[NT-1 + j - i]
    double acc[TT][S + 1], bb[TT];
#pragma unroll
    for (int j = 0; j < TT; ++j) {
      if (STD) {
        // static / delta [-0.5, 0, 0.5] / delta-delta [1, -2, 1]: the band row in closed form
        const double* a = ft[j];      // frame t-1
        const double* b = ft[j + 1];  // frame t
        const double* c = ft[j + 2];  // frame t+1
        acc[j][0] = b[0] + fma(0.25, a[1] + c[1], fma(4.0, b[2], a[2] + c[2]));
        acc[j][1] = -2.0 * (b[2] + c[2]);
        acc[j][2] = fma(-0.25, c[1], c[2]);
        bb[j] = fm[j + 1][0] + fma(0.5, fm[j][1] - fm[j + 2][1], fma(-2.0, fm[j + 1][2], fm[j][2] + fm[j + 2][2]));
      } else {
#pragma unroll
        for (int m = 0; m <= S; ++m) {
          double a = 0.0;
#pragma unroll
          for (int w = 0; w < NW; ++w)
#pragma unroll
            for (int i = 0; i + m < NT; ++i) a = fma(p.win.q[w][m][i], ft[NT - 1 + j - i][w], a);
          acc[j][m] = a;
        }
        double r = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w)
#pragma unroll
          for (int i = 0; i < NT; ++i) r = fma(p.win.c[w][i], fm[NT - 1 + j - i][w], r);
        bb[j] = r;
      }
    }
    // phase 3
#pragma unroll
    for (int j = 0; j < TT; ++j)
      if (FULL || (j < nproc && t0 + j >= L)) eliminate(t0 + j - L, acc[j], bb[j]);
    // carry the last NT-1 frames
#pragma unroll
    for (int i = 0; i < NT - 1; ++i)
#pragma unroll
      for (int w = 0; w < NW; ++w) { cft[i][w] = ft[TT + i][w]; cfm[i][w] = fm[TT + i][w]; }
  };

  {
    int s = 0;
    uint32_t par = 0;
    uint32_t mis_m = (uint32_t)(g_m & 15), mis_v = (uint32_t)(g_v & 15);  // misalignment of the current tile
    const uint32_t dmis_m = (uint32_t)(TT * ldb_m) & 15, dmis_v = (uint32_t)(TT * ldb_v) & 15;
    const int full_lo = max(L, m_edge), full_hi = (m_edge > 0) ? T - m_edge : -1;
    for (int k = 0; k < ntile; ++k) {
      mbar_wait(bars + s, par);
      const int t0 = k * TT;
      const unsigned char* sm_m = ring + (size_t)s * 2 * g.sb_in + mis_m;
      const unsigned char* sm_v = ring + (size_t)s * 2 * g.sb_in + g.sb_in + mis_v;
      if (t0 >= full_lo && t0 + TT <= full_hi)
        do_tile(FullTile<true>{}, t0, TT, TT, sm_m, sm_v);
      else
        do_tile(FullTile<false>{}, t0, min(TT, T - t0), min(TT, T + L - t0), sm_m, sm_v);
      __syncwarp();
      if (lane == 0 && k + NS < ntile) issue_in(k + NS, s);
      mis_m = (mis_m + dmis_m) & 15;
      mis_v = (mis_v + dmis_v) & 15;
      if (++s == NS) { s = 0; par ^= 1; }
    }
    // drain: the last L rows see only zero frames beyond the end
    if (ntile * TT < T + L) do_tile(FullTile<false>{}, ntile * TT, 0, T + L - ntile * TT, ring, ring);
  }
  if (bad && solve) report_not_pd(p.status, utt, chain, bad);

  // ---- backward sweep: y[t] = zs[t] - sum_j l_j[t] y[t+j] -----------------------------------------
  // the factor scratch was written with ordinary stores; order them before the async-proxy reads
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncwarp();
  uint64_t* bbar = bars + NS;
  const int nbt = (T + TTB - 1) / TTB;
  auto issue_ws = [&](int kb, int s) {  // lane 0 only; tiles are consumed from the last to the first
    const int t0 = (nbt - 1 - kb) * TTB;
    const uint32_t nb = (uint32_t)(min(T, t0 + TTB) - t0) * NT * 32 * 8;
    mbar_expect_tx(bbar + s, nb);
    bulk_g2s(ring + (size_t)s * g.sb_ws, ws0 + (size_t)t0 * (NT * 32), nb, bbar + s);
  };
  if (lane == 0)
    for (int kb = 0; kb < NS && kb < nbt; ++kb) issue_ws(kb, kb);
  double yw[S + 1];
#pragma unroll
  for (int j = 0; j <= S; ++j) yw[j] = 0.0;
  auto back = [&](int t, const double* fr) {
#pragma unroll
    for (int j = S; j > 0; --j) yw[j] = yw[j - 1];
    double y = fr[0];
#pragma unroll
    for (int j = 1; j <= S; ++j) y = fma(-fr[j * 32], yw[j], y);
    yw[0] = y;
    if (solve) st_stream(outp + (int64_t)t * p.out_ld, (Tin)y);
  };
  {
    int s = 0;
    uint32_t par = 0;
    for (int kb = 0; kb < nbt; ++kb) {
      mbar_wait(bbar + s, par);
      const int t0 = (nbt - 1 - kb) * TTB;
      const double* smw = reinterpret_cast<const double*>(ring + (size_t)s * g.sb_ws) + lane;
      if (t0 + TTB <= T) {
#pragma unroll
        for (int j = TTB - 1; j >= 0; --j) back(t0 + j, smw + j * (NT * 32));
      } else {
        for (int t = T - 1; t >= t0; --t) back(t, smw + (t - t0) * (NT * 32));
      }
      __syncwarp();
      if (lane == 0 && kb + NS < nbt) issue_ws(kb + NS, s);
      if (++s == NS) { s = 0; par ^= 1; }
    }
  }
}

// ring geometry for a given row stride; returns false if the rows are too wide for the staged kernel
template <int TT, int NS, int TTB>
static inline bool tma_geometry(int64_t in_ld, int64_t var_ld, int es, int nt, TmaGeom& g, size_t& smem_bytes) {
  const int64_t ld = in_ld > var_ld ? in_ld : var_ld;
  const size_t sb_in = ((size_t)TT * (size_t)ld * es + 32 + 15) / 16 * 16;
  const size_t sb_ws = (size_t)TTB * nt * 32 * 8;
  const size_t fwd = (size_t)NS * 2 * sb_in, bwd = (size_t)NS * sb_ws;
  const size_t tot = 128 + (fwd > bwd ? fwd : bwd);
  if (tot > (size_t)40 * 1024) return false;
  g.sb_in = (uint32_t)sb_in;
  g.sb_ws = (uint32_t)sb_ws;
  smem_bytes = tot;
  return true;
}

}  // namespace nnk
