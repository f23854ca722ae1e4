// nnk_mlpg_as.cuh -- warp-specialised MLPG forward kernel: NA ASSEMBLER warps and one SOLVER warp
// per (utterance, 32-chain group).
//
// ncu on the single-warp TMA kernel (profiles/r01_mlpg_v5_*): 141 instructions per frame issued by
// ONE warp at 3.6 cycles per instruction -- the batch has fewer warps (512) than the chip has warp
// schedulers (592), so nothing hides the fixed-latency dependencies.  Only ~35 of those
// instructions (the L D L^T elimination and the substitutions) are inherently serial in time; the
// rest (shared-memory loads, f32 reciprocals, widening, assembling the band row of P and b) is
// independent per frame.  This kernel splits the two:
//
//   warps A_0..A_{NA-1} (assemblers): tile k (TT frames) belongs to warp k mod NA.  Each warp stages
//                       its own tiles by TMA (rows k*TT-(NT-1) .. k*TT+TT-1, i.e. including the
//                       window halo, so tiles are self-contained), converts them to (tau, tau*mu),
//                       assembles the band rows acc[0..S] = P[t][t..t+S] and b[t] and publishes them as
//                       float64 through a shared-memory ring (PB ring, ND tiles, full/empty mbarriers);
//   warp S (solver):    consumes band rows in order, eliminates (L D L^T), forward-substitutes,
//                       writes the factor scratch, then runs the backward sweep (scratch staged by TMA).
// A first version with a single assembler was assembler-bound (the solver spun on the PB barrier,
// profiles/r01_mlpg_v6_*); assembly is parallel in time, so it gets NA warps.
#pragma once
#include "nnk_mlpg_tma.cuh"

// assemblers own PAIRS of consecutive tiles and carry the converted window halo from the first to the
// second tile of a pair in registers (10 instead of 12 conversions per 8 frames, 17 % fewer staged rows)
#ifndef NNK_AS_PAIRS
#define NNK_AS_PAIRS 0
#endif

namespace nnk {

struct AsGeom {
  uint32_t sb_in;     // bytes of one input stage (one array)
  uint32_t sb_ws;     // bytes of the factor part of one backward stage
  uint32_t sb_bw;     // bytes of one backward stage (factor tile [+ variance rows in GRAD mode])
  uint32_t ring_a;    // bytes of one assembler's input ring
  uint32_t off_pb;    // byte offset of the PB ring inside dynamic shared memory
};

// MODE_GRAD (paramgen/_mlpg.py:242-281, one banded solve + one stencil per chain instead of the
// reference's dense T x T right-hand side): the right-hand side of chain c is column c of grad_out
// (float32, staged in the slot the means occupy in MODE_FWD), and the backward sweep turns every
// solution value x[t] into the nw gradient columns of row r = t + L,
//     out[r][in_col + w * win_stride] = tau_w[r] * sum_i c[w][i] x[r - L + i],
// re-staging the variance rows next to the factor tiles (the assemblers have retired by then and
// their rings and the PB ring are free).

#ifdef NNK_AS_PROF
// A/B instrumentation (build with NNK_NVCC_EXTRA=-DNNK_AS_PROF): cycles per role and phase, summed over CTAs
__device__ unsigned long long g_as_prof[16];
#define AS_TICK(var) const long long var = clock64()
#define AS_ACC(slot, t0, t1) prof[slot] += (t1) - (t0)
#else
#define AS_TICK(var)
#define AS_ACC(slot, t0, t1)
#endif

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- factor scratch formats --------------------------------------------------------------------------------
// The S+1 numbers a frame needs in the backward sweep (z/d, l_1..l_S) make one scratch record per (frame, lane).
//   float64 inputs: plain doubles, [j][lane], (S+1) * 256 B per frame and warp.
//   float32 inputs: 48-bit records -- the upper 32 bits of the double [j][lane] followed by the next 16 bits
//                   [j][lane] (round to nearest): 36 mantissa bits, relative error 2^-37 = 7e-12, four orders
//                   below the float32 rounding of the result (6e-8) even at condition numbers of 1e4, and
//                   (S+1) * 192 B per frame and warp: the scratch round trip drops from 48 to 36 B per
//                   (frame, dim), the kernel's DRAM traffic from 76 to 64 B.  (Plain float32 records were
//                   measured and rejected in round 1: error 7e-8 on benign data, 7e-7 on ill-conditioned.)
template <typename Tin> struct WsFmt {  // float64: 8-byte records
  static constexpr int REC = 256;       // bytes per (frame, j) per warp
  static __device__ __forceinline__ void put(unsigned char* frame, int j, int lane, double v) {
    reinterpret_cast<double*>(frame + j * 256)[lane] = v;
  }
  static __device__ __forceinline__ double get(const unsigned char* frame, int nt, int j, int lane) {
    (void)nt;
    return reinterpret_cast<const double*>(frame + j * 256)[lane];
  }
};
template <> struct WsFmt<float> {       // float32 inputs: 6-byte records
  static constexpr int REC = 192;
  // global layout of a frame: [nt][32] uint32 (hi words) | [nt][32] uint16 (next 16 bits)
  static __device__ __forceinline__ void put_hi_lo(unsigned char* frame, int nt, int j, int lane, double v) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v) + 0x8000ull;  // round to nearest
    reinterpret_cast<uint32_t*>(frame)[j * 32 + lane] = (uint32_t)(bits >> 32);
    reinterpret_cast<uint16_t*>(frame + nt * 128)[j * 32 + lane] = (uint16_t)(bits >> 16);
  }
  static __device__ __forceinline__ double get(const unsigned char* frame, int nt, int j, int lane) {
    const uint32_t hi = reinterpret_cast<const uint32_t*>(frame)[j * 32 + lane];
    const uint32_t lo = reinterpret_cast<const uint16_t*>(frame + nt * 128)[j * 32 + lane];
    return __hiloint2double((int)hi, (int)(lo << 16));
  }
};
// -DNNK_AS_WS_F64 keeps 8-byte records for float32 inputs too (A/B builds)
#ifdef NNK_AS_WS_F64
template <typename Tin> using WsOf = WsFmt<double>;
#else
template <typename Tin> using WsOf = WsFmt<Tin>;
#endif
template <typename Tin>
__device__ __forceinline__ void ws_put(unsigned char* frame, int nt, int j, int lane, double v) {
  if (WsOf<Tin>::REC == 192) WsFmt<float>::put_hi_lo(frame, nt, j, lane, v);
  else WsFmt<double>::put(frame, j, lane, v);
}

template <typename Tin, int NW, int L, int U, bool STD, bool VARG, int MODE, int TT, int NA, int NSA, int ND, int TTB, int NSB>
__global__ void __launch_bounds__(32 * (NA + 1)) mlpg_fwd_as_kernel(const __grid_constant__ MlpgParams<Tin, NW, L, U> p,
                                                                    const AsGeom g) {
  constexpr int S = L + U;
  constexpr int NT = S + 1;
  constexpr int NR = S + 2;        // doubles per published band row: acc[0..S], b
  constexpr int NF = TT + NT - 1;  // frames an assembler converts per tile (TT + halo)
  constexpr int ES = (int)sizeof(Tin);
  constexpr bool GRAD = (MODE == MODE_GRAD);
  static_assert(MODE == MODE_FWD || MODE == MODE_GRAD, "staged kernel: forward or gradient");
  // PB-ring protocol invariant (root cause of the parked paired-tiles deadlock, tools/experiments/README.md):
  // an mbarrier wait only sees the PARITY of a phase, so a producer that gets two ring wraps ahead of the
  // consumer would pass its pb_empty wait on a stale phase and overwrite an undrained slot.  A producer's
  // wait for tile k proves that tile k - ND has been drained; its next tile is k + NA, which needs tile
  // k + NA - 2 ND drained to be alias-free -- guaranteed (the solver drains in order) iff NA <= ND.
  // With PAIRS (an assembler owns tiles 2q, 2q+1 and next 2q + 2 NA, 2q + 2 NA + 1) the largest stride is
  // 2 NA - 1: the round-1 attempt ran it with ND = 4 < 5, which is exactly the timing-dependent deadlock
  // that showed up under pytest / bench.py but not stand-alone.
  constexpr bool PAIRS = (NNK_AS_PAIRS != 0) && (NT > 1);
  static_assert((PAIRS ? 2 * NA - 1 : NA) <= ND, "producer tile stride must not exceed the PB ring depth (parity aliasing)");
  extern __shared__ __align__(128) unsigned char smem[];
  // barriers: input full [NA][NSA] | PB full [ND] | PB empty [ND] | scratch full [NSB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* in_full = bars;
  uint64_t* pb_full = bars + NA * NSA;
  uint64_t* pb_empty = pb_full + ND;
  uint64_t* ws_full = pb_empty + ND;
  static_assert((NA * NSA + 2 * ND + NSB) * 8 <= 512, "barrier block");
  unsigned char* rings = smem + 512;                                // NA input rings; ring 0 doubles as the scratch ring
  double* pb = reinterpret_cast<double*>(smem + g.off_pb);          // [ND][TT][NR][32]

  const int lane = threadIdx.x & 31;
  // 0..NA-1 = assemblers, NA = solver.  (Rotating the roles by the CTA index so that every SM sub-partition
  // hosts the same mix of assembler and solver warps was measured and is slower: configs[1] 0.148 -> 0.151 ms,
  // gradient 0.231 -> 0.254 ms; -DNNK_AS_ROTATE keeps the experiment buildable.)
#ifdef NNK_AS_ROTATE
  const int role = (int)(((threadIdx.x >> 5) + blockIdx.x) % (NA + 1));
#else
  const int role = threadIdx.x >> 5;
#endif
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

  if (threadIdx.x == 0) {
    for (int s = 0; s < NA * NSA; ++s) mbar_init(in_full + s, 1);
    for (int s = 0; s < NSB; ++s) mbar_init(ws_full + s, 1);
    for (int s = 0; s < ND; ++s) { mbar_init(pb_full + s, 32); mbar_init(pb_empty + s, 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int npb = (T + L + TT - 1) / TT;  // band-row tiles: tile k holds rows r = k*TT - L + j, j < TT
  const int64_t orow0 = p.out_off ? p.out_off[utt] : row0;
  Tin* const outp = reinterpret_cast<Tin*>(p.out) + orow0 * p.out_ld + ch.out_col;
  float* const outg = reinterpret_cast<float*>(p.out) + orow0 * p.out_ld + ch.in_col;  // GRAD: (T, D) float32

  // column span [cmin, cmax] of the variance (and means) rows this group touches
  const int lo_c = active ? ch.in_col : INT_MAX;
  const int hi_c = active ? ch.in_col + (solve ? (NW - 1) * ch.win_stride : 0) : -1;
  const int cmin = __reduce_min_sync(0xffffffffu, lo_c);
  const int cmax = __reduce_max_sync(0xffffffffu, hi_c);
  const int my_col = active ? ch.in_col : cmin;
  const int my_stride = solve ? ch.win_stride : 0;
  const int ldb_v = (int)(p.var_ld * ES);
  const uint64_t g_v = (uint64_t)p.vars + (uint64_t)((VARG ? 0 : row0 * p.var_ld + cmin) * ES);
  const uint32_t span_b = (uint32_t)(cmax - cmin + 1) * ES;
  int colb[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w) colb[w] = (my_col - cmin + w * my_stride) * ES;

  if (role < NA) {
    // =============================== assembler warp `role` ===========================================
    // first staged array: the means rows (FWD) or the 32 grad_out columns of this group (GRAD, float32)
    const int nlane = min(32, p.n_chain - grp * 32);
    const int ldb_m = GRAD ? (int)(p.go_ld * 4) : (int)(p.in_ld * ES);
    const uint64_t g_m = GRAD ? (uint64_t)p.go + (uint64_t)((row0 * p.go_ld + grp * 32) * 4)
                              : (uint64_t)p.means + (uint64_t)((row0 * p.in_ld + cmin) * ES);
    const uint32_t span_m = GRAD ? (uint32_t)nlane * 4u : span_b;
    const int colg = (active ? lane : 0) * 4;
    unsigned char* ring = rings + (size_t)role * g.ring_a;
    uint64_t* my_full = in_full + role * NSA;

    // Tile ownership.  PAIRS: consecutive tiles (2q, 2q+1) belong to assembler q mod NA; the second
    // tile of a pair re-uses the last NT-1 converted frames of the first one (kept in registers)
    // instead of staging and converting its window halo again: 2*TT + NT-1 conversions per pair
    // instead of 2*(TT + NT-1).  Otherwise tile k belongs to assembler k mod NA and every tile is
    // self-contained.
    constexpr int NH = (NT > 1) ? NT - 1 : 1;  // halo slots carried between the tiles of a pair
    auto next_tile = [&](int k) { return PAIRS ? (((k & 1) == 0) ? k + 1 : k - 1 + 2 * NA) : k + NA; };
    auto second = [&](int k) { return PAIRS && (k & 1); };
    // tile k stages frames [f_lo, f_hi): its TT own frames, preceded by the halo unless it is a second tile
    auto tile_flo = [&](int k) { return second(k) ? k * TT : max(0, k * TT - (NT - 1)); };
    auto tile_fhi = [&](int k) { return min(T, k * TT + TT); };
    // (only the last tile of a chain can have no frame of its own; it is then a second tile and stages nothing)

    // PREF: only pull the two ranges into L2 (issued PFD turns of this warp ahead of the real copy)
    auto issue_in = [&](auto pref_tag, int k, int s) {  // lane 0 only
      constexpr bool PREF = decltype(pref_tag)::value;
      const int f_lo = tile_flo(k), f_hi = tile_fhi(k);
      if (k >= npb || f_lo >= f_hi) return;
      const uint64_t A0 = g_m + (uint64_t)((int64_t)f_lo * ldb_m);
      const uint64_t a0 = A0 & ~(uint64_t)15;
      const uint32_t nb = (uint32_t)(((A0 + (uint64_t)((f_hi - f_lo - 1) * (int64_t)ldb_m) + span_m + 15) & ~(uint64_t)15) - a0);
      uint32_t nb2 = 0;
      uint64_t b0 = 0;
      if (!VARG) {
        const uint64_t B0 = g_v + (uint64_t)((int64_t)f_lo * ldb_v);
        b0 = B0 & ~(uint64_t)15;
        nb2 = (uint32_t)(((B0 + (uint64_t)((f_hi - f_lo - 1) * (int64_t)ldb_v) + span_b + 15) & ~(uint64_t)15) - b0);
      }
      if (PREF) {
        bulk_prefetch_l2(reinterpret_cast<const void*>(a0), nb);
        if (!VARG) bulk_prefetch_l2(reinterpret_cast<const void*>(b0), nb2);
        return;
      }
      mbar_expect_tx(my_full + s, nb + nb2);
      bulk_g2s(ring + (size_t)s * 2 * g.sb_in, reinterpret_cast<const void*>(a0), nb, my_full + s);
      if (!VARG) bulk_g2s(ring + (size_t)s * 2 * g.sb_in + g.sb_in, reinterpret_cast<const void*>(b0), nb2, my_full + s);
    };
#ifndef NNK_AS_PFD
#define NNK_AS_PFD 2
#endif
    constexpr int PFD = NNK_AS_PFD;  // L2 prefetch distance in turns of this warp beyond its staged tiles
    const int k_first = PAIRS ? 2 * role : role;
    if (lane == 0) {
      int k = k_first;
      for (int i = 0; i < NSA; ++i, k = next_tile(k)) issue_in(FullTile<false>{}, k, i);
      for (int i = 0; i < PFD; ++i, k = next_tile(k)) issue_in(FullTile<true>{}, k, 0);
    }

    double gtau[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w)
      gtau[w] = VARG ? recip_in_dtype<Tin>::f(p.vars[my_col + w * my_stride]) : 0.0;

    // the last NH converted frames of the previous tile of this warp (used by second tiles)
    double cft[NH][NW], cfm[NH][NW];
    float cfg[NH];
#pragma unroll
    for (int j = 0; j < NH; ++j) {
      cfg[j] = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) { cft[j][w] = 0.0; cfm[j][w] = 0.0; }
    }

    // convert the frames of tile k (slot j <-> frame k*TT - (NT-1) + j; staged row = frame - f_lo; SECOND:
    // slots 0 .. NT-2 come from the carry), assemble its TT band rows and publish them to PB slot `dst`
    auto do_tile = [&](auto full_tag, auto second_tag, int k, const unsigned char* sm_m, const unsigned char* sm_v,
                       double* dst, int stage) {
      constexpr bool FULL = decltype(full_tag)::value;
      constexpr bool SECOND = decltype(second_tag)::value;
      const int fbase = k * TT - (NT - 1);
      const int f_lo = SECOND ? k * TT : max(0, fbase);
      double ft[NF][NW], fm[NF][NW];
      float fg[NF];  // GRAD: grad_out of this lane's chain
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        if (SECOND && j < NT - 1) {
          fg[j] = cfg[j];
#pragma unroll
          for (int w = 0; w < NW; ++w) { ft[j][w] = cft[j][w]; fm[j][w] = cfm[j][w]; }
          continue;
        }
        const int f = fbase + j;
        const bool real = FULL || (f >= 0 && f < T);
        const bool edge = !FULL && ((m_edge == 0) || (f < m_edge) || (f >= T - m_edge));
        const int row = real ? (FULL ? (SECOND ? j - (NT - 1) : j) : f - f_lo) : 0;
        Tin mraw[NW];
        if (GRAD) {
          fg[j] = *reinterpret_cast<const float*>(sm_m + row * ldb_m + colg);
#pragma unroll
          for (int w = 0; w < NW; ++w) mraw[w] = Tin(0);
          // gradient of a pass-through column is grad_out itself
          if (j >= NT - 1 && copy_lane && real) st_stream(outg + (int64_t)f * p.out_ld, fg[j]);
        } else {
          fg[j] = 0.f;
#pragma unroll
          for (int w = 0; w < NW; ++w) mraw[w] = *reinterpret_cast<const Tin*>(sm_m + row * ldb_m + colb[w]);
          if (j >= NT - 1 && copy_lane && real) st_stream(outp + (int64_t)f * p.out_ld, mraw[0]);  // pass-through column
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          double tw;
          if (VARG) tw = gtau[w];
          else tw = recip_fast<Tin>::f(*reinterpret_cast<const Tin*>(sm_v + row * ldb_v + colb[w]));
          if (!FULL) tw = (!real || (w > 0 && edge)) ? 0.0 : tw;
          ft[j][w] = tw;
          fm[j][w] = (FULL || real) ? tw * (double)mraw[w] : 0.0;
        }
      }
      if (PAIRS && !SECOND) {
#pragma unroll
        for (int j = 0; j < NH; ++j) {
          cfg[j] = fg[TT + j];
#pragma unroll
          for (int w = 0; w < NW; ++w) { cft[j][w] = ft[TT + j][w]; cfm[j][w] = fm[TT + j][w]; }
        }
      }
      // the staged rows now live in registers: refill this stage before assembling / publishing
      __syncwarp();
      if (lane == 0) {
        int kn = k;
#pragma unroll
        for (int i = 0; i < NSA; ++i) kn = next_tile(kn);
        issue_in(FullTile<false>{}, kn, stage);
#pragma unroll
        for (int i = 0; i < PFD; ++i) kn = next_tile(kn);
        issue_in(FullTile<true>{}, kn, 0);
      }
#pragma unroll
      for (int j = 0; j < TT; ++j) {
        double acc[S + 1], bb;
        if (STD) {
          const double* a = ft[j];      // frame t-1
          const double* b = ft[j + 1];  // frame t
          const double* c = ft[j + 2];  // frame t+1
          acc[0] = b[0] + fma(0.25, a[1] + c[1], fma(4.0, b[2], a[2] + c[2]));
          acc[1] = -2.0 * (b[2] + c[2]);
          acc[2] = fma(-0.25, c[1], c[2]);
          bb = fm[j + 1][0] + fma(0.5, fm[j][1] - fm[j + 2][1], fma(-2.0, fm[j + 1][2], fm[j][2] + fm[j + 2][2]));
          if (GRAD) bb = (double)fg[j + 1];
        } else {
#pragma unroll
          for (int m = 0; m <= S; ++m) {
            double a = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w)
#pragma unroll
              for (int i = 0; i + m < NT; ++i) a = fma(p.win.q[w][m][i], ft[NT - 1 + j - i][w], a);
            acc[m] = a;
          }
          bb = 0.0;
#pragma unroll
          for (int w = 0; w < NW; ++w)
#pragma unroll
            for (int i = 0; i < NT; ++i) bb = fma(p.win.c[w][i], fm[NT - 1 + j - i][w], bb);
          if (GRAD) bb = (double)fg[j + U];  // band row j is frame fbase + U + j
        }
        double* row = dst + (size_t)j * (NR * 32) + lane;
#pragma unroll
        for (int m = 0; m <= S; ++m) row[m * 32] = acc[m];
        row[(S + 1) * 32] = bb;
      }
    };

    int s = 0;
    uint32_t par = 0;
#ifdef NNK_AS_PROF
    long long prof[4] = {0, 0, 0, 0};
#endif
    for (int k = k_first; k < npb; k = next_tile(k)) {
      const int ps = k % ND;
      const int f_lo = tile_flo(k);
      const bool staged_rows = f_lo < tile_fhi(k);
      AS_TICK(c0);
      mbar_wait_parked(pb_empty + ps, (uint32_t)(((k / ND) & 1) ^ 1));  // the solver has drained this PB slot
      AS_TICK(c1);
      if (staged_rows) mbar_wait(my_full + s, par);
      AS_TICK(c2);
      AS_ACC(0, c0, c1);
      AS_ACC(1, c1, c2);
      double* dst = pb + (size_t)ps * (TT * NR * 32);
      const uint32_t mis_m = (uint32_t)((g_m + (uint64_t)((int64_t)f_lo * ldb_m)) & 15);
      const uint32_t mis_v = (uint32_t)((g_v + (uint64_t)((int64_t)f_lo * ldb_v)) & 15);
      const unsigned char* sm_m = ring + (size_t)s * 2 * g.sb_in + mis_m;
      const unsigned char* sm_v = ring + (size_t)s * 2 * g.sb_in + g.sb_in + mis_v;
      const bool full = m_edge > 0 && k * TT - (NT - 1) >= m_edge && k * TT + TT <= T - m_edge;
      if (second(k)) {
        if (full) do_tile(FullTile<true>{}, FullTile<PAIRS>{}, k, sm_m, sm_v, dst, s);
        else do_tile(FullTile<false>{}, FullTile<PAIRS>{}, k, sm_m, sm_v, dst, s);
      } else {
        if (full) do_tile(FullTile<true>{}, FullTile<false>{}, k, sm_m, sm_v, dst, s);
        else do_tile(FullTile<false>{}, FullTile<false>{}, k, sm_m, sm_v, dst, s);
      }
      mbar_arrive(pb_full + ps);  // release: this lane's rows are visible to the solver
      if (staged_rows && ++s == NSA) { s = 0; par ^= 1; }
      AS_TICK(c3);
      AS_ACC(2, c2, c3);
    }
#ifdef NNK_AS_PROF
    if (lane == 0) for (int i = 0; i < 3; ++i) atomicAdd(g_as_prof + i, (unsigned long long)prof[i]);
#endif
    return;
  }

  // ================================= solver warp ========================================================
  constexpr int WREC = WsOf<Tin>::REC;  // scratch bytes per (frame, j) per warp
  unsigned char* const ws0 = reinterpret_cast<unsigned char*>(p.ws) + (size_t)item * ((size_t)p.max_T * NT * 256);
  unsigned char* wsp = ws0;  // the item's stride stays the float64 size: the workspace contract is unchanged
  double vcol[S + 1][S + 1], lcol[S + 1][S + 1], zz[S + 1];
#pragma unroll
  for (int k = 0; k <= S; ++k) {
    zz[k] = 0.0;
#pragma unroll
    for (int j = 0; j <= S; ++j) { vcol[k][j] = 0.0; lcol[k][j] = 0.0; }
  }
  double iv1 = 0.0;
  int bad = 0;
#ifdef NNK_AS_PROF
  long long prof[4] = {0, 0, 0, 0};
#endif

  auto eliminate = [&](int t, const double* row) {
    double acc[S + 1];
#pragma unroll
    for (int m = 0; m <= S; ++m) acc[m] = row[m * 32];
    double bb = row[(S + 1) * 32];
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
    bad = (bad == 0 && !(d > 0.0)) ? t + 1 : bad;  // linalg.pyx:79-82
    const double ivd = rcp_pos(d);
    ws_put<Tin>(wsp, NT, 0, lane, bb * ivd);
#pragma unroll
    for (int k = S; k >= 2; --k) {
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
        ws_put<Tin>(wsp, NT, j, lane, lj);
      }
      iv1 = ivd;
    }
    wsp += NT * WREC;
  };

  {
    int ps = 0;
    uint32_t ppar = 0;
    for (int k = 0; k < npb; ++k) {
      AS_TICK(c0);
      mbar_wait(pb_full + ps, ppar);
      AS_TICK(c1);
      AS_ACC(0, c0, c1);
      const double* src = pb + (size_t)ps * (TT * NR * 32) + lane;
      const int r0 = k * TT - L;  // row of the first band row in this tile
      if (r0 >= 0 && r0 + TT <= T) {
#pragma unroll
        for (int j = 0; j < TT; ++j) eliminate(r0 + j, src + (size_t)j * (NR * 32));
      } else {
#pragma unroll
        for (int j = 0; j < TT; ++j)
          if (r0 + j >= 0 && r0 + j < T) eliminate(r0 + j, src + (size_t)j * (NR * 32));
      }
      mbar_arrive(pb_empty + ps);
      if (++ps == ND) { ps = 0; ppar ^= 1; }
      AS_TICK(c2);
      AS_ACC(1, c1, c2);
    }
  }
  if (bad && solve) report_not_pd(p.status, utt, chain, bad);

  // ---- backward sweep (solver warp): y[t] = zs[t] - sum_j l_j[t] y[t+j] ---------------------------
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncwarp();
  unsigned char* ring = rings;  // every assembler has retired: reuse the input rings (and the PB ring behind them)
  const int nbt = (T + TTB - 1) / TTB;
  // backward stage kb holds frames [t0, t0 + TTB) of the factor scratch; GRAD adds the variance rows
  // [t0, min(T, t0 + TTB + L)) behind it (row r = t + L is emitted when x[t] becomes known)
  auto issue_ws = [&](int kb, int s) {
    const int t0 = (nbt - 1 - kb) * TTB;
    const uint32_t nb = (uint32_t)(min(T, t0 + TTB) - t0) * NT * WREC;
    uint32_t nb2 = 0;
    uint64_t b0 = 0;
    if (GRAD && !VARG) {
      const int r_hi = min(T, t0 + TTB + L);
      const uint64_t B0 = g_v + (uint64_t)((int64_t)t0 * ldb_v);
      b0 = B0 & ~(uint64_t)15;
      nb2 = (uint32_t)(((B0 + (uint64_t)((r_hi - t0 - 1) * (int64_t)ldb_v) + span_b + 15) & ~(uint64_t)15) - b0);
    }
    mbar_expect_tx(ws_full + s, nb + nb2);
    bulk_g2s(ring + (size_t)s * g.sb_bw, ws0 + (size_t)t0 * (NT * WREC), nb, ws_full + s);
    if (GRAD && !VARG) bulk_g2s(ring + (size_t)s * g.sb_bw + g.sb_ws, reinterpret_cast<const void*>(b0), nb2, ws_full + s);
  };
  if (lane == 0)
    for (int kb = 0; kb < NSB && kb < nbt; ++kb) issue_ws(kb, kb);
  double yw[S + 1];
#pragma unroll
  for (int j = 0; j <= S; ++j) yw[j] = 0.0;
  // output pointer walks backwards with the sweep; lanes that do not own a chain store nothing
  const int64_t ostep = p.out_ld;
  Tin* op = outp + (int64_t)(T - 1) * ostep;
  float* og = outg + (int64_t)(T - 1 + L) * ostep;  // GRAD: row t + L
  double gtau[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w)
    gtau[w] = (GRAD && VARG) ? recip_in_dtype<Tin>::f(p.vars[my_col + w * my_stride]) : 0.0;

  // GRAD: the nw gradient columns of row r from x[r - L .. r + U] = yw[0 .. S]; vrow = staged variance row r
  auto emit = [&](int r, const unsigned char* vrow) {
    const bool in = (r < T);
    const bool edge = (m_edge == 0) || (r < m_edge) || (r >= T - m_edge);
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      double sw;
      if (STD) {
        sw = (w == 0) ? yw[1] : (w == 1) ? 0.5 * (yw[2] - yw[0]) : fma(-2.0, yw[1], yw[0] + yw[2]);
      } else {
        sw = 0.0;
#pragma unroll
        for (int i = 0; i < NT; ++i) sw = fma(p.win.c[w][i], yw[i], sw);
      }
      double tw;
      if (VARG) tw = gtau[w];
      else tw = recip_fast<Tin>::f(*reinterpret_cast<const Tin*>(vrow + colb[w]));
      tw = (w > 0 && edge) ? 0.0 : tw;
      st_stream_if(og + w * my_stride, (float)(tw * sw), solve && in && (w < p.win.nw));
    }
    og -= ostep;
  };
  auto back = [&](const unsigned char* fr, const unsigned char* vrow, int r) {
#pragma unroll
    for (int j = S; j > 0; --j) yw[j] = yw[j - 1];
    // oldest terms first: only the last FMA (with y[t+1]) sits on the loop-carried chain
    double y = WsOf<Tin>::get(fr, NT, 0, lane);
#pragma unroll
    for (int j = S; j >= 1; --j) y = fma(-WsOf<Tin>::get(fr, NT, j, lane), yw[j], y);
    yw[0] = y;
    if (GRAD) {
      emit(r, vrow);
    } else {
      st_stream_if(op, (Tin)y, solve);
      op -= ostep;
    }
  };
  {
    int s = 0;
    uint32_t par = 0;
    for (int kb = 0; kb < nbt; ++kb) {
      AS_TICK(c0);
      mbar_wait(ws_full + s, par);
      AS_TICK(c1);
      AS_ACC(2, c0, c1);
      const int t0 = (nbt - 1 - kb) * TTB;
      const unsigned char* smw = ring + (size_t)s * g.sb_bw;
      // staged variance row of frame r sits at (r - t0) * ldb_v behind the factor tile
      const unsigned char* smv = ring + (size_t)s * g.sb_bw + g.sb_ws +
                                 (uint32_t)((g_v + (uint64_t)((int64_t)t0 * ldb_v)) & 15);
      if (t0 + TTB + (GRAD ? L : 0) <= T) {
#pragma unroll
        for (int j = TTB - 1; j >= 0; --j) back(smw + j * (NT * WREC), smv + (j + L) * ldb_v, t0 + j + L);
      } else {
        for (int t = T - 1; t >= t0; --t) {
          const int r = t + L;
          back(smw + (t - t0) * (NT * WREC), smv + (r < T ? r - t0 : 0) * ldb_v, r);
        }
      }
      if (GRAD && kb == nbt - 1) {
        // drain: rows L-1 .. 0 see x[-1], x[-2], ... = 0 (this is the stage of t0 == 0: row r sits at r)
        for (int r = L - 1; r >= 0; --r) {
#pragma unroll
          for (int j = S; j > 0; --j) yw[j] = yw[j - 1];
          yw[0] = 0.0;
          emit(r, smv + r * ldb_v);
        }
      }
      __syncwarp();
      if (lane == 0 && kb + NSB < nbt) issue_ws(kb + NSB, s);
      if (++s == NSB) { s = 0; par ^= 1; }
      AS_TICK(c2);
      AS_ACC(3, c1, c2);
    }
  }
#ifdef NNK_AS_PROF
  if (lane == 0) for (int i = 0; i < 4; ++i) atomicAdd(g_as_prof + 4 + i, (unsigned long long)prof[i]);
#endif
}

// row_bytes_m / row_bytes_v: bytes of one staged row of the first array (means, or grad_out in GRAD
// mode) and of the variances; half_l = L (GRAD stages TTB + L variance rows per backward tile)
template <int TT, int NA, int NSA, int ND, int TTB, int NSB>
static inline bool as_geometry(int64_t row_bytes_m, int64_t row_bytes_v, bool grad, int half_l, int nt, AsGeom& g,
                               size_t& smem_bytes) {
  const int64_t ld = row_bytes_m > row_bytes_v ? row_bytes_m : row_bytes_v;
  const size_t sb_in = ((size_t)(TT + nt - 1) * (size_t)ld + 32 + 15) / 16 * 16;
  const size_t sb_ws = (size_t)TTB * nt * 32 * 8;
  const size_t sb_var = grad ? ((size_t)(TTB + half_l) * (size_t)row_bytes_v + 32 + 15) / 16 * 16 : 0;
  const size_t sb_bw = sb_ws + sb_var;
  size_t ring_a = ((size_t)NSA * 2 * sb_in + 127) / 128 * 128;
  const size_t pbb = (size_t)ND * TT * (nt + 1) * 32 * 8;
  // the backward stages overlay the input rings and, behind them, the PB ring (both idle by then)
  const size_t bwd = (size_t)NSB * sb_bw;
  if ((size_t)NA * ring_a + pbb < bwd) ring_a = ((bwd - pbb) / NA + 127) / 128 * 128;
  const size_t tot = 512 + (size_t)NA * ring_a + pbb;
  if (tot > (size_t)100 * 1024) return false;
  g.sb_in = (uint32_t)sb_in;
  g.sb_ws = (uint32_t)sb_ws;
  g.sb_bw = (uint32_t)sb_bw;
  g.ring_a = (uint32_t)ring_a;
  g.off_pb = (uint32_t)(512 + (size_t)NA * ring_a);
  smem_bytes = tot;
  return true;
}

}  // namespace nnk
