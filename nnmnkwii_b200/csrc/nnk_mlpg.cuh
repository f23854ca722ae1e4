// nnk_mlpg.cuh -- declarations shared by the MLPG kernels (nnk_mlpg.cu, nnk_mlpg_tma.cuh).
#pragma once
#include "nnk_common.cuh"

namespace nnk {

enum { MODE_FWD = 0, MODE_GRAD = 1, MODE_SOLVE = 2 };

template <int NW, int L, int U>
struct WinTab {
  static constexpr int S = L + U;
  static constexpr int NT = L + U + 1;
  double c[NW][NT];         // c[w][L + k] = W_w[t, t + k]   (zero padded to the common (L, U))
  double q[NW][S + 1][NT];  // q[w][m][i] = c[w][i] * c[w][i + m]
  int nw;                   // real number of windows (<= NW)
  int m_edge;               // max_w max(l_w, u_w): dynamic windows get zero precision at that many
                            // edge frames (paramgen/_mlpg.py:177, 190-193)
};

template <typename Tin, int NW, int L, int U>
struct MlpgParams {
  const Tin* means;
  const Tin* vars;
  const void* go;
  int go_f64;
  void* out;
  int64_t in_ld, var_ld, go_ld, out_ld;
  const int64_t* utt_off;
  const int64_t* out_off;  // first output row per utterance (NULL: utt_off)
  const int32_t* utt_len;
  const int32_t* order;
  const nnk_chain_t* chains;
  int n_utt, n_chain, n_groups, max_T, urank0;
  double* ws;
  unsigned long long* status;
  WinTab<NW, L, U> win;
};

}  // namespace nnk
