// In-register per-row top-k used by the GEMM epilogues (tvc_gemm_topk.cu, tvc_gemm_topk_pair.cu).
// Thread t of an epilogue warp owns TMEM lane t = one query row and keeps its KP best (value, column)
// pairs sorted in registers.
#pragma once
#include "tvc_internal.h"
#include "tvc_ptx.cuh"

namespace tvc {

constexpr int kStageFloats = 32 * kBM;  // slow-path staging: [32 cols][128 rows] fp32 = 16 KB

template <int KP>
struct TopList {
  float v[KP];
  int id[KP];
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      v[j] = -INFINITY;
      id[j] = -1;
    }
  }
  __device__ __forceinline__ float kth() const { return v[KP - 1]; }
  // x must be > kth().  Replace the tail and bubble up; strict '>' keeps earlier (lower index)
  // entries ahead of equal newcomers, i.e. order (value desc, index asc).
  __device__ __forceinline__ void insert(float x, int col) {
    v[KP - 1] = x;
    id[KP - 1] = col;
#pragma unroll
    for (int j = KP - 1; j > 0; --j) {
      const bool sw = v[j] > v[j - 1];
      const float a = v[j - 1], b = v[j];
      const int ia = id[j - 1], ib = id[j];
      v[j - 1] = sw ? b : a;
      v[j] = sw ? a : b;
      id[j - 1] = sw ? ib : ia;
      id[j] = sw ? ia : ib;
    }
  }
};

__device__ __forceinline__ float max32(const uint32_t (&r)[32]) {
  float m0 = fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1]));
  float m1 = fmaxf(__uint_as_float(r[2]), __uint_as_float(r[3]));
#pragma unroll
  for (int j = 4; j < 32; j += 2) {
    m0 = fmaxf(m0, __uint_as_float(r[j]));
    m1 = fmaxf(m1, __uint_as_float(r[j + 1]));
  }
  return fmaxf(m0, m1);
}

// One 128 x 256 accumulator buffer: read this thread's lane 32 columns at a time; a chunk whose
// maximum does not beat the current KP-th best costs one tcgen05.ld + a max tree; otherwise the chunk
// is staged through shared memory (keeps `top` in registers) and walked with bubble inserts.
// LOCAL_STAGE: the slow path keeps the chunk in thread-local memory (a 128-byte stack array per thread, L1
// resident) instead of the 16 KB shared-memory staging buffer - for kernels that need that shared memory.
template <int KP, int NCOLS = kBN, bool LOCAL_STAGE = false>
__device__ __forceinline__ void topk_consume_tile(TopList<KP>& top, float& thr, uint32_t t_addr, float* my_stage,
                                                  int col_base, int n_rows, long long self_col,
                                                  int debug = 0) {
#pragma unroll 1
  for (int c = 0; c < NCOLS / 32; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(t_addr + static_cast<uint32_t>(c * 32), r);
    tmem_ld_wait();
    if (max32(r) > thr && !(debug & 2)) {
      // hit mask from the registers (static indices), values through smem / local memory for the dynamic walk;
      // ascending bit order = ascending column, which the tie rule needs
      uint32_t hits = 0;
      float local_stage[LOCAL_STAGE ? 32 : 1];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (LOCAL_STAGE)
          local_stage[LOCAL_STAGE ? j : 0] = __uint_as_float(r[j]);
        else
          my_stage[j * kBM] = __uint_as_float(r[j]);
        hits |= (__uint_as_float(r[j]) > thr) ? (1u << j) : 0u;
      }
      const int col0 = col_base + c * 32;
      while (hits) {
        const int j = __ffs(hits) - 1;
        hits &= hits - 1;
        const float x = LOCAL_STAGE ? local_stage[LOCAL_STAGE ? j : 0] : my_stage[j * kBM];
        const int col = col0 + j;
        if (x > thr && col < n_rows && col != self_col) {
          top.insert(x, col);
          thr = top.kth();
        }
      }
    }
  }
}

template <int KP>
__device__ __forceinline__ void topk_store(const TopList<KP>& top, float* cand_val, int32_t* cand_idx, size_t base) {
  float4* vo = reinterpret_cast<float4*>(cand_val + base);
  int4* io = reinterpret_cast<int4*>(cand_idx + base);
#pragma unroll
  for (int j = 0; j < KP / 4; ++j) {
    vo[j] = make_float4(top.v[4 * j], top.v[4 * j + 1], top.v[4 * j + 2], top.v[4 * j + 3]);
    io[j] = make_int4(top.id[4 * j], top.id[4 * j + 1], top.id[4 * j + 2], top.id[4 * j + 3]);
  }
}

}  // namespace tvc
