// Persistent, warp-specialised bf16 GEMM for sm_100a:  C = epilogue(alpha * A * B^T)
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (accumulators live in TMEM, 2 buffers)
//   warps 2..9  : epilogue (tcgen05.ld -> registers -> bias/activation/residual/RoPE/SwiGLU -> global); two warps per
//                 TMEM lane quarter, each taking half of the tile's columns (the epilogue is issue-bound: with one
//                 warp per SM sub-partition short-K GEMMs (CLIP / Whisper) were epilogue-limited)
//
// Epilogue note (measured, round 1): staging C through shared memory for fully coalesced stores was tried and is
// SLOWER (LLaMA GEMMs 1384 -> 1240 TFLOP/s, CLIP 315 -> 228): with cta_group::1 and a 128x256 tile the tensor core
// already reads 96 B/cycle/SM of the 128 B/cycle shared-memory bandwidth, so any extra smem traffic stalls the MMA.
// Accumulator rows therefore go registers -> global directly (each thread owns one row: 64 B contiguous per chunk).
//
// Tile = 128 (M) x BN (N) x 64 (K) per stage; UMMA shape 128 x BN x 16.  Two TMEM accumulator buffers let the
// epilogue of tile i overlap the main loop of tile i+1.  One CTA per SM, static round-robin tile schedule with
// grouped rasterisation for L2 reuse.
//
// Wide, multi-wave problems (the LLaMA GEMMs) run as CTA PAIRS: a 2-CTA cluster is ONE tcgen05.mma.cta_group::2 unit on a
// 256 x 256 tile (template parameter CG2; see the comment at the kernel) — or, for MN-major B, two cta_group::1 MMAs that
// share a TMA-multicast B tile (MC).  A partial last wave can be split along K over all CTAs (stream-K tail, gemm_work()).
//
// Thin problems (decode steps) are launched with the operands swapped and `c_trans` set: the weight matrix takes the
// 128-row A side, the few activation rows the narrow B side, and the standard epilogue stores transposed.
//
// Reference call sites replaced: see include/macaw_b200.h (mm_gemm_fwd).
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/macaw_b200.h"
#include <stdlib.h>

namespace mm {

struct GemmKParams {
  int M, N, K, batch, batch2;
  int m_tiles, n_tiles, num_k;
  int b_shared, b2_shared;
  void* C;
  long long ldc, c_bs, c_bs2;
  int c_fp32;
  int act;
  float alpha;
  const bf16* bias;
  long long bias_bs;
  const float* row_scale;
  const bf16* residual;
  long long ldr, r_bs, r_bs2;
  int res_row_mod;
  const float* rope_cos;
  const float* rope_sin;
  int rope_T, rope_cols;
  const int* rope_pos;
  int c_trans;
  int c_fp16;
  int aux_f16;     // bias / bias2 / residual tensors are fp16 (the calling thread's activation format), else bf16
  uint32_t idesc;  // tcgen05 instruction descriptor (operand formats are run-time properties)
  const float* bias_rs;
  const bf16* bias2;
  const float* bias2_rs;
  float* sumsq_out;        // [M][sumsq_parts]: per-(row, 32-column chunk) sums of squares of the STORED outputs
  const float* rs_sumsq;   // [M][rs_parts]: row_scale = rsqrt(sum_j rs_sumsq[row][j] / K + rs_eps) (RMSNorm of the A rows)
  int sumsq_parts, rs_parts;
  float rs_eps;
  int vec_ok;
  int group_m;  // rasterisation: M units per group
  // stream-K tail (0 tiles = off): the last sk_tiles tiles (indices >= sk_first) are split along K into equal shares,
  // one per CTA; partial accumulators travel through sk_ws (fp32, one 128 x BN slot per CTA), sk_flags signals them
  int sk_tiles, sk_first;
  float* sk_ws;
  int* sk_flags;
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr uint32_t kABytes = kBlockM * kBlockK * 2;  // 16 KiB per stage

__host__ __device__ constexpr int gemm_stages(int BN) {
  // keep the ring within ~200 KiB
  return BN >= 256 ? 4 : (BN >= 128 ? 6 : 8);
}
__host__ __device__ constexpr uint32_t gemm_tmem_cols(int BN) {
  return 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
}
__host__ __device__ constexpr size_t gemm_smem_bytes(int BN) {
  return 1024 /*align slack*/ + (size_t)gemm_stages(BN) * (kABytes + BN * kBlockK * 2) + 256 /*barriers*/;
}

struct TileCoord {
  int b, b_lo, b_hi, m_blk, n_blk;  // b = b_hi * batch + b_lo
};
__device__ __forceinline__ TileCoord tile_coord(int idx, const GemmKParams& p, int m_units) {
  const int per_batch = m_units * p.n_tiles;
  TileCoord t;
  t.b = idx / per_batch;
  t.b_hi = t.b / p.batch;
  t.b_lo = t.b - t.b_hi * p.batch;
  int r = idx - t.b * per_batch;
  const int in_group = p.group_m * p.n_tiles;
  const int g = r / in_group;
  const int first_m = g * p.group_m;
  const int gsz = min(m_units - first_m, p.group_m);
  const int rr = r - g * in_group;
  t.m_blk = first_m + rr % gsz;
  t.n_blk = rr / gsz;
  return t;
}

// One unit of a CTA's schedule: a whole tile (role 0), or — in the stream-K tail — a K-range of a tile whose partial
// accumulator is handed over (role 1, "contributor") or which ends the tile and folds the others' partials in before
// the epilogue (role 2, "finisher"; contributors are the CTAs c0 .. c0 + nc - 1, each with exactly one slot).
//
// Tail schedule: the sk_tiles * num_k k-blocks of the tail are cut into n_workers equal contiguous shares (share w =
// [w U / W, (w + 1) U / W)).  A share is shorter than one tile's K extent (sk_tiles < n_workers), so it touches at most
// two tiles: it may END one tile (finisher piece) and BEGIN the next (contributor piece), or sit inside one tile
// (contributor).  Every CTA runs its contributor piece FIRST and never waits in it; finishers wait only for
// contributors, so there is no cycle — and all CTAs are co-resident (grid <= SM count, one CTA per SM).
struct GemmWork {
  int tile, kb0, kb1, role, c0, nc;
};
__device__ __forceinline__ bool gemm_work(const GemmKParams& p, int worker, int n_workers, int total_tiles, int it,
                                          GemmWork& w) {
  const int first = p.sk_tiles > 0 ? p.sk_first : total_tiles;
  const int n_full = first > worker ? (first - worker + n_workers - 1) / n_workers : 0;
  w.c0 = 0;
  w.nc = 0;
  // The tail pieces come FIRST (contributor, then finisher), the CTA's full tiles after them: the hand-over (partial
  // accumulators through L2, flag latency, the finisher's extra loads) then overlaps the main loops of the full tiles
  // instead of sitting exposed at the end of the kernel.
  int n_sk = 0;
  unsigned nk = 1, U = 0, u0 = 0, u1 = 0, ta = 0, end_a = 0;
  bool has_b = false;
  if (p.sk_tiles > 0) {
    // 32-bit arithmetic (the host checks sk_tiles * num_k * (n_workers + 1) < 2^31): no 64-bit divisions on the tile path
    nk = p.num_k;
    U = static_cast<unsigned>(p.sk_tiles) * nk;
    u0 = static_cast<unsigned>(worker) * U / n_workers;
    u1 = static_cast<unsigned>(worker + 1) * U / n_workers;
    if (u0 < u1) {
      ta = u0 / nk;
      end_a = u1 < (ta + 1) * nk ? u1 : (ta + 1) * nk;
      has_b = u1 > end_a;  // the share runs on into tile ta + 1 (then piece A ends tile ta)
      n_sk = has_b ? 2 : 1;
    }
  }
  if (it >= n_sk) {
    const int f = it - n_sk;
    if (f >= n_full) return false;
    w.tile = worker + f * n_workers;
    w.kb0 = 0;
    w.kb1 = p.num_k;
    w.role = 0;
    return true;
  }
  if (has_b && it == 0) {  // contributor piece first
    w.tile = first + static_cast<int>(ta) + 1;
    w.kb0 = 0;
    w.kb1 = static_cast<int>(u1 - end_a);
    w.role = 1;
    return true;
  }
  w.tile = first + static_cast<int>(ta);
  w.kb0 = static_cast<int>(u0 - ta * nk);
  w.kb1 = static_cast<int>(end_a - ta * nk);
  w.role = (w.kb1 == p.num_k) ? 2 : 1;
  if (w.role == 2) {  // contributors: the CTAs below this one whose shares reach into tile ta
    int c0 = worker;
    while (c0 > 0 && static_cast<unsigned>(c0) * U / n_workers > ta * nk) --c0;
    w.c0 = c0;
    w.nc = worker - c0;
  }
  return true;
}

__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid(x) = 0.5 tanh(0.5 x) + 0.5: one MUFU op (tanh.approx.f32, rel. error ~2^-11 — far below the bf16 output rounding)
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
// exact-GELU's erf via Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7): 2 MUFU + ~10 FMA, branch-free
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = 1.0f - poly * exp2f(-1.4426950408889634f * z * z);  // erf(|x| / sqrt 2)
  return 0.5f * x + 0.5f * fabsf(x) * e;                                 // 0.5 x (1 + sign(x) erf(|x|/sqrt2))
}

// bias / residual elements in the run-time activation format
__device__ __forceinline__ float aux_ld(const bf16* p, long long i, int f16) {
  const uint16_t raw = reinterpret_cast<const uint16_t*>(p)[i];
  return f16 ? __half2float(__ushort_as_half(raw)) : __uint_as_float(static_cast<uint32_t>(raw) << 16);
}
__device__ __forceinline__ void aux_add8(float* v, const uint4& u, int f16) {
  if (f16) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(h[k]);
      v[2 * k] += f.x;
      v[2 * k + 1] += f.y;
    }
  } else {
    v[0] += bf16lo(u.x); v[1] += bf16hi(u.x); v[2] += bf16lo(u.y); v[3] += bf16hi(u.y);
    v[4] += bf16lo(u.z); v[5] += bf16hi(u.z); v[6] += bf16lo(u.w); v[7] += bf16hi(u.w);
  }
}

// Activation over a 32-wide chunk: the switch is hoisted so each case is a straight unrolled loop.
__device__ __forceinline__ void apply_act32(float (&v)[32], int act) {
  if (act == MM_ACT_QUICK_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= sigmoid_fast(1.702f * v[i]);
  } else if (act == MM_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
  } else if (act == MM_ACT_SILU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= sigmoid_fast(v[i]);
  }
}

// Store 32 consecutive outputs of one row (columns col0..col0+31), masked by N.
__device__ __forceinline__ void store_row32(const GemmKParams& p, void* crow, int col0, int ncols_total,
                                            const float (&v)[32]) {
  if (p.c_fp32) {
    float* c = reinterpret_cast<float*>(crow) + col0;
    if (p.vec_ok && col0 + 32 <= ncols_total) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        reinterpret_cast<float4*>(c)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < ncols_total) c[i] = v[i];
    }
  } else {
    bf16* c = reinterpret_cast<bf16*>(crow) + col0;
    if (p.vec_ok && col0 + 32 <= ncols_total) {
      if (p.c_fp16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_f16x2(v[8 * i + 0], v[8 * i + 1]);
          u.y = pack_f16x2(v[8 * i + 2], v[8 * i + 3]);
          u.z = pack_f16x2(v[8 * i + 4], v[8 * i + 5]);
          u.w = pack_f16x2(v[8 * i + 6], v[8 * i + 7]);
          reinterpret_cast<uint4*>(c)[i] = u;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
          u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
          u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
          u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
          reinterpret_cast<uint4*>(c)[i] = u;
        }
      }
    } else if (p.c_fp16) {
      __half* ch = reinterpret_cast<__half*>(c);
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < ncols_total) ch[i] = __float2half_rn(v[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < ncols_total) c[i] = __float2bfloat16(v[i]);
    }
  }
}

// MC = true: clusters of 2 CTAs work on vertically adjacent tiles (m_blk = 2u, 2u+1; same n_blk) and share the B tile:
// each CTA TMA-loads half of it and multicasts to both, cutting L2->SM operand traffic per CTA from 48 to 32 KiB per
// k-block.  A smem slot is free only when BOTH CTAs' MMAs have retired (they both receive the peer's multicast).
// A_MN = true: A is given as [K][M] with M contiguous (the transpose of a row-major [tokens][features] activation): the
// weight-gradient GEMM dW = dY^T X reads dY and X exactly as the forward pass wrote them, no transpose copies.
// CG2 = true (with MC): the pair runs as ONE cta_group::2 MMA unit — 256 x BN tile, each CTA stages its own 128 rows of A
// and HALF of the B tile (no multicast: the tensor core reads both halves across the pair), so a CTA's shared memory
// delivers 32 KiB per k-block instead of 48 and the ring is 6 stages deep.  Only the leader (rank 0) issues MMAs; both
// producers credit their loads to the leader's full barrier; commits are multicast to both CTAs' barriers; the epilogue
// warps of both CTAs release the accumulator on the leader's barrier.
template <int BN, int EPI, bool B_MN, bool MC, bool A_MN = false, bool CG2 = false>
__global__ void __launch_bounds__(320, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmKParams p) {
  static_assert(!CG2 || (MC && !A_MN && BN == 256), "cta_group::2 variant: pair schedule, K-major A, BN 256");
  constexpr int STAGES = CG2 ? 6 : gemm_stages(BN);
  constexpr uint32_t B_BYTES = (CG2 ? BN / 2 : BN) * kBlockK * 2;
  constexpr uint32_t TMEM_COLS = gemm_tmem_cols(BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = MC ? static_cast<int>(cluster_ctarank()) : 0;
  const int worker = MC ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int n_workers = MC ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int m_units = MC ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const int total_tiles = p.batch * p.batch2 * m_units * p.n_tiles;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);                          // CG2: only the LEADER's barrier is used (see the producer)
        mbar_init(&empty_bar[s], CG2 ? 1 : (MC ? 2 : 1));    // CG2: one multicast commit from the leader
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tfull_bar[s], 1);
        mbar_init(&tempty_bar[s], CG2 ? 16 : 8);             // CG2: the epilogue warps of BOTH CTAs, on the leader's barrier
      }
      fence_mbar_init();
    }
    __syncwarp();
    if constexpr (CG2) {
      tmem_alloc2(tmem_slot, TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (MC) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // prologue done (barriers, TMEM, descriptors): let the next kernel start its own, then wait for our inputs
  griddep_launch();
  griddep_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      GemmWork wk;
      for (int it = 0; gemm_work(p, worker, n_workers, total_tiles, it, wk); ++it) {
        TileCoord t = tile_coord(wk.tile, p, m_units);
        if constexpr (MC) t.m_blk = 2 * t.m_blk + cta_rank;
        const int bb = p.b_shared ? 0 : t.b_lo;
        const int bh = p.b2_shared ? 0 : t.b_hi;
        for (int kb = wk.kb0; kb < wk.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if constexpr (CG2) {
            // Both CTAs' loads credit the LEADER's barrier; the leader alone announces the bytes of the pair (a remote
            // arrive.expect_tx per k-block from the peer is a cluster-scope release in the producer's critical loop:
            // measured 767 instead of 1483 TFLOP/s).  The peer's bytes may land before the announcement: the phase cannot
            // complete until the leader's arrival is in.
            const uint32_t lfull = mapa_u32(smem_u32(&full_bar[stage]), 0);
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (kABytes + B_BYTES));
            tma_load_4d_cg2(&tmA, lfull, sA + stage * kABytes, kb * kBlockK, t.m_blk * kBlockM, t.b_lo, t.b_hi);
            if constexpr (!B_MN) {
              tma_load_4d_cg2(&tmB, lfull, sB + stage * B_BYTES, kb * kBlockK, t.n_blk * BN + cta_rank * (BN / 2), bb, bh);
            } else {  // MN-major B: this CTA's half of the N extent = BN / 128 boxes of 64 columns x 64 k
#pragma unroll
              for (int jj = 0; jj < BN / 128; ++jj)
                tma_load_4d_cg2(&tmB, lfull, sB + stage * B_BYTES + jj * 8192,
                                t.n_blk * BN + (cta_rank * (BN / 128) + jj) * 64, kb * kBlockK, bb, bh);
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], kABytes + B_BYTES);
          if constexpr (A_MN) {
#pragma unroll
            for (int j = 0; j < kBlockM / 64; ++j)
              tma_load_4d(&tmA, &full_bar[stage], sA + stage * kABytes + j * 8192, t.m_blk * kBlockM + j * 64,
                          kb * kBlockK, t.b_lo, t.b_hi);
          } else {
            tma_load_4d(&tmA, &full_bar[stage], sA + stage * kABytes, kb * kBlockK, t.m_blk * kBlockM, t.b_lo, t.b_hi);
          }
          if constexpr (MC) {
            // this CTA fetches its half of the B tile and multicasts it to both CTAs of the pair
            if constexpr (!B_MN) {
              tma_load_4d_mc(&tmB, &full_bar[stage], sB + stage * B_BYTES + cta_rank * (BN / 2) * 128, kb * kBlockK,
                             t.n_blk * BN + cta_rank * (BN / 2), bb, bh, 3);
            } else {
#pragma unroll
              for (int jj = 0; jj < BN / 128; ++jj) {
                const int j = cta_rank * (BN / 128) + jj;
                tma_load_4d_mc(&tmB, &full_bar[stage], sB + stage * B_BYTES + j * 8192, t.n_blk * BN + j * 64,
                               kb * kBlockK, bb, bh, 3);
              }
            }
          } else if constexpr (!B_MN) {
            tma_load_4d(&tmB, &full_bar[stage], sB + stage * B_BYTES, kb * kBlockK, t.n_blk * BN, bb, bh);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_4d(&tmB, &full_bar[stage], sB + stage * B_BYTES + j * 8192, t.n_blk * BN + j * 64,
                          kb * kBlockK, bb, bh);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (CG2: the pair's leader only)
    if ((!CG2 || cta_rank == 0) && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      GemmWork wk;
      for (int it = 0; gemm_work(p, worker, n_workers, total_tiles, it, wk); ++it) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = wk.kb0; kb < wk.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * kABytes);
          const uint32_t b_addr = smem_u32(sB + stage * B_BYTES);
          // K-major operand: 8-row groups are 1024 B apart (SBO); LBO unused with 128B swizzle.
          const uint64_t a_desc = A_MN ? make_sdesc_sw128(a_addr, 8192, 1024) : make_sdesc_sw128(a_addr, 16, 1024);
          // MN-major operand: 64-column blocks are 8192 B apart (LBO); 8-k groups 1024 B apart (SBO).
          const uint64_t b_desc = B_MN ? make_sdesc_sw128(b_addr, 8192, 1024) : make_sdesc_sw128(b_addr, 16, 1024);
#pragma unroll
          for (int kk = 0; kk < kBlockK / 16; ++kk) {
            const uint64_t a_k = a_desc + static_cast<uint64_t>(A_MN ? kk * 128 : kk * 2);  // +2048 B | +32 B along K
            const uint64_t b_k = b_desc + static_cast<uint64_t>(B_MN ? kk * 128 : kk * 2);  // +2048 B | +32 B
            if constexpr (CG2) umma_bf16_cg2(d_tmem, a_k, b_k, p.idesc, (kb != wk.kb0 || kk != 0) ? 1u : 0u);
            else umma_bf16(d_tmem, a_k, b_k, p.idesc, (kb != wk.kb0 || kk != 0) ? 1u : 0u);
          }
          // frees the smem slot once these MMAs retire (in both CTAs of a multicast pair)
          if constexpr (CG2) umma_commit_cg2_mc(&empty_bar[stage], 3);
          else if constexpr (MC) umma_commit_mc(&empty_bar[stage], 3);
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (CG2) umma_commit_cg2_mc(&tfull_bar[acc], 3);  // both CTAs' epilogues read their own 128 rows
        else umma_commit(&tfull_bar[acc]);                           // accumulator complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // which half of the tile's columns this warp handles
    int acc = 0;
    uint32_t acc_phase = 0;
    const int n_out_total = (EPI == MM_EPI_SWIGLU) ? p.N / 2 : p.N;
    const int wi = warp - 2;  // 0..7: this warp's private flag / workspace lane of the stream-K hand-over
    GemmWork wk;
    for (int it = 0; gemm_work(p, worker, n_workers, total_tiles, it, wk); ++it) {
      TileCoord t = tile_coord(wk.tile, p, m_units);
      if constexpr (MC) t.m_blk = 2 * t.m_blk + cta_rank;
      const int row = t.m_blk * kBlockM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      if (wk.role == 1) {
        // ---- stream-K contributor: hand the raw fp32 partial accumulator of this K-range to the tile's finisher.
        // Slot layout [32-col chunk][4-col group 0..7][row 0..127][4 floats]: every warp store / load instruction moves
        // 512 contiguous bytes; warp (q, half) writes exactly the part the finisher's warp (q, half) reads, so the
        // hand-over is per warp.
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        constexpr int CPH_S = BN >= 64 ? BN / 64 : 1;
        float* slot = p.sk_ws + static_cast<long long>(worker) * (kBlockM * BN);
#pragma unroll 1
        for (int c = half * CPH_S; c < (half + 1) * CPH_S && c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(slot) + static_cast<long long>(c) * 8 * kBlockM + q * 32 + lane;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            dst[i * kBlockM] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                 __uint_as_float(r[4 * i + 3]));
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release_gpu(p.sk_flags + worker * 8 + wi, 1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
        continue;
      }
      // stream-K finisher: r[] += the contributors' partials of accumulator columns col_off .. col_off + 31 (fixed
      // order: deterministic); a no-op for ordinary tiles (warp-uniform branch)
      // (a CTA whose share of the tail is empty — more CTAs than tail k-blocks — has no piece and is skipped)
      const unsigned sk_units = static_cast<unsigned>(p.sk_tiles) * p.num_k;
      auto sk_has = [&](int sidx) {
        return static_cast<unsigned>(sidx) * sk_units / n_workers < static_cast<unsigned>(sidx + 1) * sk_units / n_workers;
      };
      auto sk_add = [&](uint32_t (&r)[32], int col_off) {
        if (wk.nc == 0) return;
        const int c = col_off >> 5;
        for (int sidx = wk.c0; sidx < wk.c0 + wk.nc; ++sidx) {
          if (!sk_has(sidx)) continue;
          const float4* src = reinterpret_cast<const float4*>(p.sk_ws + static_cast<long long>(sidx) * (kBlockM * BN)) +
                              static_cast<long long>(c) * 8 * kBlockM + q * 32 + lane;
          float4 fv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) fv[i] = __ldcg(src + i * kBlockM);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 f = fv[i];
            r[4 * i] = __float_as_uint(__uint_as_float(r[4 * i]) + f.x);
            r[4 * i + 1] = __float_as_uint(__uint_as_float(r[4 * i + 1]) + f.y);
            r[4 * i + 2] = __float_as_uint(__uint_as_float(r[4 * i + 2]) + f.z);
            r[4 * i + 3] = __float_as_uint(__uint_as_float(r[4 * i + 3]) + f.w);
          }
        }
      };
      if (wk.nc > 0) {  // wait for this warp's share of every contributor's partial
        if (lane == 0)
          for (int sidx = wk.c0; sidx < wk.c0 + wk.nc; ++sidx)
            if (sk_has(sidx)) {
              unsigned spins = 0;
              while (ld_acquire_gpu(p.sk_flags + sidx * 8 + wi) == 0) {
                __nanosleep(64);
                if (++spins > (1u << 25)) asm volatile("trap;");  // seconds: a scheduling bug must fail, never hang
              }
            }
        __syncwarp();
      }
      char* crow = reinterpret_cast<char*>(p.C) +
                   (static_cast<long long>(t.b_lo) * p.c_bs + static_cast<long long>(t.b_hi) * p.c_bs2 +
                    static_cast<long long>(row) * p.ldc) * (p.c_fp32 ? 4 : 2);
      float rs = 1.0f;
      if (p.row_scale != nullptr && row_ok && !p.c_trans) rs = p.row_scale[static_cast<long long>(t.b) * p.M + row];
      if (p.rs_sumsq != nullptr && row_ok) {
        // RMSNorm statistic of this A row from the partial sums the PRODUCING GEMM's epilogue left behind (fixed summation
        // order: deterministic); done while this tile's MMAs are still running
        const float4* sp = reinterpret_cast<const float4*>(p.rs_sumsq + static_cast<long long>(row) * p.rs_parts);
        float ssum = 0.f;
        for (int j = 0; j < p.rs_parts / 4; ++j) {
          const float4 f = __ldg(sp + j);
          ssum += (f.x + f.y) + (f.z + f.w);
        }
        rs = rsqrtf(ssum / static_cast<float>(p.K) + p.rs_eps);
      }
      rs *= p.alpha;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();

      if constexpr (EPI == MM_EPI_STD) {
        if (p.c_trans) {
          // "swap-AB" launches (few activation rows, many weight rows): the tile's rows are OUTPUT FEATURES (weights ride
          // the 128-row A operand so every MMA row is useful) and its columns are the activation rows.  C / residual are
          // addressed transposed, bias is per tile row, row_scale per tile column.  Outputs are tiny: scalar stores.
          constexpr int CPH_T = BN >= 64 ? BN / 64 : 1;
          const float bias_r = (p.bias != nullptr && row_ok) ? aux_ld(p.bias, row, p.aux_f16) : 0.f;
#pragma unroll 1
          for (int c = half * CPH_T; c < (half + 1) * CPH_T && c < BN / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(taddr + c * 32, r);
            tmem_ld_wait();
            sk_add(r, c * 32);
            const int col0 = t.n_blk * BN + c * 32;
            if (col0 >= p.N || !row_ok) continue;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int col = col0 + i;
              float cs = p.alpha;
              if (p.row_scale != nullptr && col < p.N) cs *= p.row_scale[col];
              v[i] = __uint_as_float(r[i]) * cs + bias_r;
            }
            if (p.act != MM_ACT_NONE) apply_act32(v, p.act);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int col = col0 + i;
              if (col < p.N) {
                float o = v[i];
                if (p.residual != nullptr) o += aux_ld(p.residual, static_cast<long long>(col) * p.ldr + row, p.aux_f16);
                if (p.c_fp32)
                  reinterpret_cast<float*>(p.C)[static_cast<long long>(col) * p.ldc + row] = o;
                else if (p.c_fp16)
                  reinterpret_cast<__half*>(p.C)[static_cast<long long>(col) * p.ldc + row] = __float2half_rn(o);
                else
                  reinterpret_cast<bf16*>(p.C)[static_cast<long long>(col) * p.ldc + row] = __float2bfloat16(o);
              }
            }
          }
        } else {
        const bf16* bias = p.bias ? p.bias + static_cast<long long>(t.b_lo) * p.bias_bs : nullptr;
        const bf16* rrow = nullptr;
        if (p.residual != nullptr && row_ok) {
          const int rr = p.res_row_mod > 0 ? row % p.res_row_mod : row;
          rrow = p.residual + static_cast<long long>(t.b_lo) * p.r_bs + static_cast<long long>(t.b_hi) * p.r_bs2 +
                 static_cast<long long>(rr) * p.ldr;
        }
        constexpr int CPH = BN >= 64 ? BN / 64 : 1;  // 32-column chunks per half
#pragma unroll 1
        for (int c = half * CPH; c < (half + 1) * CPH && c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          sk_add(r, c * 32);
          const int col0 = t.n_blk * BN + c * 32;
          if (col0 >= p.N) continue;  // warp-uniform
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * rs;
          const bool full = p.vec_ok && (col0 + 32 <= p.N);
          if (p.bias_rs != nullptr || p.bias2 != nullptr) {
            // row-scaled bias terms (value-side biases of the absorbed alignment attention); narrow GEMMs only
            const long long ri = static_cast<long long>(t.b) * p.M + row;
            const float s1 = (p.bias_rs != nullptr && row_ok) ? p.bias_rs[ri] : 1.0f;
            const float s2 = (p.bias2_rs != nullptr && row_ok) ? p.bias2_rs[ri] : 1.0f;
            const bf16* b2 = p.bias2 ? p.bias2 + static_cast<long long>(t.b_lo) * p.bias_bs : nullptr;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i < p.N) {
                if (bias != nullptr) v[i] = fmaf(s1, aux_ld(bias, col0 + i, p.aux_f16), v[i]);
                if (b2 != nullptr) v[i] = fmaf(s2, aux_ld(b2, col0 + i, p.aux_f16), v[i]);
              }
          } else if (bias != nullptr) {
            if (full) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                aux_add8(v + 8 * i, __ldg(reinterpret_cast<const uint4*>(bias + col0) + i), p.aux_f16);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) v[i] += aux_ld(bias, col0 + i, p.aux_f16);
            }
          }
          if (p.act != MM_ACT_NONE) apply_act32(v, p.act);
          if (rrow != nullptr) {
            if (full) {
#pragma unroll
              for (int i = 0; i < 4; ++i) aux_add8(v + 8 * i, *(reinterpret_cast<const uint4*>(rrow + col0) + i), p.aux_f16);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) v[i] += aux_ld(rrow, col0 + i, p.aux_f16);
            }
          }
          if (row_ok) store_row32(p, crow, col0, n_out_total, v);
          if (p.sumsq_out != nullptr && row_ok) {
            // sum of squares of the values AS STORED (rounded to the output format), for the next RMSNorm
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float r = p.c_fp16 ? __half2float(__float2half_rn(v[i])) : __bfloat162float(__float2bfloat16(v[i]));
              ss = fmaf(r, r, ss);
            }
            p.sumsq_out[static_cast<long long>(row) * p.sumsq_parts + (col0 >> 5)] = ss;
          }
        }
        }  // !c_trans
      } else if constexpr (EPI == MM_EPI_SWIGLU) {
#pragma unroll 1
        for (int c = half * (BN / 128); c < (half + 1) * (BN / 128); ++c) {
          uint32_t g[32], u[32];
          tmem_ld32(taddr + c * 64, g);
          tmem_ld32(taddr + c * 64 + 32, u);
          tmem_ld_wait();
          sk_add(g, c * 64);
          sk_add(u, c * 64 + 32);
          const int col_in = t.n_blk * BN + c * 64;
          if (col_in >= p.N) continue;
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float gg = __uint_as_float(g[i]) * rs;
            const float uu = __uint_as_float(u[i]) * rs;
            v[i] = gg * sigmoid_fast(gg) * uu;
          }
          if (row_ok) store_row32(p, crow, col_in / 2, n_out_total, v);
        }
      } else {  // MM_EPI_ROPE, head_dim 128: pairs (i, i + 64) within each head
        const int pos = (row_ok ? (row % p.rope_T) : 0) + (p.rope_pos != nullptr ? __ldg(p.rope_pos) : 0);
        const float* cs = p.rope_cos + static_cast<long long>(pos) * 64;
        const float* sn = p.rope_sin + static_cast<long long>(pos) * 64;
#pragma unroll 1
        for (int un = half * (BN / 128); un < (half + 1) * (BN / 128); ++un) {
          {
            const int h = un >> 1, hc = un & 1;
            uint32_t x1[32], x2[32];
            tmem_ld32(taddr + h * 128 + hc * 32, x1);
            tmem_ld32(taddr + h * 128 + 64 + hc * 32, x2);
            tmem_ld_wait();
            sk_add(x1, h * 128 + hc * 32);
            sk_add(x2, h * 128 + 64 + hc * 32);
            const int col1 = t.n_blk * BN + h * 128 + hc * 32;
            if (col1 >= p.N) continue;
            float o1[32], o2[32];
            if (col1 < p.rope_cols) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(cs + hc * 32) + i);
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(sn + hc * 32) + i);
                const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
                const float ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float a = __uint_as_float(x1[4 * i + j]) * rs;
                  const float b = __uint_as_float(x2[4 * i + j]) * rs;
                  o1[4 * i + j] = a * cc[j] - b * ss[j];
                  o2[4 * i + j] = b * cc[j] + a * ss[j];
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                o1[i] = __uint_as_float(x1[i]) * rs;
                o2[i] = __uint_as_float(x2[i]) * rs;
              }
            }
            if (row_ok) {
              store_row32(p, crow, col1, n_out_total, o1);
              store_row32(p, crow, col1 + 64, n_out_total, o2);
            }
          }
        }
      }
      if (wk.nc > 0) {  // partials consumed: re-arm this warp's flags for the next launch (stream order separates launches)
        __syncwarp();
        if (lane == 0)
          for (int sidx = wk.c0; sidx < wk.c0 + wk.nc; ++sidx)
            if (sk_has(sidx)) p.sk_flags[sidx * 8 + wi] = 0;
      }
      // release this accumulator buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));  // the leader's barrier
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  if constexpr (MC) cluster_sync_all(); else __syncthreads();  // a pair exits together: the peer may still signal us
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG2) tmem_dealloc2(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 4-D map {inner, rows, batch, batch2}; strides in elements; 128B swizzle; box {64, box_rows, 1, 1}.
int make_map(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint64_t batch, uint64_t batch2,
             int64_t ld, int64_t bs, int64_t bs2, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return 1;
  }
  cuuint64_t dims[4] = {inner, rows, batch, batch2};
  if (bs <= 0) bs = static_cast<int64_t>(rows) * ld;
  if (bs2 <= 0) bs2 = static_cast<int64_t>(batch) * bs;
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(bs) * 2,
                           static_cast<cuuint64_t>(bs2) * 2};
  cuuint32_t box[4] = {64, box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p dims={%llu,%llu,%llu,%llu} ld=%lld bs=%lld bs2=%lld box_rows=%u",
              static_cast<int>(r), ptr, (unsigned long long)inner, (unsigned long long)rows,
              (unsigned long long)batch, (unsigned long long)batch2, (long long)ld, (long long)bs, (long long)bs2,
              box_rows);
    return 1;
  }
  return 0;
}

// pair launches as ONE cta_group::2 MMA unit (default; 0 = multicast pairs of cta_group::1 MMAs); MACAW_B200_GEMM_CG2.
// Measured on B200, cfg4 LLaMA GEMMs at M = 16896: QKV 1153 -> 1086 us, o_proj 414 -> 383, gate-up 1950 -> 1877, down 1180 ->
// 1041 (1.46 - 1.64 PFLOP/s); in-step under the power cap: 212.8 -> 206.2 ms (fp16), 202.6 -> 196.8 ms (bf16).
static int& cg2_mode() {
  static int mode = []() { const char* e = getenv("MACAW_B200_GEMM_CG2"); return e ? (atoi(e) != 0 ? 1 : 0) : 1; }();
  return mode;
}

// switch for rule (b) above (MACAW_B200_GEMM_CG2_ODD, default on).  Measured at M = 2112 (tools/profile_gemms.py --batch 4):
// QKV 165.0 -> 154.8 us, down_proj 152.7 -> 144.5, o_proj 56.9 -> 56.7; bench.py --global-batch 4: 30.7 / 31.2 -> 29.6 ms.
static int& odd_cg2_mode() {
  static int mode = []() { const char* e = getenv("MACAW_B200_GEMM_CG2_ODD"); return e ? atoi(e) : 1; }();
  return mode;
}

// process-wide stream-K policy; initial value from MACAW_B200_GEMM_STREAMK (default 1)
static int& streamk_mode() {
  static int mode = []() { const char* e = getenv("MACAW_B200_GEMM_STREAMK"); const int v = e ? atoi(e) : 1; return v < 0 || v > 2 ? 1 : v; }();
  return mode;
}

template <int BN, int EPI, bool B_MN, bool MC, bool A_MN = false, bool CG2 = false>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmKParams& p, cudaStream_t st) {
  static bool attr_set[kMaxDevices] = {};
  constexpr size_t smem = gemm_smem_bytes(BN);
  auto kern = gemm_bf16_kernel<BN, EPI, B_MN, MC, A_MN, CG2>;
  if (int rc = ensure_smem_attr(kern, smem, attr_set, "mm_gemm_fwd")) return rc;
  const int m_units = MC ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const int total = p.batch * p.batch2 * m_units * p.n_tiles;
  cudaError_t e;
  if constexpr (MC) {
    const int pairs = total < num_sms() / 2 ? total : num_sms() / 2;
    e = launch_kernel(kern, dim3(2 * pairs), dim3(320), smem, st, 2, ta, tb, p);
  } else {
    const int grid = (total < num_sms() && p.sk_tiles == 0) ? total : num_sms();  // stream-K shares the tail over ALL SMs
    e = launch_kernel(kern, dim3(grid), dim3(320), smem, st, 1, ta, tb, p);
  }
  if (e != cudaSuccess) {
    set_error("mm_gemm_fwd: launch failed: %s", cudaGetErrorString(e));
    return 2;
  }
  return check_launch("mm_gemm_fwd");
}

}  // namespace mm

using namespace mm;

// The whole host side of a GEMM call: argument checks, tile width, pair / cta_group::2 / stream-K decisions, rasterisation
// group, tensor maps, launch.  With `plan` set it stops after the decisions (nothing is dereferenced, encoded or launched):
// mm_gemm_plan() exposes the schedule to the CPU test tier and to tools/gemm_plan.py.
static int32_t gemm_dispatch(const mm_gemm_args* a, void* stream, mm_gemm_schedule* plan) {
  MM_REQUIRE(a != nullptr, "mm_gemm_fwd: null args");
  MM_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0 && a->batch > 0 && a->batch2 >= 0,
             "mm_gemm_fwd: bad shape M=%d N=%d K=%d batch=%d batch2=%d", a->M, a->N, a->K, a->batch, a->batch2);
  const int batch2 = a->batch2 > 0 ? a->batch2 : 1;
  MM_REQUIRE(a->A && a->B && a->C, "mm_gemm_fwd: null operand");
  MM_REQUIRE((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0,
             "mm_gemm_fwd: A/B must be 16-byte aligned");
  MM_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0, "mm_gemm_fwd: lda/ldb must be multiples of 8 elements (lda=%lld ldb=%lld)",
             (long long)a->lda, (long long)a->ldb);
  MM_REQUIRE(a->batch == 1 || (a->a_bs % 8 == 0 && a->b_bs % 8 == 0), "mm_gemm_fwd: batch strides must be multiples of 8");
  MM_REQUIRE(batch2 == 1 || (a->a_bs2 % 8 == 0 && a->b_bs2 % 8 == 0), "mm_gemm_fwd: batch2 strides must be multiples of 8");
  MM_REQUIRE(a->epi >= MM_EPI_STD && a->epi <= MM_EPI_ROPE, "mm_gemm_fwd: bad epilogue %d", a->epi);

  GemmKParams p;
  p.M = a->M; p.N = a->N; p.K = a->K; p.batch = a->batch; p.batch2 = batch2;
  p.num_k = (a->K + kBlockK - 1) / kBlockK;
  p.m_tiles = (a->M + kBlockM - 1) / kBlockM;
  p.b_shared = (a->batch > 1 && a->b_bs == 0) ? 1 : 0;
  p.b2_shared = (batch2 > 1 && a->b_bs2 == 0) ? 1 : 0;
  p.C = a->C; p.ldc = a->ldc; p.c_bs = a->c_bs; p.c_bs2 = a->c_bs2; p.c_fp32 = a->c_fp32;
  p.act = a->act; p.alpha = a->alpha;
  p.bias = reinterpret_cast<const bf16*>(a->bias); p.bias_bs = a->bias_bs;
  p.row_scale = a->row_scale;
  p.residual = reinterpret_cast<const bf16*>(a->residual); p.ldr = a->ldr; p.r_bs = a->r_bs; p.r_bs2 = a->r_bs2;
  p.res_row_mod = a->res_row_mod;
  p.rope_cos = a->rope_cos; p.rope_sin = a->rope_sin; p.rope_T = a->rope_T; p.rope_cols = a->rope_cols;
  p.rope_pos = a->rope_pos;
  p.c_trans = a->c_trans;
  p.c_fp16 = a->c_fp16;
  p.aux_f16 = act_f16() ? 1 : 0;
  p.bias_rs = a->bias_rs; p.bias2 = reinterpret_cast<const bf16*>(a->bias2); p.bias2_rs = a->bias2_rs;
  p.sumsq_out = a->sumsq_out; p.sumsq_parts = (a->N + 31) / 32;
  p.rs_sumsq = a->rs_sumsq; p.rs_parts = a->rs_parts; p.rs_eps = a->rs_eps;
  MM_REQUIRE(a->sumsq_out == nullptr || (a->epi == MM_EPI_STD && !a->c_trans && !a->c_fp32 && a->batch == 1 && batch2 == 1 &&
                                         a->N % 32 == 0),
             "mm_gemm_fwd: sumsq_out needs the standard epilogue, 16-bit output, no batching and N %% 32 == 0");
  MM_REQUIRE(a->rs_sumsq == nullptr || (a->rs_parts > 0 && a->rs_parts % 4 == 0 && !a->c_trans && a->batch == 1 && batch2 == 1 &&
                                        (reinterpret_cast<uintptr_t>(a->rs_sumsq) & 15) == 0),
             "mm_gemm_fwd: rs_sumsq needs rs_parts %% 4 == 0, a 16-byte aligned buffer and no batching");
  MM_REQUIRE(!(a->c_fp16 && a->c_fp32), "mm_gemm_fwd: c_fp16 and c_fp32 are exclusive");
  MM_REQUIRE((a->a_fp16 != 0) == (a->b_fp16 != 0),
             "mm_gemm_fwd: A and B must share one 16-bit format (sm_100a faults on mixed f16 x bf16 tcgen05.mma)");
  MM_REQUIRE((!a->bias_rs && !a->bias2 && !a->bias2_rs) || (a->epi == MM_EPI_STD && !a->c_trans),
             "mm_gemm_fwd: row-scaled bias terms need the standard, non-transposed epilogue");
  MM_REQUIRE(!(a->c_fp16 && a->c_fp32), "mm_gemm_fwd: c_fp16 and c_fp32 are exclusive");
  MM_REQUIRE(!a->c_trans || (a->epi == MM_EPI_STD && a->batch == 1 && batch2 == 1),
             "mm_gemm_fwd: c_trans needs the standard epilogue and no batching");

  const int esz = a->c_fp32 ? 4 : 2;
  bool vec = (reinterpret_cast<uintptr_t>(a->C) % 16 == 0) && ((a->ldc * esz) % 16 == 0) &&
             ((a->c_bs * esz) % 16 == 0) && ((a->c_bs2 * esz) % 16 == 0);
  if (a->bias) vec = vec && (reinterpret_cast<uintptr_t>(a->bias) % 16 == 0) && (a->bias_bs % 8 == 0);
  if (a->residual)
    vec = vec && (reinterpret_cast<uintptr_t>(a->residual) % 16 == 0) && (a->ldr % 8 == 0) && (a->r_bs % 8 == 0) &&
          (a->r_bs2 % 8 == 0);
  p.vec_ok = vec ? 1 : 0;

  // ---- tile width: widest BN that still yields enough tiles to occupy the SMs
  const int sms = num_sms();
  int BN = 256;
  if (a->epi == MM_EPI_ROPE) {
    MM_REQUIRE(a->N % 128 == 0 && a->rope_cos && a->rope_sin && a->rope_T > 0 && vec && a->rope_cols % 128 == 0,
               "mm_gemm_fwd: RoPE epilogue needs N %% 128 == 0, cos/sin tables and vector-aligned C");
    const long long t256 = (long long)a->batch * batch2 * p.m_tiles * ((a->N + 255) / 256);
    BN = (a->N % 256 == 0 && t256 >= sms) ? 256 : 128;
  } else if (a->epi == MM_EPI_SWIGLU) {
    MM_REQUIRE(a->N % 64 == 0, "mm_gemm_fwd: SwiGLU epilogue needs N %% 64 == 0");
    const long long t256 = (long long)a->batch * batch2 * p.m_tiles * ((a->N + 255) / 256);
    BN = (t256 >= sms) ? 256 : 128;
  } else {
    // Cost model (cycles per 64-deep k-block of one tile, cta_group::1): the MMA needs 2*BN cycles, shared memory must
    // deliver 16 KiB of A + 128*BN bytes of B at 128 B/cycle -> 128 + BN cycles; BN = 128 sits exactly on the smem
    // limit (measured slower than the model, hence the 10 % penalty).  Waves = ceil(tiles / SMs).
    const int cands[4] = {256, 128, 64, 32};
    const long long kcost[4] = {512, 282, 192, 160};
    long long best = -1;
    BN = 32;
    for (int i = 0; i < 4; ++i) {
      const int bn = cands[i];
      if (a->b_mn_major && bn < 64) continue;
      if (bn > 32 && a->N <= bn / 2) continue;  // do not waste more than half a tile on padding
      const long long tiles = (long long)a->batch * batch2 * p.m_tiles * ((a->N + bn - 1) / bn);
      const long long waves = (tiles + sms - 1) / sms;
      // + a per-tile constant for the epilogue / pipeline fill (in k-block units of the same cost scale)
      const long long cost = waves * ((long long)p.num_k * kcost[i] + 2 * kcost[i]);
      if (best < 0 || cost < best) {
        best = cost;
        BN = bn;
      }
    }
    if (a->b_mn_major && BN < 64) BN = 64;
  }
  p.n_tiles = (a->N + BN - 1) / BN;
  p.idesc = make_idesc_f16(kBlockM, BN, a->a_mn_major != 0, a->b_mn_major != 0, a->a_fp16 != 0, a->b_fp16 != 0);
  static const int gm_env = []() { const char* e = getenv("MACAW_B200_GEMM_GROUPM"); return e ? atoi(e) : 0; }();
  // multicast pairs: wide tiles, several waves of work, and at most ~3 % of rows lost to an odd number of M tiles
  static const int mc_env = []() { const char* e = getenv("MACAW_B200_GEMM_MC"); return e ? atoi(e) : 1; }();
  const long long tiles256 = (long long)a->batch * batch2 * p.m_tiles * p.n_tiles;
  // (measured A/B on one box, cfg4: LLaMA GEMMs 1282 -> 1325 TFLOP/s; short-K CLIP GEMMs do not gain, hence K >= 2048)
  // Pairs with an odd number of M tiles leave the last pair with an idle half.  Two cases are worth it: (a) the idle rows are
  // <= 3 % of the tile rows; (b) the pair runs as a cta_group::2 unit (worth 6 - 12 % on these shapes) and the pair schedule
  // needs no more waves than the single-CTA schedule would — M = 2112 (17 M tiles, the per-GPU batch of the 8-GPU run):
  // QKV 432 pair-tiles on 74 pairs = 6 waves vs 816 tiles on 148 CTAs = 6 waves; gate-up 11 vs 10 waves stays single.
  const int pairs_m = (p.m_tiles + 1) / 2;
  const long long pair_tiles = (long long)a->batch * batch2 * pairs_m * p.n_tiles;
  const long long pair_waves = (pair_tiles + sms / 2 - 1) / (sms / 2), single_waves = (tiles256 + sms - 1) / sms;
  const bool odd_small = (pairs_m * 2 - p.m_tiles) * 32 <= p.m_tiles && tiles256 >= 2LL * sms;
  const bool odd_cg2 = cg2_mode() != 0 && odd_cg2_mode() != 0 && tiles256 > sms && pair_waves <= single_waves;
  static const int min_k = []() { const char* e = getenv("MACAW_B200_GEMM_PAIR_MIN_KBLOCKS"); return e ? atoi(e) : 32; }();
  const bool use_mc = mc_env != 0 && !a->a_mn_major && BN == 256 && p.m_tiles >= 2 && p.num_k >= min_k && (odd_small || odd_cg2);
  // rasterisation: keep one group's A rows (~32 MiB) resident in the 126 MB L2 while its B tiles stream
  // (measured on cfg4: 16 pairs at K=4096 is the optimum; 4 / 8 / 32 cost +5 % / +1 % / +9 % step time)
  {
    const long long unit_bytes = (long long)(use_mc ? 2 : 1) * kBlockM * a->K * 2;
    long long g = ((32LL << 20) + unit_bytes / 2) / unit_bytes;
    g = g < 2 ? 2 : (g > 32 ? 32 : g);
    p.group_m = gm_env > 0 ? gm_env : static_cast<int>(g);
  }
  // ---- stream-K tail: only when the last wave is clearly partial (<= 85 % full) and K is long enough to split
  p.sk_tiles = 0; p.sk_first = 0; p.sk_ws = nullptr; p.sk_flags = nullptr;
  const int sk_env = streamk_mode();  // 0 off, 1 when it pays (default), 2 whenever the schedule allows (tests)
  // (needs at least one full wave: the tail pieces run FIRST and their hand-over hides behind the full tiles' main loops;
  //  with fewer tiles than SMs — the thin GEMMs of a decode step — it would be exposed: measured 27 vs 22 us per GEMM
  //  against a fixed split-K of 4 + reduce)
  if (sk_env != 0 && a->sk_workspace != nullptr && !use_mc && p.num_k >= 8 && tiles256 > sms &&
      static_cast<long long>(sms) * p.num_k * (sms + 1) < (1LL << 31)) {
    const int rem = static_cast<int>(tiles256 % sms);
    const long long need = 8192 + static_cast<long long>(sms) * kBlockM * BN * 4;
    // Worth it only when the saved MMA time clearly exceeds the hand-over cost (partials through L2 compete with the
    // operand traffic: ~9 us measured at 128 x 256 tiles).  Saved time = (1 - rem / SMs) of one tile = (SMs - rem) / SMs *
    // num_k * 0.44 us at BN = 256; measured on B200, M = 2112: QKV (76 tiles left, K = 4096) 163.0 -> 159.6 us, down_proj
    // (124 left, K = 11008) 150.2 -> 146.6 us, o_proj (124 left, K = 4096) 56.9 -> 60.5 us — hence the threshold.
    const bool pays = sk_env == 2 || static_cast<long long>(sms - rem) * p.num_k * BN >= 3400LL * 256;
    if (rem > 0 && pays && a->sk_workspace_bytes >= need &&
        (reinterpret_cast<uintptr_t>(a->sk_workspace) & 15) == 0) {
      p.sk_tiles = rem;
      p.sk_first = static_cast<int>(tiles256 - rem);
      p.sk_flags = reinterpret_cast<int*>(a->sk_workspace);
      p.sk_ws = reinterpret_cast<float*>(reinterpret_cast<char*>(a->sk_workspace) + 8192);
    }
  }
  MM_REQUIRE(!a->a_mn_major || (a->b_mn_major && a->epi == MM_EPI_STD && !a->c_trans),
             "mm_gemm_fwd: MN-major A needs MN-major B and the standard epilogue");
  MM_REQUIRE(!a->b_mn_major || a->epi == MM_EPI_STD, "mm_gemm_fwd: MN-major B only with the standard epilogue");
  if (plan != nullptr) {
    const int m_units = use_mc ? (p.m_tiles + 1) / 2 : p.m_tiles;
    const long long units = (long long)a->batch * batch2 * m_units * p.n_tiles;
    const int workers = use_mc ? sms / 2 : sms;  // CTAs, or CTA pairs
    plan->block_n = BN;
    plan->pairs = use_mc ? (cg2_mode() != 0 ? 2 : 1) : 0;
    plan->m_tiles = p.m_tiles;
    plan->n_tiles = p.n_tiles;
    plan->k_blocks = p.num_k;
    plan->units = units;
    plan->workers = workers;
    plan->grid = use_mc ? 2 * static_cast<int>(units < workers ? units : workers)
                        : static_cast<int>((units < sms && p.sk_tiles == 0) ? units : sms);
    plan->waves = static_cast<int>((units + workers - 1) / workers);
    plan->group_m = p.group_m;
    plan->streamk_tiles = p.sk_tiles;
    plan->smem_bytes = static_cast<int32_t>(gemm_smem_bytes(BN));
    plan->vectorised_epilogue = p.vec_ok;
    return 0;
  }
  CUtensorMap ta, tb;
  if (a->a_mn_major) {
    if (make_map(&ta, a->A, a->M, a->K, a->batch, batch2, a->lda, a->a_bs, a->a_bs2, 64)) return 1;
  } else if (make_map(&ta, a->A, a->K, a->M, a->batch, batch2, a->lda, a->a_bs, a->a_bs2, kBlockM)) {
    return 1;
  }
  const uint64_t b_batch = p.b_shared ? 1 : a->batch;
  const uint64_t b_batch2 = p.b2_shared ? 1 : batch2;
  if (!a->b_mn_major) {
    if (make_map(&tb, a->B, a->K, a->N, b_batch, b_batch2, a->ldb, a->b_bs, a->b_bs2, use_mc ? BN / 2 : BN)) return 1;
  } else {
    if (make_map(&tb, a->B, a->N, a->K, b_batch, b_batch2, a->ldb, a->b_bs, a->b_bs2, 64)) return 1;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

#define MM_LAUNCH(BN_, EPI_, MN_) return launch_gemm<BN_, EPI_, MN_, false>(ta, tb, p, st)
#define MM_LAUNCH_MC(EPI_, MN_) return launch_gemm<256, EPI_, MN_, true>(ta, tb, p, st)
  if (use_mc && cg2_mode() != 0) {
    p.idesc = make_idesc_f16(2 * kBlockM, BN, false, a->b_mn_major != 0, a->a_fp16 != 0, a->b_fp16 != 0);
    if (a->b_mn_major) return launch_gemm<256, MM_EPI_STD, true, true, false, true>(ta, tb, p, st);
    if (a->epi == MM_EPI_ROPE) return launch_gemm<256, MM_EPI_ROPE, false, true, false, true>(ta, tb, p, st);
    if (a->epi == MM_EPI_SWIGLU) return launch_gemm<256, MM_EPI_SWIGLU, false, true, false, true>(ta, tb, p, st);
    return launch_gemm<256, MM_EPI_STD, false, true, false, true>(ta, tb, p, st);
  }
  if (use_mc) {
    if (a->epi == MM_EPI_ROPE) MM_LAUNCH_MC(MM_EPI_ROPE, false);
    if (a->epi == MM_EPI_SWIGLU) MM_LAUNCH_MC(MM_EPI_SWIGLU, false);
    if (a->b_mn_major) MM_LAUNCH_MC(MM_EPI_STD, true);
    MM_LAUNCH_MC(MM_EPI_STD, false);
  }
  if (a->epi == MM_EPI_ROPE) {
    if (BN == 256) MM_LAUNCH(256, MM_EPI_ROPE, false);
    MM_LAUNCH(128, MM_EPI_ROPE, false);
  }
  if (a->epi == MM_EPI_SWIGLU) {
    if (BN == 256) MM_LAUNCH(256, MM_EPI_SWIGLU, false);
    MM_LAUNCH(128, MM_EPI_SWIGLU, false);
  }
  if (a->a_mn_major) {
    if (BN == 256) return launch_gemm<256, MM_EPI_STD, true, false, true>(ta, tb, p, st);
    if (BN == 128) return launch_gemm<128, MM_EPI_STD, true, false, true>(ta, tb, p, st);
    return launch_gemm<64, MM_EPI_STD, true, false, true>(ta, tb, p, st);
  }
  if (a->b_mn_major) {
    if (BN == 256) MM_LAUNCH(256, MM_EPI_STD, true);
    if (BN == 128) MM_LAUNCH(128, MM_EPI_STD, true);
    MM_LAUNCH(64, MM_EPI_STD, true);
  }
  if (BN == 256) MM_LAUNCH(256, MM_EPI_STD, false);
  if (BN == 128) MM_LAUNCH(128, MM_EPI_STD, false);
  if (BN == 64) MM_LAUNCH(64, MM_EPI_STD, false);
  MM_LAUNCH(32, MM_EPI_STD, false);
#undef MM_LAUNCH
#undef MM_LAUNCH_MC
}

extern "C" int32_t mm_gemm_fwd(const mm_gemm_args* a, void* stream) { return gemm_dispatch(a, stream, nullptr); }

extern "C" int32_t mm_gemm_plan(const mm_gemm_args* a, mm_gemm_schedule* plan) {
  MM_REQUIRE(plan != nullptr, "mm_gemm_plan: null plan");
  return gemm_dispatch(a, nullptr, plan);
}

extern "C" int32_t mm_gemm_cg2_mode(int32_t mode) {
  const int prev = cg2_mode();
  if (mode == 0 || mode == 1) cg2_mode() = mode;
  return prev;
}

extern "C" int32_t mm_gemm_streamk_mode(int32_t mode) {
  const int prev = streamk_mode();
  if (mode >= 0 && mode <= 2) streamk_mode() = mode;
  return prev;
}

extern "C" int64_t mm_gemm_streamk_workspace_bytes(void) {
  return 8192 + static_cast<int64_t>(num_sms()) * kBlockM * 256 * 4;
}

// ------------------------------------------------------------------------------------------------ split-K reduce
namespace mm {
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int M, int N,
                                     const bf16* __restrict__ bias, bf16* __restrict__ out, long long ldo, int out_fp16,
                                     int bias_f16) {
  const long long total = static_cast<long long>(M) * N;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / N), n = static_cast<int>(i % N);
    float acc = bias ? aux_ld(bias, n, bias_f16) : 0.0f;
    for (int s = 0; s < splits; ++s) acc += part[static_cast<long long>(s) * total + i];
    if (out_fp16)
      reinterpret_cast<__half*>(out)[static_cast<long long>(m) * ldo + n] = __float2half_rn(acc);
    else
      out[static_cast<long long>(m) * ldo + n] = __float2bfloat16(acc);
  }
}
}  // namespace mm

extern "C" int32_t mm_splitk_reduce(const float* partial, int32_t splits, int32_t M, int32_t N, const void* bias,
                                    void* out, int64_t ldo, int32_t out_fp16, void* stream) {
  MM_REQUIRE(partial && out && splits > 0 && M > 0 && N > 0, "mm_splitk_reduce: bad arguments");
  const long long total = static_cast<long long>(M) * N;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  splitk_reduce_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      partial, splits, M, N, reinterpret_cast<const bf16*>(bias), reinterpret_cast<bf16*>(out), ldo, out_fp16,
      act_f16() ? 1 : 0);
  return check_launch("mm_splitk_reduce");
}
