// Flash attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA) for head_dim 64 / 96 / 128.
//
// One CTA = 128 queries of one (batch, head); key tiles of KT = 64 keys so that TWO CTAs fit per SM (smem <= 113 KB,
// 256 TMEM columns each): one CTA's prologue / epilogue / softmax latency hides behind the other's MMAs.  Warp roles:
//   warp 0      : TMA producer — Q once, then K / V tiles of KT keys into 2-stage rings (128B swizzle)
//   warp 1      : single-thread tcgen05.mma issuer:  S_j = Q K_j^T (TMEM, double-buffered)  and  O += P_j V_j (TMEM)
//   warps 2..5  : softmax — ONE THREAD PER QUERY ROW (tcgen05.ld 32x32b gives each thread its row, so the row max /
//                 row sum need no shuffles); writes P_j as bf16 into a 128B-swizzled smem tile that is the A operand
//                 of the PV MMA; V is consumed as an MN-major B operand straight from its row-major layout.
// S_{j+2} and PV_j are issued while the softmax warps work on S_{j+1}, so tensor pipe and MUFU overlap.
// O stays in TMEM across key tiles; it is rescaled (tcgen05.ld -> scale -> tcgen05.st) only when a row's running max
// grows by more than 2^8 since the max the accumulator is expressed in ("lazy rescale"); P and the row sums use that
// same reference max, so the final O / l is exact.
//
// Masking: causal (key j visible to query i iff j <= i + Tk - Tq), key padding mask, and the Tk bound, evaluated as
// 32-bit masks per 32-key chunk.  Rows with every key masked produce zeros (see DESIGN.md "unspecified rows").
//
// Reference call sites replaced: see include/macaw_b200.h (mm_attn_fwd).  head_dim 96 (video-long self-attention,
// reference modeling.py:1078) runs on the HD = 128 instantiation: the Q / K / V tensor maps carry the real head dim, so
// the TMA boxes of the second 64-column block are zero-filled past column 96; S = Q K^T issues 6 of 8 k-steps and
// O += P V uses an N = 96 instruction.  The mma.sync kernel in attn.cu is kept only as a second implementation for tests.
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/macaw_b200.h"

namespace mm {

struct FaParams {
  bf16* out;
  int B, H, Tq, Tk;
  long long o_bs, o_ts, o_hs;
  const int* key_mask;
  const int* tk_dev;  // optional device-side number of valid keys (<= Tk)
  int causal;
  float scale_log2;
  int hd;   // actual head dim (<= HD, multiple of 32): 96 runs on the HD = 128 instantiation — TMA zero-fills the
            // out-of-range columns of the second 64-column block, S skips the all-zero k-steps, PV uses N = hd
  int qpc;  // query tiles per CTA (walked heaviest first)
  int rev;  // launch query-tile groups in descending order (heavier causal groups first)
};

constexpr int kFaMQ = 128;  // queries per CTA
constexpr int kFaMaskWords = 64;  // key-validity words kept in shared memory (keys < 2048; longer rows load per tile)
constexpr float kFaTau = 8.0f;  // lazy-rescale threshold (log2 domain)

__device__ __forceinline__ float max3(float a, float b, float c) {  // one FMNMX3 on sm_100
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint64_t pack_b32x2(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) { return pack_b32x2(__float_as_uint(lo), __float_as_uint(hi)); }
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// HD head dim (64 | 128), KT keys per tile (64 | 128), NPB number of P buffers (1 | 2)
template <int HD, int KT, int NPB>
__host__ __device__ constexpr size_t fa_smem_bytes() {
  // Q + 2 K stages + 2 V stages + NPB P buffers + barriers (the dynamic smem window itself is 1024-byte aligned)
  // ... + 21 barriers + TMEM slot (256 B) + key-validity bits of up to kFaMaskWords * 32 keys (256 B)
  return (size_t)(HD / 64) * 16384 + 4 * (size_t)(HD / 64) * KT * 128 + (size_t)NPB * 128 * KT * 2 + 256 + 256;
}

template <int HD, int KT, int NPB, bool F16>
__global__ void __launch_bounds__(192, 2)
fa_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const FaParams p) {
  constexpr int kFaKT = KT;
  constexpr int KB = HD / 64;                 // 64-column blocks of the head dim
  constexpr uint32_t QBYTES = KB * 16384;     // bytes of the Q tile
  constexpr uint32_t QB = KB * KT * 128;      // bytes of one K / V stage
  constexpr uint32_t PB = 128 * KT * 2;       // bytes of one P buffer
  constexpr int NCH = KT / 32;                // 32-key chunks per tile
  constexpr uint32_t IDESC_S = make_idesc_f16(kFaMQ, KT, false, false, F16, F16);
  const uint32_t IDESC_O = make_idesc_f16(kFaMQ, p.hd, false, true, F16, F16);
  constexpr uint32_t S_COL = 0, O_COL = 2 * KT;  // TMEM columns: S0 [0,KT), S1 [KT,2KT), O [2KT, 2KT+HD)
  constexpr uint32_t TMEM_COLS = (2 * KT + HD) <= 256 ? 256 : 512;
  static_assert(2 * KT + HD <= 512, "TMEM budget");

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B-swizzled TMA / UMMA tiles need 1024-byte alignment
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + QBYTES;
  uint8_t* sV = sK + 2 * QB;
  uint8_t* sP = sV + 2 * QB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + NPB * PB);
  uint64_t* bar_q = bars;            // 1  Q tile landed
  uint64_t* bar_k = bars + 1;        // 2  K stage landed
  uint64_t* bar_v = bars + 3;        // 2  V stage landed
  uint64_t* bar_kfree = bars + 5;    // 2  S MMA finished reading the K stage
  uint64_t* bar_vfree = bars + 7;    // 2  PV MMA finished reading the V stage
  uint64_t* bar_s = bars + 9;        // 2  S tile complete in TMEM
  uint64_t* bar_sfree = bars + 11;   // 2  softmax has S in registers
  uint64_t* bar_p = bars + 13;       // 2  P tile written to smem
  uint64_t* bar_pfree = bars + 15;   // 2  PV MMA finished reading the P buffer
  uint64_t* bar_o = bars + 17;       // 1  all PV of the current query tile complete
  uint64_t* bar_qfree = bars + 18;   // 1  all S MMAs of the current query tile complete (Q may be replaced)
  uint64_t* bar_ofree = bars + 19;   // 1  softmax warps have read O (next query tile may overwrite it)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  uint32_t* s_kbits = reinterpret_cast<uint32_t*>(bars + 32);  // [kFaMaskWords]: bit i of word c = key 32 c + i is attendable

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  int Tk = p.Tk;          // replaced below by the device-side length when one is given
  int shift = Tk - p.Tq;  // causal: key j visible to query i iff j <= i + shift
  // This CTA walks `p.qpc` consecutive query tiles of (b, h), heaviest first; heavy groups are launched first.
  const int nq = (p.Tq + kFaMQ - 1) / kFaMQ;
  // Causal: the ragged query tile sits at the START of the sequence (tiles are aligned to the END), where it sees a single
  // key tile, instead of at the end where a mostly empty 128-row tile would walk every key tile (T = 528: 25 instead of
  // 29 key-tile iterations per head).  Rows before 0 are TMA zero fill and are neither attended nor stored.
  const int q_base = p.causal ? p.Tq - nq * kFaMQ : 0;
  const int grp = p.rev ? static_cast<int>(gridDim.x - 1 - blockIdx.x) : static_cast<int>(blockIdx.x);
  const int q_lo = grp * p.qpc, q_hi = min(nq, q_lo + p.qpc);  // query tiles [q_lo, q_hi)
  auto tiles_of = [&](int mt) {
    int kv_end = Tk;
    if (p.causal) kv_end = min(Tk, q_base + mt * kFaMQ + kFaMQ + shift);
    return kv_end > 0 ? (kv_end + kFaKT - 1) / kFaKT : 0;
  };

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    if (elect_one()) {
      mbar_init(bar_q, 1);
      mbar_init(bar_o, 1);
      mbar_init(bar_qfree, 1);
      mbar_init(bar_ofree, 4);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bar_k[i], 1);
        mbar_init(&bar_v[i], 1);
        mbar_init(&bar_kfree[i], 1);
        mbar_init(&bar_vfree[i], 1);
        mbar_init(&bar_s[i], 1);
        mbar_init(&bar_sfree[i], 4);
        mbar_init(&bar_p[i], 4);
        mbar_init(&bar_pfree[i], 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();  // prologue done: the next kernel may begin its own
  griddep_wait();    // q / k / v (and the device-side length) come from earlier kernels
  if (p.tk_dev != nullptr) {
    Tk = min(*p.tk_dev, p.Tk);
    shift = Tk - p.Tq;
  }

  // Every role walks the same sequence: query tiles mt = q_hi-1 .. q_lo, key tiles j = 0 .. n-1 of each.
  // G = running key-tile index across query tiles (ring stages / mbarrier parities), qa = running count of
  // non-empty query tiles.
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int G = 0, qa = 0;
      for (int mt = q_hi - 1; mt >= q_lo; --mt) {
        const int n_tiles = tiles_of(mt);
        if (n_tiles == 0) continue;
        if (qa >= 1) mbar_wait(bar_qfree, (qa - 1) & 1);  // the previous query tile's S MMAs are done with sQ
        mbar_arrive_expect_tx(bar_q, QBYTES);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) tma_load_4d(&tmQ, bar_q, sQ + kb * 16384, kb * 64, q_base + mt * kFaMQ, h, b);
        for (int j = 0; j < n_tiles; ++j, ++G) {
          const int st = G & 1;
          const uint32_t par = ((G >> 1) - 1) & 1;  // parity of the previous use of this stage
          if (G >= 2) mbar_wait(&bar_kfree[st], par);
          mbar_arrive_expect_tx(&bar_k[st], QB);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
            tma_load_4d(&tmK, &bar_k[st], sK + st * QB + kb * (KT * 128), kb * 64, j * kFaKT, h, b);
          if (G >= 2) mbar_wait(&bar_vfree[st], par);
          mbar_arrive_expect_tx(&bar_v[st], QB);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
            tma_load_4d(&tmV, &bar_v[st], sV + st * QB + kb * (KT * 128), kb * 64, j * kFaKT, h, b);
        }
        ++qa;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int G0 = 0, qa = 0;
      for (int mt = q_hi - 1; mt >= q_lo; --mt) {
        const int n_tiles = tiles_of(mt);
        if (n_tiles == 0) continue;
        auto issue_s = [&](int jj) {
          const int Gj = G0 + jj;
          const int st = Gj & 1;
          mbar_wait(&bar_k[st], (Gj >> 1) & 1);
          if (Gj >= 2) mbar_wait(&bar_sfree[st], ((Gj >> 1) - 1) & 1);
          tc_fence_after();
          const uint32_t qa_ = smem_u32(sQ), ka = smem_u32(sK + st * QB);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            if (k * 16 >= p.hd) break;  // columns >= hd are TMA zero fill
            const uint64_t ad = make_sdesc_sw128(qa_ + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
            const uint64_t bd = make_sdesc_sw128(ka + (k >> 2) * (KT * 128) + (k & 3) * 32, 16, 1024);
            umma_bf16(tmem_base + S_COL + st * kFaKT, ad, bd, IDESC_S, k != 0 ? 1u : 0u);
          }
          umma_commit(&bar_s[st]);
          umma_commit(&bar_kfree[st]);
          if (jj == n_tiles - 1) umma_commit(bar_qfree);  // last S of this query tile: sQ may be refilled
        };
        mbar_wait(bar_q, qa & 1);
        issue_s(0);
        if (n_tiles > 1) issue_s(1);
        if (qa >= 1) mbar_wait(bar_ofree, (qa - 1) & 1);  // the previous query tile's O has been read out of TMEM
        for (int j = 0; j < n_tiles; ++j) {
          const int Gj = G0 + j;
          const int st = Gj & 1;
          const int pb = (NPB == 2) ? st : 0;
          if (NPB == 2) mbar_wait(&bar_p[st], (Gj >> 1) & 1); else mbar_wait(&bar_p[0], Gj & 1);
          mbar_wait(&bar_v[st], (Gj >> 1) & 1);
          tc_fence_after();
          const uint32_t pa = smem_u32(sP + pb * PB), va = smem_u32(sV + st * QB);
#pragma unroll
          for (int k = 0; k < kFaKT / 16; ++k) {
            const uint64_t ad = make_sdesc_sw128(pa + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
            // V: MN-major B (head dim contiguous); 64-column blocks KT*128 B apart (LBO), 8-key groups 1 KiB apart (SBO)
            const uint64_t bd = make_sdesc_sw128(va + k * 2048, KT * 128, 1024);
            umma_bf16(tmem_base + O_COL, ad, bd, IDESC_O, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&bar_pfree[pb]);
          umma_commit(&bar_vfree[st]);
          if (j + 2 < n_tiles) issue_s(j + 2);
        }
        umma_commit(bar_o);
        G0 += n_tiles;
        ++qa;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps: one thread per query row
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row within the tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int* kmask = p.key_mask ? p.key_mask + static_cast<long long>(b) * p.Tk : nullptr;
    // Key-validity bits (Tk bound & key-padding mask) of the whole key range, built ONCE per CTA: loading the mask words
    // per key tile put two dependent global loads (~700 cycles of long-scoreboard stall, 13 % of all samples in
    // profiles/r2_ncu_attention_cfg4.txt) at the head of every tile's softmax.
    const bool bits_in_smem = Tk <= kFaMaskWords * 32;
    if (bits_in_smem) {
      const int n_words = (Tk + 31) >> 5;
      uint32_t mine[kFaMaskWords / 4];
#pragma unroll
      for (int k = 0; k < kFaMaskWords / 4; ++k) {  // all loads in flight, then the ballots
        const int key = (q + 4 * k) * 32 + lane;
        mine[k] = (q + 4 * k < n_words && key < Tk && (kmask == nullptr || kmask[key] != 0)) ? 1u : 0u;
      }
#pragma unroll
      for (int k = 0; k < kFaMaskWords / 4; ++k) {
        const uint32_t bits = __ballot_sync(0xffffffffu, mine[k] != 0u);
        if (lane == 0 && q + 4 * k < n_words) s_kbits[q + 4 * k] = bits;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four softmax warps only
    }
    const int sw = r & 7;
    bf16* og = p.out + static_cast<long long>(b) * p.o_bs + static_cast<long long>(h) * p.o_hs;
    int G0 = 0, qa = 0;

    for (int mt = q_hi - 1; mt >= q_lo; --mt) {
      const int n_tiles = tiles_of(mt);
      const int qrow = q_base + mt * kFaMQ + r;
      float m_used = -INFINITY;  // reference max the accumulator / row sum are expressed in (log2 domain)
      float l = 0.f;

      for (int j = 0; j < n_tiles; ++j) {
        const int Gj = G0 + j;
        const int st = Gj & 1;
        mbar_wait(&bar_s[st], (Gj >> 1) & 1);
        tc_fence_after();
        const uint32_t s_addr = tmem_base + lane_base + S_COL + st * kFaKT;
        const int key0 = j * kFaKT;
        // per-32-key validity bits: Tk bound + key padding mask (warp-cooperative) & causal limit (per row)
        uint32_t okb[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          uint32_t bits;
          if (bits_in_smem) {
            bits = (key0 + c * 32 < Tk) ? s_kbits[(key0 >> 5) + c] : 0u;
          } else {
            const int key = key0 + c * 32 + lane;
            bool ok = key < Tk;
            if (ok && kmask != nullptr) ok = kmask[key] != 0;
            bits = __ballot_sync(0xffffffffu, ok);
          }
          if (p.causal) {
            const int lim = qrow + shift - (key0 + c * 32);  // keys with index <= lim inside this chunk are visible
            bits &= lim >= 31 ? 0xffffffffu : (lim < 0 ? 0u : ((2u << lim) - 1u));
          }
          okb[c] = bits;
        }
        // ---- the whole score row of this tile into registers (all loads in flight, one wait)
        uint32_t v[NCH][32];
#pragma unroll
        for (int c = 0; c < NCH; ++c) tmem_ld32(s_addr + c * 32, v[c]);
        tmem_ld_wait();
        // S_j is in registers: the MMA warp may overwrite this S buffer with S_{j+2}
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_sfree[st]);
        // tiles without any masked element (the common case away from the causal diagonal / sequence end) take a
        // select-free path: the softmax warps are ALU/MUFU-bound, every instruction per element counts
        uint32_t allb = 0xffffffffu;
#pragma unroll
        for (int c = 0; c < NCH; ++c) allb &= okb[c];
        const bool full_tile = __all_sync(0xffffffffu, allb == 0xffffffffu);
        float mx = -INFINITY;
        if (full_tile) {
#pragma unroll
          for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int i = 0; i < 32; i += 2) mx = max3(mx, __uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1]));
        } else {
#pragma unroll
          for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (okb[c] & (1u << i)) mx = fmaxf(mx, __uint_as_float(v[c][i]));
        }
        mx *= p.scale_log2;  // scale > 0, so max commutes with the scaling
        // ---- decide the reference max for this tile
        float alpha = 1.0f;
        bool rescale = false;
        if (mx > m_used + kFaTau || m_used == -INFINITY) {
          if (mx != -INFINITY) {
            if (m_used != -INFINITY) {
              alpha = ex2_approx(m_used - mx);
              rescale = true;
            }
            m_used = mx;
          }
        }
        const float m_ref = (m_used == -INFINITY) ? 0.f : m_used;
        // ---- P = exp2(s * scale - m_ref) -> bf16 -> swizzled smem (A operand of the PV MMA)
        const int pb = (NPB == 2) ? st : 0;
        if (NPB == 2) {
          if (Gj >= 2) mbar_wait(&bar_pfree[st], ((Gj >> 1) - 1) & 1);  // PV_{G-2} has finished reading this P buffer
        } else {
          if (Gj >= 1) mbar_wait(&bar_pfree[0], (Gj - 1) & 1);          // PV_{G-1} has finished reading the P buffer
        }
        uint8_t* prow = sP + pb * PB + r * 128;
        float psum = 0.f;
        if (full_tile) {
          // packed fp32x2 arithmetic (FFMA2 / FADD2, sm_100): the softmax warps are issue-bound, and the scale-and-shift
          // and the row-sum accumulation are the two fp32 ops per element besides the ex2 — half the instructions each
          const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2), nm2 = pack_f32x2(-m_ref, -m_ref);
          uint64_t ps2 = pack_f32x2(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              uint64_t t2;
              asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t2) : "l"(pack_b32x2(v[c][2 * i], v[c][2 * i + 1])), "l"(sc2), "l"(nm2));
              float t0, t1;
              unpack_f32x2(t2, t0, t1);
              const float e0 = ex2_approx(t0), e1 = ex2_approx(t1);
              asm("add.rn.f32x2 %0, %1, %2;" : "=l"(ps2) : "l"(ps2), "l"(pack_f32x2(e0, e1)));
              pk[i] = pack2<F16>(e0, e1);
            }
            uint8_t* blk = prow + (c >> 1) * 16384;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<uint4*>(blk + ((((c & 1) * 4 + i) ^ sw) << 4)) =
                  make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          }
          float s0, s1;
          unpack_f32x2(ps2, s0, s1);
          psum = s0 + s1;
        } else {
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float e0 = ex2_approx(fmaf(__uint_as_float(v[c][2 * i]), p.scale_log2, -m_ref));
              float e1 = ex2_approx(fmaf(__uint_as_float(v[c][2 * i + 1]), p.scale_log2, -m_ref));
              e0 = (okb[c] & (1u << (2 * i))) ? e0 : 0.f;
              e1 = (okb[c] & (1u << (2 * i + 1))) ? e1 : 0.f;
              psum += e0 + e1;
              pk[i] = pack2<F16>(e0, e1);
            }
            uint8_t* blk = prow + (c >> 1) * 16384;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<uint4*>(blk + ((((c & 1) * 4 + i) ^ sw) << 4)) =
                  make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          }
        }
        l = l * alpha + psum;
        // ---- lazy rescale of the TMEM accumulator (rare): needs PV_{j-1} complete, must finish before PV_j starts
        if (__any_sync(0xffffffffu, rescale)) {
          if (NPB == 2) {
            if (Gj >= 1) mbar_wait(&bar_pfree[(Gj - 1) & 1], ((Gj - 1) >> 1) & 1);
          } else {
            if (Gj >= 1) mbar_wait(&bar_pfree[0], (Gj - 1) & 1);
          }
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < p.hd / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(tmem_base + lane_base + O_COL + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tmem_base + lane_base + O_COL + c * 32, o);
          }
          tmem_st_wait();
          tc_fence_before();
        }
        fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_p[pb]);
      }

      // ---- epilogue of this query tile: O / l -> bf16 -> global (each thread owns one row: HD*2 contiguous bytes)
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      if (n_tiles > 0) {
        mbar_wait(bar_o, qa & 1);
        tc_fence_after();
      }
      bf16* orow = og + static_cast<long long>(qrow) * p.o_ts;
#pragma unroll 1
      for (int c = 0; c < p.hd / 32; ++c) {
        uint32_t o[32];
        if (n_tiles > 0) {
          tmem_ld32(tmem_base + lane_base + O_COL + c * 32, o);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = 0u;
        }
        if (qrow >= 0 && qrow < p.Tq) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = pack2<F16>(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
            u.y = pack2<F16>(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
            u.z = pack2<F16>(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
            u.w = pack2<F16>(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c * 32 + i * 8) = u;
          }
        }
      }
      if (n_tiles > 0) {
        // O has left TMEM: the MMA warp may start the next query tile's PV (accumulate = 0 overwrites it)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ofree);
        G0 += n_tiles;
        ++qa;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int fa_make_map(CUtensorMap* m, const void* ptr, int hd, int T, int H, int B, int64_t ts, int64_t hs, int64_t bs,
                       uint32_t box_rows) {
  static EncodeTiledFn2 fn = nullptr;
  if (fn == nullptr) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn2>(f);
  }
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return 1;
  }
  // a size-1 dimension may carry any stride; keep every stride a positive multiple of 16 bytes
  if (hs <= 0) hs = static_cast<int64_t>(hd);
  if (bs <= 0) bs = static_cast<int64_t>(T) * ts;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(hd), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(ts) * 2, static_cast<cuuint64_t>(hs) * 2,
                           static_cast<cuuint64_t>(bs) * 2};
  cuuint32_t box[4] = {64, box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("mm_attn_fwd: cuTensorMapEncodeTiled failed (%d) hd=%d T=%d H=%d B=%d ts=%lld hs=%lld bs=%lld",
              static_cast<int>(r), hd, T, H, B, (long long)ts, (long long)hs, (long long)bs);
    return 1;
  }
  return 0;
}

template <int HD, int KT, int NPB, bool F16>
static int launch_fa(const mm_attn_args* a, cudaStream_t st) {
  static bool attr_set[kMaxDevices] = {};
  constexpr size_t smem = fa_smem_bytes<HD, KT, NPB>();
  if (int rc = ensure_smem_attr(fa_tcgen05_kernel<HD, KT, NPB, F16>, smem, attr_set, "mm_attn_fwd")) return rc;
  CUtensorMap tq, tk, tv;
  const int hd = a->head_dim;  // the maps carry the ACTUAL head dim: boxes reaching past it are zero-filled
  if (fa_make_map(&tq, a->q, hd, a->Tq, a->H, a->B, a->q_ts, a->q_hs, a->q_bs, kFaMQ)) return 1;
  if (fa_make_map(&tk, a->k, hd, a->Tk, a->H, a->B, a->k_ts, a->k_hs, a->k_bs, KT)) return 1;
  if (fa_make_map(&tv, a->v, hd, a->Tk, a->H, a->B, a->v_ts, a->v_hs, a->v_bs, KT)) return 1;
  FaParams p;
  p.hd = hd;
  p.out = reinterpret_cast<bf16*>(a->out);
  p.B = a->B; p.H = a->H; p.Tq = a->Tq; p.Tk = a->Tk;
  p.o_bs = a->o_bs; p.o_ts = a->o_ts; p.o_hs = a->o_hs;
  p.key_mask = a->key_mask;
  p.tk_dev = a->tk_dev;
  p.causal = a->causal;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  // query tiles per CTA: amortise the per-CTA prologue (TMEM alloc, barrier init, first TMA round trip) while keeping
  // at least ~4 waves of CTAs (2 CTAs per SM) for balance
  const int nq = (a->Tq + kFaMQ - 1) / kFaMQ;
  const int sms = num_sms();
  long long qpc = ((long long)nq * a->H * a->B) / (4LL * 2 * sms);
  if (qpc < 1) qpc = 1;
  if (qpc > nq) qpc = nq;
  if (qpc > 8) qpc = 8;
  p.qpc = static_cast<int>(qpc);
  p.rev = (a->causal && nq % p.qpc == 0) ? 1 : 0;  // a trailing partial group is the light one: launch it last
  dim3 grid((nq + p.qpc - 1) / p.qpc, a->H, a->B);
  cudaError_t e = launch_kernel(fa_tcgen05_kernel<HD, KT, NPB, F16>, grid, dim3(192), smem, st, 1, tq, tk, tv, p);
  if (e != cudaSuccess) {
    set_error("mm_attn_fwd: launch failed: %s", cudaGetErrorString(e));
    return 2;
  }
  return check_launch("mm_attn_fwd(tcgen05)");
}

// called from mm_attn_fwd (attn.cu) for head_dim 64 / 96 / 128 when scale > 0 (96 rides the 128 instantiation)
int attn_tcgen05_dispatch(const mm_attn_args* a, cudaStream_t st) {
  if (act_f16()) return a->head_dim == 64 ? launch_fa<64, 64, 2, true>(a, st) : launch_fa<128, 64, 1, true>(a, st);
  return a->head_dim == 64 ? launch_fa<64, 64, 2, false>(a, st) : launch_fa<128, 64, 1, false>(a, st);
}

}  // namespace mm
