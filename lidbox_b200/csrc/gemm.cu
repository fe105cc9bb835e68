// TMA-fed tcgen05 GEMM for the x-vector TDNN (sm_100a).
//
// One persistent, warp-specialised kernel serves every dense contraction of lidbox/models/xvector.py:
//   layout NT :  C[M,N] = A[M,K] . B[N,K]^T      forward Conv1D / Dense (xvector.py:38-43,53-64) and data gradients
//   layout TN :  C[M,N] = A[K,M]^T . B[K,N]      weight gradients (contraction over the batch*time rows), split-K
// A causal strided Conv1D is an NT GEMM without im2col: activations are NWC with k-1 zero rows in front of every
// utterance, so row (b,t) of the implicit matrix is the k*C_in contiguous elements starting at padded time t*stride;
// the A operand is a plain 2-D TMA view whose row pitch (stride*C_in) is smaller than its width (k*C_in).
//
// Roles (384 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer, warps 2-3 idle,
// warps 4..11 = epilogue (TMEM -> registers -> bias/ReLU/mask -> global).  bf16 results take the lean epilogue
// (prefetched tcgen05.ld, mask tiles by TMA bulk loads, one TMA bulk store per 32-column chunk); fp32 / atomic /
// read-modify-write / residual-plane results take the general one.  Operands are bf16 in 128B-swizzled
// shared memory (4 stages x 48 KB), accumulators fp32 in TMEM, double-buffered (2 x 256 columns) so the epilogue
// of tile i overlaps the main loop of tile i+1.  "bf16x3" mode runs three accumulating passes
// (A_hi B_hi + A_hi B_lo + A_lo B_hi) for fp32-grade results from bf16 tensor cores.
#include "tc_ptx.cuh"
#include <mutex>

namespace lbx {

#ifndef LBX_BIAS_SMEM
#define LBX_BIAS_SMEM 1
#endif
#ifndef LBX_CTRL_WARPS
#define LBX_CTRL_WARPS 4
#endif
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
constexpr int EPI_WARPS = 8;                  // two warps per TMEM sub-partition, each takes half of the tile's columns
constexpr int EPI_WARPS_WIDE = 16;            // "wide epilogue" variant: four warps per sub-partition, a quarter each
constexpr int GEMM_THREADS = 32 * LBX_CTRL_WARPS + 32 * EPI_WARPS;   // warpgroup 0: TMA / MMA / 2 idle warps; warpgroups 1-2: epilogue
constexpr int GEMM_THREADS_WIDE = 32 * LBX_CTRL_WARPS + 32 * EPI_WARPS_WIDE;
constexpr int BIAS_SMEM_FLOATS = 3072;         // the bias vector is staged in shared memory when N fits
constexpr int STAGE_PITCH = 80;                // bytes per staged row: 32 bf16 + 16 B pad (conflict-free 16-byte accesses)
constexpr int STAGE_BYTES_PER_WARP = 32 * STAGE_PITCH;
// fast bf16 epilogue: per warp one 32 x 64 B tile in the TMA 64-byte swizzle as the source of bulk tensor stores, and
// (ReLU-backward launches) one more as the destination of bulk tensor loads of the mask
constexpr int FAST_TILE_BYTES = 32 * 64;
constexpr int BAR_BYTES = 1024;                // mbarriers + TMEM pointer (keeps the epilogue tiles 1024-byte aligned)

// tile-N variants: 256 (large problems), 128, and 64 (the small-M dense layers: enough CTAs without split-K)
// PAIR: two CTAs of a cluster share one 256 x BN tile through tcgen05 cta_group::2 — each CTA stages its own 128 rows of
// A and only HALF of the B tile, which cuts the operand bytes every SM has to ingest per k-block from 48 KB to 32 KB
template <int BN, bool PAIR = false, bool HAS_MASK = false, bool WIDE = false>
struct Cfg {
  static constexpr int EW = WIDE ? EPI_WARPS_WIDE : EPI_WARPS;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // the mask tiles of the ReLU-backward launches take the shared memory of one pipeline stage; so do the extra staging
  // tiles of the wide epilogue
  // (the wide ReLU-backward variant keeps 5 stages: it has no bias staging area, see AUX_BYTES)
  static constexpr int STAGES = (PAIR ? 6 : (BN == 256 ? 4 : (BN == 128 ? 6 : 8))) - (HAS_MASK ? 1 : 0) -
                                ((WIDE && !HAS_MASK) ? 1 : 0);
  static constexpr int TMEM_COLS = 2 * BN;    // 2 accumulator stages x BN fp32 columns
  static constexpr int EPI_BYTES = HAS_MASK ? EW * 2 * FAST_TILE_BYTES : (WIDE ? EW * FAST_TILE_BYTES : EW * STAGE_BYTES_PER_WARP);
  static_assert(WIDE || (EPI_BYTES >= EW * STAGE_BYTES_PER_WARP && EPI_BYTES >= EW * FAST_TILE_BYTES), "epilogue tiles");
  // ReLU-backward launches of the wide variant take the lean epilogue only with bias == NULL: no bias staging area
  static constexpr int AUX_BYTES = (WIDE && HAS_MASK) ? 0 : BIAS_SMEM_FLOATS * 4;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + BAR_BYTES + EPI_BYTES + AUX_BYTES;
  static_assert(SMEM <= 232448, "shared memory budget");
};

struct GemmParams {
  int M, N, K;             // output rows, output cols, contraction length
  int layout;              // 0 NT, 1 TN
  int n_terms;             // accumulating passes over the contraction (1..LBX_GEMM_MAX_TERMS)
  int term_a[LBX_GEMM_MAX_TERMS], term_b[LBX_GEMM_MAX_TERMS];   // which A / B tensor map each pass reads (0 or 1)
  int term_arow[LBX_GEMM_MAX_TERMS];   // row offset added to the A coordinate of each pass (NT / NN layouts)
  int term_brow[LBX_GEMM_MAX_TERMS];   // row offset added to the B row coordinate of each pass (NT: N index, NN: K index)
  int term_ncol[LBX_GEMM_MAX_TERMS];   // > 0: the pass only contributes to output columns < this (a multiple of the tile width)
  int k_splits;            // split-K partitions (>= 1)
  int epi_atomic;          // 1: atomicAdd fp32 into out (split-K / gradient accumulation)
  int out_dtype;           // LBX_F32 / LBX_BF16
  void* out;
  void* out_lo;            // optional bf16 residual plane (x - bf16(x)); out_dtype must be BF16
  long long ldo;           // output row pitch in elements (may be < N for the overlapping dgrad view)
  const float* bias;       // [N] or NULL
  int relu;
  const float* post_scale; // optional per-column affine AFTER the activation (inference-time BatchNormalization)
  const float* post_shift;
  int rows_per_utt;        // > 0: rows with (m % rows_per_utt) >= valid_rows are not stored
  int valid_rows;
  const __nv_bfloat16* mask_src;   // optional: keep x only where mask_src[m*ldo + n] > 0 (ReLU backward)
  int accumulate;          // 1: out = out + x (read-modify-write, non-atomic)
  float* colsum;           // optional: colsum[n % colsum_mod] += sum_m x[m, n] (bias gradient of the layer below)
  int colsum_mod;
  int fast;                // 1: bf16 output through the lean epilogue (TMA stores / TMA mask loads, prefetched TMEM loads)
};

// ------------------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------------------
// load 32 consecutive bf16 (16-byte vectors when possible) as floats
__device__ __forceinline__ void load_bf16x32(const __nv_bfloat16* src, bool vec, int ncols, float (&f)[32]) {
  if (vec) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 u = reinterpret_cast<const uint4*>(src)[q];
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        f[8 * q + 2 * i] = __uint_as_float(w[i] << 16);
        f[8 * q + 2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = j < ncols ? __bfloat162float(src[j]) : 0.0f;
  }
}

// WIDE: 16 epilogue warps instead of 8 (four per TMEM sub-partition, a quarter of the tile's columns each) for the
// launches whose pace is set by the epilogue (small K: the main loop of a tile is shorter than draining it).  Only the
// lean bf16 epilogue exists in this variant (640 threads leave 102 registers per thread), without the one-chunk-ahead
// TMEM prefetch: with four warps per scheduler the other warps hide that latency.
template <int LAYOUT, int BN, bool HAS_MASK, bool PAIR, bool WIDE = false>
__global__ void __launch_bounds__(WIDE ? GEMM_THREADS_WIDE : GEMM_THREADS, 1)
    gemm_bf16_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                     const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
                     const __grid_constant__ CUtensorMap mapOut, const __grid_constant__ CUtensorMap mapMask,
                     const GemmParams p) {
  using C = Cfg<BN, PAIR, HAS_MASK, WIDE>;
  constexpr int EW = C::EW;                      // epilogue warps
  constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, TMEM_COLS = C::TMEM_COLS;
  constexpr int BN_LOCAL = PAIR ? BN / 2 : BN;   // B columns staged by this CTA
  constexpr bool A_MN = LAYOUT == 1;            // A stored [K, M] (contraction index is the slow one)
  constexpr bool B_MN = LAYOUT != 0;            // B stored [K, N]
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* mask_bar = tempty_bar + 2;          // one per epilogue warp (fast epilogue: mask tile landed)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(mask_bar + EW);
  static_assert((2 * 8 + 4 + EW) * 8 + 4 <= BAR_BYTES, "barrier area");
  unsigned char* s_stage = smem + (size_t)STAGES * STAGE_BYTES + BAR_BYTES;        // 1024-byte aligned epilogue tiles
  float* s_bias = reinterpret_cast<float*>(s_stage + C::EPI_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int cta_rank = PAIR ? (int)cluster_ctarank() : 0;      // rank inside the CTA pair (0 = leader, issues the MMAs)
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, n_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
  const int m_units = PAIR ? (m_tiles + 1) / 2 : m_tiles;     // a pair owns two consecutive 128-row tiles
  const int mn_tiles = m_units * n_tiles;
  const int total_tiles = mn_tiles * p.k_splits;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_per_split = (kb_total + p.k_splits - 1) / p.k_splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapB0);
    if (p.n_terms > 1) {
      tma_prefetch_desc(&mapA1);
      tma_prefetch_desc(&mapB1);
    }
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar + i, PAIR ? 2 : 1);  // pair: the producers of both CTAs arrive on the leader's barrier
      mbar_init(empty_bar + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar + i, 1);
      mbar_init(tempty_bar + i, PAIR ? 2 * EW : EW);   // one arrival per epilogue warp (of both CTAs)
    }
    for (int i = 0; i < EW; ++i) mbar_init(mask_bar + i, 1);
    if (p.fast) {
      tma_prefetch_desc(&mapOut);
      if (HAS_MASK) tma_prefetch_desc(&mapMask);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                 // the peer's barriers are initialised before anyone arrives remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // programmatic dependent launch: everything above overlapped the tail of the previous kernel in the stream;
  // from here on global memory written by it is read (and memory it reads is overwritten)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = worker; tile < total_tiles; tile += n_workers) {
        const int split = tile / mn_tiles, rem = tile - split * mn_tiles;
        const int m_unit = rem / n_tiles, n_blk = rem - m_unit * n_tiles;
        const int m_blk = PAIR ? 2 * m_unit + cta_rank : m_unit;
        const int n_col0 = n_blk * BN + cta_rank * BN_LOCAL;       // first B column staged by this CTA
        const int kb0 = split * kb_per_split, kb1 = min(kb_total, kb0 + kb_per_split);
        for (int term = 0; term < p.n_terms; ++term) {
          const CUtensorMap* mA = p.term_a[term] ? &mapA1 : &mapA0;
          const CUtensorMap* mB = p.term_b[term] ? &mapB1 : &mapB0;
          if (p.term_ncol[term] > 0 && n_blk * BN >= p.term_ncol[term]) continue;   // pass does not reach these columns
          const int arow = p.term_arow[term], brow = p.term_brow[term];
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(empty_bar + stage, phase ^ 1);
            unsigned char* sA = smem + (size_t)stage * STAGE_BYTES;
            unsigned char* sB = sA + A_BYTES;
            if (!PAIR) {
              mbar_expect_tx(full_bar + stage, (uint32_t)STAGE_BYTES);
              if (!A_MN) {
                tma_load_2d(mA, full_bar + stage, sA, kb * BK, m_blk * BM + arow);
              } else {
#pragma unroll
                for (int i = 0; i < BM / 64; ++i)
                  tma_load_2d(mA, full_bar + stage, sA + i * 8192, m_blk * BM + i * 64, kb * BK);
              }
              if (!B_MN) {
                tma_load_2d(mB, full_bar + stage, sB, kb * BK, n_blk * BN + brow);
              } else {
#pragma unroll
                for (int i = 0; i < BN / 64; ++i)
                  tma_load_2d(mB, full_bar + stage, sB + i * 8192, n_blk * BN + i * 64, kb * BK + brow);
              }
            } else {
              // all bytes of both CTAs are counted on the LEADER's full barrier (the MMA issuer waits there)
              const uint32_t lead_full = mapa_u32(smem_u32(full_bar + stage), 0);
              if (cta_rank == 0) mbar_expect_tx(full_bar + stage, 2u * (uint32_t)STAGE_BYTES);
              else mbar_arrive_remote(lead_full);
              if (!A_MN) {
                tma_load_2d_pair(mA, lead_full, sA, kb * BK, m_blk * BM + arow);
              } else {
#pragma unroll
                for (int i = 0; i < BM / 64; ++i)
                  tma_load_2d_pair(mA, lead_full, sA + i * 8192, m_blk * BM + i * 64, kb * BK);
              }
              if (!B_MN) {
                tma_load_2d_pair(mB, lead_full, sB, kb * BK, n_col0 + brow);
              } else {
#pragma unroll
                for (int i = 0; i < BN_LOCAL / 64; ++i)
                  tma_load_2d_pair(mB, lead_full, sB + i * 8192, n_col0 + i * 64, kb * BK + brow);
              }
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc<BN, PAIR ? 256 : 128>(A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = worker; tile < total_tiles; tile += n_workers) {
        const int split = tile / mn_tiles;
        const int kb0 = split * kb_per_split, kb1 = min(kb_total, kb0 + kb_per_split);
        const int n_blk_t = (tile - split * mn_tiles) % n_tiles;
        int terms_here = 0;                         // passes that reach this tile's columns (same test as the producer)
        for (int term = 0; term < p.n_terms; ++term)
          terms_here += (p.term_ncol[term] <= 0 || n_blk_t * BN < p.term_ncol[term]) ? 1 : 0;
        const int iters = (kb1 - kb0) * terms_here;
        mbar_wait(tempty_bar + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * BN;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + (size_t)stage * STAGE_BYTES);
          const uint32_t sB = sA + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: 8-row groups are 1024 B apart; advance 32 B per UMMA_K inside the swizzle atom
            // MN-major: 64-element blocks are 8192 B apart (LBO), 8 k-rows = 1024 B (SBO), 2048 B per UMMA_K
            const uint64_t da = A_MN ? make_smem_desc(sA + k * (UMMA_K * 128), 8192, 1024)
                                     : make_smem_desc(sA + k * (UMMA_K * 2), 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc(sB + k * (UMMA_K * 128), 8192, 1024)
                                     : make_smem_desc(sB + k * (UMMA_K * 2), 16, 1024);
            if (PAIR) tc_mma_bf16_pair(tmem_d, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
            else tc_mma_bf16(tmem_d, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          if (PAIR) tc_commit_pair(empty_bar + stage);   // frees the stage in both CTAs when these MMAs retire
          else tc_commit(empty_bar + stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (PAIR) tc_commit_pair(tfull_bar + acc);      // accumulators (in both CTAs' TMEM) ready for the epilogues
        else tc_commit(tfull_bar + acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= LBX_CTRL_WARPS) {
    // ===================================== epilogue =====================================
    const int sub = warp & 3;                    // TMEM sub-partition this warp may read: lanes [32*sub, 32*sub+32)
    const int chalf = (warp - LBX_CTRL_WARPS) >> 2;           // which part of the tile's columns this warp drains
    constexpr int CPARTS = EW / 4;               // warps sharing a sub-partition (2, wide epilogue: 4)
    constexpr int WCOLS = BN / CPARTS;           // columns per warp
    constexpr int CHUNKS = WCOLS / 32;           // 32-column chunks per warp
    int acc = 0;
    uint32_t acc_phase = 0;
    // the bias vector is read by every tile: stage it once (global loads in the epilogue's critical path cost an L2
    // round trip per 32-column chunk with only two warps per scheduler to hide it)
    const bool bias_smem = LBX_BIAS_SMEM && p.bias != nullptr && p.N <= BIAS_SMEM_FLOATS;
    if (bias_smem) {      // zero-filled up to the next multiple of 32 so that the last (partial) chunk needs no guard
      for (int i = threadIdx.x - 32 * LBX_CTRL_WARPS; i < ((p.N + 31) & ~31) && i < BIAS_SMEM_FLOATS; i += 32 * EW)
        s_bias[i] = i < p.N ? __ldg(p.bias + i) : 0.0f;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
    }
    uint32_t mphase = 0;                         // parity of this warp's mask barrier (fast epilogue)
    for (int tile = worker; tile < total_tiles; tile += n_workers) {
      const int rem = tile % mn_tiles;
      // split-K partial tiles are summed atomically: the bias is added by the first split only
      const bool add_bias = !p.epi_atomic || tile < mn_tiles;
      const int m_unit = rem / n_tiles, n_blk = rem - m_unit * n_tiles;
      const int m_blk = PAIR ? 2 * m_unit + cta_rank : m_unit;
      const int m = m_blk * BM + sub * 32 + lane;
      const bool in_range = m < p.M;
      // rows past the valid part of an utterance are stored as zeros: they land on rows of the destination that
      // must stay zero anyway (junk rows / the next utterance's causal padding)
      const bool row_zero = p.rows_per_utt > 0 && ((m % p.rows_per_utt) >= p.valid_rows);
      const long long row_off = (long long)m * p.ldo;
      const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * BN + chalf * WCOLS);
      const int col_base = n_blk * BN + chalf * WCOLS;
      // coalescing through a per-warp staging tile: a thread owns one accumulator ROW, so direct 16-byte accesses touch
      // 32 different lines per instruction (half sectors); staged, 4 lanes cover 64 contiguous bytes of one row
      unsigned char* st = s_stage + (warp - LBX_CTRL_WARPS) * STAGE_BYTES_PER_WARP;
      const int warp_row0 = m_blk * BM + sub * 32;
      const int sr = lane >> 2, sseg = lane & 3;           // staged access: rows sr + 8j, 16-byte segment sseg
      const bool ld_al = (p.ldo & 7) == 0;
      const bool stage_out = p.out_dtype == LBX_BF16 && !p.epi_atomic && !p.accumulate && ld_al &&
                             ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                             (p.out_lo == nullptr || (reinterpret_cast<uintptr_t>(p.out_lo) & 15) == 0);
      const bool stage_mask = HAS_MASK && ld_al && ((reinterpret_cast<uintptr_t>(p.mask_src) & 15) == 0);
      const __nv_bfloat16* mrow = HAS_MASK ? p.mask_src + row_off : nullptr;
      const __nv_bfloat16* prow = reinterpret_cast<const __nv_bfloat16*>(p.out) + row_off;
      const bool mvec = HAS_MASK && ((reinterpret_cast<uintptr_t>(mrow) & 15) == 0);
      const bool pvec = p.accumulate && p.out_dtype == LBX_BF16 && !p.epi_atomic && in_range &&
                        ((reinterpret_cast<uintptr_t>(prow) & 15) == 0);
      if (p.fast) {
        // ---------------- lean bf16 epilogue ----------------
        // per 32-column chunk: TMEM load (the next chunk's load is already in flight) -> bias -> ReLU-backward mask
        // (tile fetched by a bulk tensor load one chunk ahead) -> zero junk rows -> bf16 (ReLU folded into the
        // conversion) -> swizzled shared-memory tile -> one bulk tensor store (clipped at M / N by the tensor map)
        const int wi = warp - LBX_CTRL_WARPS;
        unsigned char* st_out = s_stage + wi * FAST_TILE_BYTES;
        unsigned char* st_msk = s_stage + (EW + wi) * FAST_TILE_BYTES;
        uint64_t* mbar_m = mask_bar + wi;
        const int ncols_rem = p.N - col_base;
        const int nch = (ncols_rem <= 0 || warp_row0 >= p.M) ? 0 : min(CHUNKS, (ncols_rem + 31) >> 5);
        const uint32_t sw = (uint32_t)((lane >> 1) & 3);           // 64-byte swizzle: 16-byte chunk index ^= (row / 2) % 4
        const bool kill_row = row_zero || !in_range;
        if (HAS_MASK && nch > 0 && lane == 0) {
          mbar_expect_tx(mbar_m, FAST_TILE_BYTES);
          tma_load_2d(&mapMask, mbar_m, st_msk, col_base, warp_row0);
        }
        mbar_wait(tfull_bar + acc, acc_phase);
        tc_fence_after();
        uint32_t v[32];
        if (!WIDE && nch > 0) tc_ld32(taddr, v);
        bool released = false;
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
          const int n0 = col_base + c * 32;
          if (WIDE) tc_ld32(taddr + c * 32, v);            // no prefetch: four warps per scheduler hide the latency
          tc_wait_ld();
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
          if (c + 1 < nch) {
            if (!WIDE) tc_ld32(taddr + (c + 1) * 32, v);
          } else {
            // every TMEM read of this warp has landed: hand the accumulator stage back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR) mbar_arrive_remote(mapa_u32(smem_u32(tempty_bar + acc), 0));
              else mbar_arrive(tempty_bar + acc);
            }
            released = true;
          }
          if (bias_smem) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 b4 = *reinterpret_cast<const float4*>(s_bias + n0 + 4 * q);
              x[4 * q] += b4.x; x[4 * q + 1] += b4.y; x[4 * q + 2] += b4.z; x[4 * q + 3] += b4.w;
            }
          }
          if (HAS_MASK) {
            mbar_wait(mbar_m, mphase);
            mphase ^= 1;
            uint4 mk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
              mk[q] = *reinterpret_cast<const uint4*>(st_msk + lane * 64 + (((uint32_t)q ^ sw) << 4));
            __syncwarp();                                          // all lanes have read the tile: refill it
            if (c + 1 < nch && lane == 0) {
              mbar_expect_tx(mbar_m, FAST_TILE_BYTES);
              tma_load_2d(&mapMask, mbar_m, st_msk, n0 + 32, warp_row0);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t w[4] = {mk[q].x, mk[q].y, mk[q].z, mk[q].w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {     // keep x where the bf16 mask value is > 0 (sign clear, magnitude non-zero)
                if ((int)(w[i] << 16) <= 0) x[8 * q + 2 * i] = 0.0f;
                if ((int)(w[i] & 0xFFFF0000u) <= 0) x[8 * q + 2 * i + 1] = 0.0f;
              }
            }
          }
          if (kill_row) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = 0.0f;
          }
          uint32_t pk[16];
          if (p.post_scale != nullptr) {          // activation, then the per-column affine (zero rows stay zero)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float v = p.relu ? fmaxf(x[j], 0.0f) : x[j];
              const int nn = n0 + j < p.N ? n0 + j : p.N - 1;
              v = fmaf(v, __ldg(p.post_scale + nn), __ldg(p.post_shift + nn));
              x[j] = kill_row ? 0.0f : v;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = cvt_bf16x2(x[2 * j], x[2 * j + 1]);
          } else if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = cvt_bf16x2_relu(x[2 * j], x[2 * j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = cvt_bf16x2(x[2 * j], x[2 * j + 1]);
          }
          if (lane == 0) bulk_wait_read0();                        // the previous store has finished reading the tile
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(st_out + lane * 64 + (((uint32_t)q ^ sw) << 4)) =
                make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&mapOut, smem_u32(st_out), n0, warp_row0);
            bulk_commit();
          }
          if (p.colsum != nullptr) {             // host guarantees relu == 0 here: x is what was stored (before rounding)
            // (tried in round 2: reading the staged bf16 tile back from shared memory, 16 loads + one shuffle per lane
            // instead of this 31-shuffle transpose — measured SLOWER, tensor-core replay 0.263 -> 0.295 ms per step)
            const int ncols = min(32, p.N - n0);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
              const bool up = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < off; ++i) {
                const float send = up ? x[i] : x[i + off];
                const float recv = __shfl_xor_sync(0xffffffffu, send, off);
                x[i] = (up ? x[i + off] : x[i]) + recv;
              }
            }
            // (also tried: per-warp shared-memory accumulators over all tiles of the CTA and ONE global atomic per column
            // and CTA at the end — no gain: the cost of this block is the transpose, not the atomics; tools/dgrad_probe.py)
            if (lane < ncols) atomicAdd(p.colsum + (n0 + lane) % p.colsum_mod, x[0]);
          }
        }
        if (!released) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_remote(mapa_u32(smem_u32(tempty_bar + acc), 0));
            else mbar_arrive(tempty_bar + acc);
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      if constexpr (WIDE) continue;                // (the host launches the wide variant for lean-epilogue problems only)
      mbar_wait(tfull_bar + acc, acc_phase);
      tc_fence_after();
#ifndef LBX_EPI_GROUP
#define LBX_EPI_GROUP 1
#endif
      constexpr int GROUP = LBX_EPI_GROUP, NGROUPS = CHUNKS / GROUP;     // 32*GROUP columns in flight per warp
#pragma unroll 1
      for (int grp = 0; grp < NGROUPS; ++grp) {
      uint32_t v[GROUP][32];
#pragma unroll
      for (int c = 0; c < GROUP; ++c)
        if (col_base + (grp * GROUP + c) * 32 < p.N) tc_ld32(taddr + (grp * GROUP + c) * 32, v[c]);   // warp-uniform
      tc_wait_ld();
      if (grp == NGROUPS - 1) {
        // all TMEM reads of this warp are done: release the accumulator stage before touching global memory
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_remote(mapa_u32(smem_u32(tempty_bar + acc), 0));   // the leader's MMA issuer waits
          else mbar_arrive(tempty_bar + acc);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (m_blk * BM + sub * 32 >= p.M && p.colsum == nullptr) continue;   // whole warp past the last row
#pragma unroll
      for (int c = 0; c < GROUP; ++c) {
        const int ci = grp * GROUP + c;
        const int n0 = col_base + ci * 32;
        if (n0 >= p.N) break;
        const int ncols = min(32, p.N - n0);
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[c][j]);
        if (!add_bias) {
          // (a later split of an atomically accumulated tile)
        } else if (bias_smem) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (4 * q < ncols) {                 // N is a multiple of 4 for every biased layer that takes this path
              const float4 b4 = *reinterpret_cast<const float4*>(s_bias + n0 + 4 * q);
              x[4 * q] += b4.x; x[4 * q + 1] += b4.y; x[4 * q + 2] += b4.z; x[4 * q + 3] += b4.w;
            }
          }
        } else if (p.bias != nullptr) {
          const float* bp = p.bias + n0;
          if (ncols == 32 && ((reinterpret_cast<uintptr_t>(bp) & 15) == 0)) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + q);
              x[4 * q] += b4.x; x[4 * q + 1] += b4.y; x[4 * q + 2] += b4.z; x[4 * q + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) x[j] += __ldg(bp + j);
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.0f);
        }
        if (p.post_scale != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncols) x[j] = fmaf(x[j], __ldg(p.post_scale + n0 + j), __ldg(p.post_shift + n0 + j));
        }
        if (row_zero || !in_range) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = 0.0f;
        }
        if (HAS_MASK && stage_mask && ncols == 32) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = sr + 8 * j;
            uint4 val = make_uint4(0u, 0u, 0u, 0u);
            if (warp_row0 + r < p.M)
              val = __ldg(reinterpret_cast<const uint4*>(p.mask_src + (long long)(warp_row0 + r) * p.ldo + n0) + sseg);
            *reinterpret_cast<uint4*>(st + r * STAGE_PITCH + sseg * 16) = val;
          }
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 u = *reinterpret_cast<const uint4*>(st + lane * STAGE_PITCH + q * 16);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {       // bf16 > 0  <=>  sign clear and magnitude non-zero
              const uint32_t lo16 = w[i] & 0xFFFFu, hi16 = w[i] >> 16;
              if (lo16 == 0u || lo16 >= 0x8000u) x[8 * q + 2 * i] = 0.0f;
              if (hi16 == 0u || hi16 >= 0x8000u) x[8 * q + 2 * i + 1] = 0.0f;
            }
          }
          __syncwarp();
        } else if (HAS_MASK && in_range) {
          float mk[32];
          load_bf16x32(mrow + n0, mvec && ncols == 32, ncols, mk);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (!(mk[j] > 0.0f)) x[j] = 0.0f;
        }
        if (stage_out && ncols == 32) {
#pragma unroll
          for (int plane = 0; plane < 2; ++plane) {
            if (plane == 1 && p.out_lo == nullptr) break;
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(plane == 0 ? p.out : p.out_lo);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float y[8];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                y[i] = plane == 0 ? x[8 * q + i] : x[8 * q + i] - __bfloat162float(__float2bfloat16_rn(x[8 * q + i]));
              *reinterpret_cast<uint4*>(st + lane * STAGE_PITCH + q * 16) =
                  make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]),
                             pack_bf16x2(y[6], y[7]));
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int r = sr + 8 * j;
              if (warp_row0 + r < p.M)
                *(reinterpret_cast<uint4*>(dst + (long long)(warp_row0 + r) * p.ldo + n0) + sseg) =
                    *reinterpret_cast<const uint4*>(st + r * STAGE_PITCH + sseg * 16);
            }
            __syncwarp();
          }
        } else if (!in_range) {
          // nothing to store; the lane only takes part in the column-sum reduction below
        } else if (p.epi_atomic) {
          float* o = reinterpret_cast<float*>(p.out) + row_off + n0;
          if (ncols == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * q), "f"(x[4 * q]),
                           "f"(x[4 * q + 1]), "f"(x[4 * q + 2]), "f"(x[4 * q + 3])
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) atomicAdd(o + j, x[j]);
          }
        } else if (p.out_dtype == LBX_F32) {
          float* o = reinterpret_cast<float*>(p.out) + row_off + n0;
          const bool vec = ncols == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
          if (p.accumulate) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) x[j] += o[j];
          }
          if (vec) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              reinterpret_cast<float4*>(o)[q] = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) o[j] = x[j];
          }
        } else {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + row_off + n0;
          const bool vec = ncols == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
          float prev[32];
          if (p.accumulate) {
            load_bf16x32(o, pvec && ncols == 32, ncols, prev);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) prev[j] = 0.0f;
          }
          if (vec) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              reinterpret_cast<uint4*>(o)[q] = make_uint4(
                  pack_bf16x2(x[8 * q] + prev[8 * q], x[8 * q + 1] + prev[8 * q + 1]),
                  pack_bf16x2(x[8 * q + 2] + prev[8 * q + 2], x[8 * q + 3] + prev[8 * q + 3]),
                  pack_bf16x2(x[8 * q + 4] + prev[8 * q + 4], x[8 * q + 5] + prev[8 * q + 5]),
                  pack_bf16x2(x[8 * q + 6] + prev[8 * q + 6], x[8 * q + 7] + prev[8 * q + 7]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) o[j] = __float2bfloat16_rn(x[j] + prev[j]);
          }
          if (p.out_lo != nullptr) {
            __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(p.out_lo) + row_off + n0;
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] -= __bfloat162float(__float2bfloat16_rn(x[j]));
            if (vec && ((reinterpret_cast<uintptr_t>(ol) & 15) == 0)) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                reinterpret_cast<uint4*>(ol)[q] =
                    make_uint4(pack_bf16x2(x[8 * q], x[8 * q + 1]), pack_bf16x2(x[8 * q + 2], x[8 * q + 3]),
                               pack_bf16x2(x[8 * q + 4], x[8 * q + 5]), pack_bf16x2(x[8 * q + 6], x[8 * q + 7]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols) ol[j] = __float2bfloat16_rn(x[j]);
            }
          }
        }
        if (p.colsum != nullptr) {
          // transpose-reduce over the 32 rows of this warp: lane j ends up with the sum of column n0 + j
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; ++i) {
              const float send = up ? x[i] : x[i + off];
              const float recv = __shfl_xor_sync(0xffffffffu, send, off);
              x[i] = (up ? x[i + off] : x[i]) + recv;
            }
          }
          if (lane < ncols) atomicAdd(p.colsum + (n0 + lane) % p.colsum_mod, x[0]);
        }
      }
      }
    }
  }

  if (p.fast && warp >= LBX_CTRL_WARPS && lane == 0) bulk_wait0();   // bulk stores complete before shared memory goes away
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                 // the peer may still be reading this CTA's shared memory / barriers
  if (warp == 1) {
    tc_fence_after();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

// 2-D bf16 tensor map: inner extent `cols` (pitch 1), outer extent `rows` with pitch `ld` elements (ld may be < cols)
int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_cols,
             int box_rows, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(LBX_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(LBX_ECUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld base=%p", (int)r, rows,
                     cols, ld, base);
  return LBX_OK;
}

static int g_num_sms = 0;
static int g_use_pair = 1;
static int g_use_fast_epi = 1;
static int g_use_wide_epi = 1;
static int g_wide_max_kb = 64;

}  // namespace lbx

using namespace lbx;

extern "C" int lbx_gemm_bf16(const lbx_gemm_t* g, void* stream) {
  LBX_CHECK_ARG(g != nullptr, "NULL gemm descriptor");
  LBX_CHECK_ARG(g->layout >= 0 && g->layout <= 2, "layout must be 0 (NT), 1 (TN) or 2 (NN)");
  LBX_CHECK_ARG(g->n_terms >= 1 && g->n_terms <= LBX_GEMM_MAX_TERMS, "n_terms must be in [1, %d]", LBX_GEMM_MAX_TERMS);
  LBX_CHECK_ARG(g->a_rows >= 0 && g->a_cols >= 0 && g->b_rows >= 0 && g->b_cols >= 0, "negative extent");
  GemmParams p{};
  p.layout = g->layout;
  if (g->layout == 0) {
    LBX_CHECK_ARG(g->a_cols == g->b_cols, "NT: A and B must share the contraction length (%d vs %d)", g->a_cols,
                  g->b_cols);
    LBX_CHECK_ARG(g->a_rows <= 2147483647LL && g->b_rows <= 2147483647LL, "extent too large");
    p.M = (int)g->a_rows; p.N = (int)g->b_rows; p.K = g->a_cols;
  } else if (g->layout == 1) {
    LBX_CHECK_ARG(g->a_rows == g->b_rows, "TN: A and B must share the contraction length (%lld vs %lld)", g->a_rows,
                  g->b_rows);
    LBX_CHECK_ARG(g->a_rows <= 2147483647LL, "extent too large");
    p.M = g->a_cols; p.N = g->b_cols; p.K = (int)g->a_rows;
  } else {
    LBX_CHECK_ARG((long long)g->a_cols == g->b_rows, "NN: A and B must share the contraction length (%d vs %lld)",
                  g->a_cols, g->b_rows);
    LBX_CHECK_ARG(g->a_rows <= 2147483647LL, "extent too large");
    p.M = (int)g->a_rows; p.N = g->b_cols; p.K = g->a_cols;
  }
  if (p.M == 0 || p.N == 0) return LBX_OK;
  LBX_CHECK_ARG(p.K > 0, "empty contraction");
  LBX_CHECK_ARG(g->a0 && g->b0 && g->out, "NULL operand");
  for (int t = 0; t < g->n_terms; ++t) {
    LBX_CHECK_ARG((g->term_a[t] == 0 || (g->term_a[t] == 1 && g->a1)) && (g->term_b[t] == 0 || (g->term_b[t] == 1 && g->b1)),
                  "term %d selects an operand plane that was not given", t);
    LBX_CHECK_ARG(g->term_a_row[t] == 0 || g->layout != 1, "row offsets are not available in the TN layout");
    LBX_CHECK_ARG(g->term_b_row[t] == 0 || g->layout != 1, "row offsets are not available in the TN layout");
    // rows of the B view past b_map_rows read as zeros (TMA out-of-bounds fill): a pass may hang over the end of the view
    LBX_CHECK_ARG(g->term_b_row[t] >= 0 && g->term_b_row[t] < (g->b_map_rows > 0 ? g->b_map_rows : g->b_rows),
                  "term %d: B row offset %d is outside the B view", t, g->term_b_row[t]);
    p.term_a[t] = g->term_a[t]; p.term_b[t] = g->term_b[t]; p.term_arow[t] = g->term_a_row[t];
    p.term_brow[t] = g->term_b_row[t];
    LBX_CHECK_ARG(g->term_col_limit[t] >= 0 && g->term_col_limit[t] % 256 == 0 && (t > 0 || g->term_col_limit[t] == 0),
                  "term_col_limit must be a multiple of 256 (the widest tile) and 0 for the first pass");
    p.term_ncol[t] = g->term_col_limit[t];
  }
  LBX_CHECK_ARG(g->lda % 8 == 0 && g->ldb % 8 == 0, "operand pitches must be multiples of 8 elements (16 bytes)");
  LBX_CHECK_ARG((reinterpret_cast<uintptr_t>(g->a0) & 15) == 0 && (reinterpret_cast<uintptr_t>(g->b0) & 15) == 0,
                "operands must be 16-byte aligned");
  LBX_CHECK_ARG(g->out_dtype == LBX_F32 || g->out_dtype == LBX_BF16, "bad out_dtype");
  LBX_CHECK_ARG(!g->epi_atomic || g->out_dtype == LBX_F32, "atomic epilogue needs an fp32 output");
  LBX_CHECK_ARG(g->out_lo == nullptr || g->out_dtype == LBX_BF16, "out_lo needs a bf16 output");
  LBX_CHECK_ARG(g->k_splits >= 1, "k_splits must be >= 1");
  LBX_CHECK_ARG(g->k_splits == 1 || g->epi_atomic, "split-K needs the atomic epilogue");
  LBX_CHECK_ARG(!(g->epi_atomic && g->relu), "an activation cannot be applied to atomically accumulated partial sums");
  p.n_terms = g->n_terms;
  const int kb_total = (p.K + BK - 1) / BK;
  int ks = g->k_splits < kb_total ? g->k_splits : kb_total;
  const int per = (kb_total + ks - 1) / ks;
  ks = (kb_total + per - 1) / per;      // no empty split
  p.k_splits = ks;
  p.epi_atomic = g->epi_atomic;
  p.out_dtype = g->out_dtype;
  p.out = g->out; p.out_lo = g->out_lo; p.ldo = g->ldo;
  p.bias = g->bias; p.relu = g->relu;
  LBX_CHECK_ARG((g->post_scale == nullptr) == (g->post_shift == nullptr), "post_scale and post_shift come together");
  LBX_CHECK_ARG(g->post_scale == nullptr || (!g->epi_atomic && !g->accumulate && g->mask_src == nullptr && g->colsum == nullptr),
                "the post-activation affine is a forward-pass epilogue (no atomics / accumulate / mask / colsum)");
  p.post_scale = g->post_scale; p.post_shift = g->post_shift;
  p.rows_per_utt = g->rows_per_utt; p.valid_rows = g->valid_rows;
  p.mask_src = reinterpret_cast<const __nv_bfloat16*>(g->mask_src);
  p.accumulate = g->accumulate;
  p.colsum = g->colsum;
  p.colsum_mod = g->colsum_mod > 0 ? g->colsum_mod : 1;
  LBX_CHECK_ARG(!(g->colsum && (g->out_lo || g->epi_atomic)), "colsum cannot be combined with out_lo / atomic epilogues");

  CUtensorMap mA0, mA1, mB0, mB1, mOut, mMask;
  int rc;
  // lean epilogue: bf16 result, plain store (no atomics / read-modify-write / residual plane), 16-byte aligned rows
  const bool al16 = (g->ldo % 8 == 0) && (reinterpret_cast<uintptr_t>(g->out) & 15) == 0;
  p.fast = g_use_fast_epi && g->out_dtype == LBX_BF16 && !g->epi_atomic && !g->accumulate && g->out_lo == nullptr &&
           al16 && (g->bias == nullptr || ((p.N + 31) & ~31) <= BIAS_SMEM_FLOATS) && !(g->colsum && g->relu) &&
           (g->mask_src == nullptr || (reinterpret_cast<uintptr_t>(g->mask_src) & 15) == 0);
  // tile-N 256 unless the problem is narrower than 128 columns (measured: 128 never wins on the TDNN shapes)
  const int bn = (g->tile_n == 64 || g->tile_n == 128 || g->tile_n == 256) ? g->tile_n : (p.N <= 64 ? 64 : (p.N <= 128 ? 128 : 256));
  // CTA pairs (tcgen05 cta_group::2) for the 256-wide tiles: every CTA stages only half of the B tile
  const bool pair = g_use_pair && bn == 256;
  const int boxA_rows = g->layout == 1 ? 64 : BM, boxB_rows = g->layout == 0 ? (pair ? bn / 2 : bn) : 64;
  if ((rc = make_map(&mA0, g->a0, g->a_rows, g->a_cols, g->lda, 64, boxA_rows))) return rc;
  const long long b_map_rows = g->b_map_rows > 0 ? g->b_map_rows : g->b_rows;
  if ((rc = make_map(&mB0, g->b0, b_map_rows, g->b_cols, g->ldb, 64, boxB_rows))) return rc;
  mA1 = mA0; mB1 = mB0;
  if (g->a1 && (rc = make_map(&mA1, g->a1, g->a_rows, g->a_cols, g->lda, 64, boxA_rows))) return rc;
  if (g->b1 && (rc = make_map(&mB1, g->b1, b_map_rows, g->b_cols, g->ldb, 64, boxB_rows))) return rc;
  mOut = mA0; mMask = mA0;
  if (p.fast) {
    if ((rc = make_map(&mOut, g->out, p.M, p.N, g->ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if (g->mask_src && (rc = make_map(&mMask, g->mask_src, p.M, p.N, g->ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)))
      return rc;
  }
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    LBX_CUDA(cudaGetDevice(&dev));
    LBX_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
#define LBX_SET_SMEM(L, N_)                                                                                        \
  LBX_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<L, N_, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)Cfg<N_>::SMEM));                                                              \
  LBX_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<L, N_, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                (int)Cfg<N_, false, true>::SMEM))
#define LBX_SET_SMEM_PAIR(L)                                                                                       \
  LBX_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<L, 256, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)Cfg<256, true>::SMEM));                                                       \
  LBX_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<L, 256, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                (int)Cfg<256, true, true>::SMEM))
    LBX_SET_SMEM(0, 256); LBX_SET_SMEM(1, 256); LBX_SET_SMEM(2, 256);
    LBX_SET_SMEM(0, 128); LBX_SET_SMEM(1, 128); LBX_SET_SMEM(2, 128);
    LBX_SET_SMEM(0, 64); LBX_SET_SMEM(1, 64); LBX_SET_SMEM(2, 64);
    LBX_SET_SMEM_PAIR(0); LBX_SET_SMEM_PAIR(1); LBX_SET_SMEM_PAIR(2);
    // wide-epilogue variants (CTA pairs, lean epilogue only): forward (NN) and ReLU-backward data gradient (NT + mask)
    LBX_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<2, 256, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)Cfg<256, true, false, true>::SMEM));
    LBX_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<0, 256, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)Cfg<256, true, true, true>::SMEM));
#undef LBX_SET_SMEM
#undef LBX_SET_SMEM_PAIR
    g_num_sms = n;
  }
  const long long m_tiles_h = (p.M + BM - 1) / BM;
  const long long tiles = (pair ? (m_tiles_h + 1) / 2 : m_tiles_h) * ((p.N + bn - 1) / bn) * p.k_splits;
  const int max_workers = pair ? g_num_sms / 2 : g_num_sms;
  const int workers = (int)(tiles < max_workers ? tiles : max_workers);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(pair ? 2 * workers : workers));
  cfg.blockDim = dim3(GEMM_THREADS);
  const bool hm = p.mask_src != nullptr;
  cfg.dynamicSmemBytes =
      pair ? (hm ? Cfg<256, true, true>::SMEM : Cfg<256, true>::SMEM)
           : (bn == 256 ? (hm ? Cfg<256, false, true>::SMEM : Cfg<256>::SMEM)
                        : (bn == 128 ? (hm ? Cfg<128, false, true>::SMEM : Cfg<128>::SMEM)
                                     : (hm ? Cfg<64, false, true>::SMEM : Cfg<64>::SMEM)));
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g_use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (pair) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t le;
  // Wide epilogue (16 epilogue warps) for the lean-epilogue CTA-pair launches: every forward / data-gradient GEMM of the
  // TDNN drains its tiles more slowly than it computes them (measured on B200, step time vs the k-block threshold:
  // off 0.411 ms, <= 8: 0.399, <= 16: 0.397, <= 24 (all of them): 0.389), so the threshold only excludes very long
  // contractions, where one pipeline stage more is worth more than the extra warps.
  const int kb_per_tile = ((p.K + BK - 1) / BK) * p.n_terms;
  const bool wide = g_use_wide_epi && pair && p.fast && p.k_splits == 1 && kb_per_tile <= g_wide_max_kb &&
                    ((g->layout == 2 && !hm) || (g->layout == 0 && hm && g->bias == nullptr));
  if (wide) {
    cfg.blockDim = dim3(GEMM_THREADS_WIDE);
    if (hm) {
      cfg.dynamicSmemBytes = Cfg<256, true, true, true>::SMEM;
      le = cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<0, 256, true, true, true>, mA0, mA1, mB0, mB1, mOut, mMask, p);
    } else {
      cfg.dynamicSmemBytes = Cfg<256, true, false, true>::SMEM;
      le = cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<2, 256, false, true, true>, mA0, mA1, mB0, mB1, mOut, mMask, p);
    }
    if (le != cudaSuccess) return set_error(LBX_ECUDA, "GEMM launch failed: %s", cudaGetErrorString(le));
    LBX_LAUNCH_CHECK();
    return LBX_OK;
  }
#define LBX_GEMM_LAUNCH(L, N_, P_)                                                                          \
  le = p.mask_src ? cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<L, N_, true, P_>, mA0, mA1, mB0, mB1, mOut, mMask, p)  \
                  : cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<L, N_, false, P_>, mA0, mA1, mB0, mB1, mOut, mMask, p)
  if (pair) {
    if (g->layout == 0) LBX_GEMM_LAUNCH(0, 256, true); else if (g->layout == 1) LBX_GEMM_LAUNCH(1, 256, true); else LBX_GEMM_LAUNCH(2, 256, true);
  } else if (bn == 256) {
    if (g->layout == 0) LBX_GEMM_LAUNCH(0, 256, false); else if (g->layout == 1) LBX_GEMM_LAUNCH(1, 256, false); else LBX_GEMM_LAUNCH(2, 256, false);
  } else if (bn == 128) {
    if (g->layout == 0) LBX_GEMM_LAUNCH(0, 128, false); else if (g->layout == 1) LBX_GEMM_LAUNCH(1, 128, false); else LBX_GEMM_LAUNCH(2, 128, false);
  } else {
    if (g->layout == 0) LBX_GEMM_LAUNCH(0, 64, false); else if (g->layout == 1) LBX_GEMM_LAUNCH(1, 64, false); else LBX_GEMM_LAUNCH(2, 64, false);
  }
#undef LBX_GEMM_LAUNCH
  if (le != cudaSuccess) return set_error(LBX_ECUDA, "GEMM launch failed: %s", cudaGetErrorString(le));
  LBX_LAUNCH_CHECK();
  return LBX_OK;
}

// lean bf16 epilogue (TMA stores, TMA mask loads, prefetched TMEM loads); LBX_GEMM_FAST_EPI=0 selects the general one
extern "C" int lbx_set_gemm_fast_epilogue(int enabled) {
  g_use_fast_epi = enabled ? 1 : 0;
  return LBX_OK;
}

// 16-warp epilogue for epilogue-bound lean launches (tiles of at most max_kb 64-wide k-blocks); 0 disables
extern "C" int lbx_set_gemm_wide_epilogue(int enabled, int max_kb) {
  g_use_wide_epi = enabled ? 1 : 0;
  if (max_kb > 0) g_wide_max_kb = max_kb;
  return LBX_OK;
}

// CTA-pair (tcgen05 cta_group::2) execution of the 256-wide tiles; LBX_GEMM_PAIR in the Python host's environment
extern "C" int lbx_set_gemm_pair(int enabled) {
  g_use_pair = enabled ? 1 : 0;
  return LBX_OK;
}
