// Fused dense head of the x-vector (sm_100a): the segment layers (lidbox/models/xvector.py:61-63) forward, and their
// whole backward, each as ONE persistent launch.
//
// With a batch of a few hundred utterances the head is a chain of GEMMs with 256 rows: < 1 GFLOP each, 1 % of the
// step's FLOPs, but as separate tcgen05 launches (prologue -> TMA pipeline -> epilogue -> teardown, 6-17 us apiece) they
// were 18 % of the training step.  Here one grid of co-resident CTAs (one per SM) walks a short list of STEPS separated
// by grid-wide barriers:
//   forward :  slab[s] = pooled[:, K-range s] . W1[K-range s, :]  (split-K over the CTAs, one fp32 slab per split)  |
//              H1 = bf16(relu(sum_s slab[s] + b1))  (fixed summation order: bit-reproducible)  |
//              H2 = bf16(relu(H1 . W2 + b2))
//   backward:  dH1 = (dH2 . W2^T) * (H1 > 0), db1 += colsum(dH1), dW2 += H1^T . dH2  |
//              gpool = dH1 . W1^T, dW1 += pooled^T . dH1
// Tiles are 64 x 128 x 64 per CTA on mma.sync.m16n8k16 (bf16 in, fp32 accumulate: a 256-row problem cannot fill
// tcgen05's 128-row tiles on 148 SMs).  Operands arrive by TMA tensor maps into 128B-swizzled shared-memory tiles in
// whatever orientation the buffers have (K-contiguous or M/N-contiguous: ldmatrix / ldmatrix.trans, conflict-free), so
// no transposed copy of weights or activations exists; stages are handed over with mbarriers (no bar.sync in the loop);
// every warp turns its 32 x 32 accumulator block around in shared memory so that results leave in 16-byte,
// row-contiguous accesses.  LBX_HEAD_PROFILE / LBX_HEAD_NO_MMA are measurement builds (tools/head_probe.py).
#include "tc_ptx.cuh"

namespace lbx {

typedef __nv_bfloat16 bf16;

#ifndef LBX_HT_STAGES
#define LBX_HT_STAGES 6
#endif
constexpr int HT_M = 64, HT_N = 128, HT_K = 64, HT_STAGES = LBX_HT_STAGES, HT_THREADS = 256;
// every operand tile is a stack of rows of 64 bf16 = 128 bytes in the TMA 128-byte swizzle (16-byte chunk c of row r
// is stored at chunk c ^ (r % 8)): ldmatrix reads 8 rows at one logical chunk = 8 different bank groups, conflict-free
// in both orientations
constexpr int HT_A_BYTES = HT_M * HT_K * 2;             // 8 KB: [64 rows][64] either way round
constexpr int HT_B_BYTES = HT_N * HT_K * 2;             // 16 KB: [128 n][64 k], or two [64 k][64 n] halves
constexpr int HT_STAGE_BYTES = HT_A_BYTES + HT_B_BYTES;
constexpr int HT_EPI_PITCH = 40;                         // floats per row of a warp's 32 x 32 epilogue tile
constexpr int HT_EPI_BYTES = (HT_THREADS / 32) * 32 * HT_EPI_PITCH * 4;
constexpr int HT_SMEM = HT_STAGES * HT_STAGE_BYTES + HT_EPI_BYTES + 1024;
static_assert(HT_STAGE_BYTES % 1024 == 0, "swizzled tiles need 1024-byte alignment");

enum { HEPI_ATOMIC = 0, HEPI_STORE_F32 = 1, HEPI_BIAS_ACT_BF16 = 2, HEPI_MASK_BF16_COLSUM = 3 };
enum { HSTEP_NONE = 0, HSTEP_GEMM = 1, HSTEP_FINALIZE = 2 };

struct HeadGemm {
  int map_a, map_b;        // indices into HeadChain::maps
  // a_trans = 0: A stored [M, K] (K contiguous); 1: stored [K, M] (M contiguous)
  // b_trans = 0: B stored [N, K] (K contiguous); 1: stored [K, N] (N contiguous)
  int M, N, K;
  int a_trans, b_trans;
  int k_splits, m_tiles, n_tiles;
  int epi;
  float* out_f32;
  bf16* out_bf16;
  long long ldo;
  long long split_stride;  // HEPI_STORE_F32 with k_splits > 1: split i stores its partial tile into slab i (deterministic)
  const float* bias;
  int relu;
  const bf16* mask;        // HEPI_MASK_BF16_COLSUM: keep x where mask[m, n] > 0 (pitch ldo)
  float* colsum;           // HEPI_MASK_BF16_COLSUM: colsum[n] += sum_m x[m, n]
};

struct HeadStep {
  int kind;
  int n_gemms;
  HeadGemm g[2];
  // HSTEP_FINALIZE: out[m, n] = bf16(act(sum_i z[i][m, n] + bias[n]))   (n_slabs dense [M, N] slabs, summed in order;
  // N % 4 == 0)
  float* z;
  int n_slabs;
  const float* bias;
  bf16* out;
  int M, N, relu;
  long long ldo;
};

struct HeadChain {
  CUtensorMap maps[8];     // 2-D bf16 maps, box = 64 elements (128 bytes, SWIZZLE_128B) x 64 or 128 rows
  HeadStep step[3];
  int n_steps;
  int n_maps;
  unsigned int* sync_ws;   // grid-barrier workspace, see grid_barrier()
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* smem_src) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_src);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* smem_src) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_src);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#ifdef LBX_HEAD_NO_MMA          // measurement only: everything but the tensor instruction
  c[0] += __uint_as_float(a[0] ^ b0); c[1] += __uint_as_float(a[1] ^ b1); c[2] += __uint_as_float(a[2]); c[3] += __uint_as_float(a[3]);
  return;
#endif
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned long long hd_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned int hd_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid-wide barrier number `b` (1-based count of all barriers ever executed on this workspace).  All CTAs of the grid
// are resident (grid <= number of SMs, one CTA per SM), so spinning cannot deadlock; the spin is bounded anyway (a GPU
// that hangs is worse than a flagged wrong answer): on time-out sync_ws[2] is set.
// Two levels, because 148 atomics on ONE address serialise at ~14 ns each in L2 (2 us per barrier): CTAs arrive on the
// counter of their group of 16 (one 128-byte line per group), the last arriver of a group arrives on the top-level
// counter, everybody polls the top-level counter.  Workspace layout (uint32): [0] top counter, [1] barriers completed
// by previous launches, [2] error flag, [32 + 32 * group] group counters.
constexpr int HB_GROUP = 16;
__device__ __forceinline__ void grid_barrier(unsigned int* sync_ws, unsigned int b) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int G = gridDim.x, grp = blockIdx.x / HB_GROUP, n_groups = (G + HB_GROUP - 1) / HB_GROUP;
    const unsigned int gs = min((unsigned int)HB_GROUP, G - grp * HB_GROUP);
    __threadfence();
    if (atomicAdd(sync_ws + 32 + 32 * grp, 1u) + 1u == b * gs) {
      __threadfence();
      atomicAdd(sync_ws, 1u);
    }
    const unsigned int target = b * n_groups;
    long long it = 0;
    while ((int)(hd_ld_acquire(sync_ws) - target) < 0) {
      if (++it > (1LL << 24)) { sync_ws[2] = 1u; break; }
      __nanosleep(20);
    }
    __threadfence();
  }
  __syncthreads();
}

// Operand pipeline: one elected thread issues TMA tile loads (2-D tensor maps: out-of-range rows / columns arrive as
// zeros, no predicates anywhere), "full" mbarriers count the bytes, "empty" mbarriers take one arrival per warp after
// its last ldmatrix of the stage.  (First version: per-thread 16-byte cp.async — measured 0.9-1.05 us per 24 KB k-chunk
// whatever the number of stages: an SM sustains only ~30 GB/s that way; TMA boxes do not have that limit.)
// The fill / use counters run on across tiles and steps, so the barriers are initialised once per kernel.
#ifdef LBX_HEAD_PROFILE
__device__ unsigned long long g_head_prof[8];
#endif
struct HeadPipe {
  uint64_t* full;
  uint64_t* empty;
  uint32_t fills, uses;
};

// One 64 x 128 output tile over the k-chunks [kc0, kc1) of HT_K.
template <int AT, int BT>
__device__ __forceinline__ void gemm_tile(const HeadGemm& g, const CUtensorMap* mapA, const CUtensorMap* mapB, int m_tile,
                                          int n_tile, int split, int kc0, int kc1, unsigned char* smem, HeadPipe& pipe) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 2, wn = warp & 3;                 // 2 x 4 warps, 32 x 32 outputs each
  const int m0 = m_tile * HT_M, n0 = n_tile * HT_N;
  const int M = g.M, N = g.N;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.0f;

  auto fill = [&](int kc) {                   // next free stage <- chunk kc (thread 0 only; counters advance everywhere)
    if (tid == 0) {
      const uint32_t stage = pipe.fills % HT_STAGES, round = pipe.fills / HT_STAGES;
      if (round > 0) mbar_wait(pipe.empty + stage, (round - 1) & 1);   // every warp has finished the previous use
      unsigned char* sA = smem + stage * HT_STAGE_BYTES;
      unsigned char* sB = sA + HT_A_BYTES;
      const int k0 = kc * HT_K;
      mbar_expect_tx(pipe.full + stage, (uint32_t)HT_STAGE_BYTES);
      if (AT == 0) tma_load_2d(mapA, pipe.full + stage, sA, k0, m0);          // [64 m rows][64 k]
      else tma_load_2d(mapA, pipe.full + stage, sA, m0, k0);                  // [64 k rows][64 m]
      if (BT == 0) {
        tma_load_2d(mapB, pipe.full + stage, sB, k0, n0);                     // [128 n rows][64 k]
      } else {
        tma_load_2d(mapB, pipe.full + stage, sB, n0, k0);                     // [64 k rows][64 n] x 2
        tma_load_2d(mapB, pipe.full + stage, sB + 8192, n0 + 64, k0);
      }
    }
    ++pipe.fills;
  };

  const int nk = kc1 - kc0;
#ifdef LBX_HEAD_PROFILE
  const long long t_e0 = clock64();
#endif
#pragma unroll 1
  for (int s = 0; s < HT_STAGES - 1; ++s)
    if (s < nk) fill(kc0 + s);
#ifdef LBX_HEAD_PROFILE
  const long long t_e1 = clock64();
#endif
  const int l7 = lane & 7, l8 = (lane >> 3) & 1, l16 = lane >> 4;
  // per-lane ldmatrix addressing inside a stage: byte offset of the lane's row and its logical 16-byte chunk at k16-step 0
  int a_row[2], a_chk[2], b_row[2], b_chk[2];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    const int ml = wm * 32 + mi * 16;
    a_row[mi] = AT == 0 ? (ml + l7 + l8 * 8) * 128 : (l7 + l16 * 8) * 128;
    a_chk[mi] = AT == 0 ? l16 : (ml >> 3) + l8;
  }
#pragma unroll
  for (int nj = 0; nj < 2; ++nj) {
    const int nl = wn * 32 + nj * 16;
    b_row[nj] = BT == 0 ? (nl + l7 + l16 * 8) * 128 : (wn >> 1) * 8192 + (l7 + l8 * 8) * 128;
    b_chk[nj] = BT == 0 ? l8 : ((nl & 63) >> 3) + l16;
  }
  for (int i = 0; i < nk; ++i) {
    if (i + HT_STAGES - 1 < nk) fill(kc0 + i + HT_STAGES - 1);
    const uint32_t stage = pipe.uses % HT_STAGES;
#ifdef LBX_HEAD_PROFILE
    const long long t_w0 = clock64();
#endif
    mbar_wait(pipe.full + stage, (pipe.uses / HT_STAGES) & 1);
#ifdef LBX_HEAD_PROFILE
    const long long t_w1 = clock64();
#endif
    const unsigned char* sA = smem + stage * HT_STAGE_BYTES;
    const unsigned char* sB = sA + HT_A_BYTES;
#pragma unroll
    for (int ks = 0; ks < HT_K / 16; ++ks) {
      uint32_t a[2][4], b[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        // K-contiguous tile: the k16 step moves 2 chunks along the row; M-contiguous: 16 rows down
        if (AT == 0) ldsm_x4(a[mi], sA + a_row[mi] + (((a_chk[mi] + 2 * ks) ^ l7) << 4));
        else ldsm_x4_trans(a[mi], sA + a_row[mi] + ks * 2048 + ((a_chk[mi] ^ l7) << 4));
      }
#pragma unroll
      for (int nj = 0; nj < 2; ++nj) {
        if (BT == 0) ldsm_x4(b[nj], sB + b_row[nj] + (((b_chk[nj] + 2 * ks) ^ l7) << 4));
        else ldsm_x4_trans(b[nj], sB + b_row[nj] + ks * 2048 + ((b_chk[nj] ^ l7) << 4));
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) mma_bf16_16816(acc[mi][ni], a[mi], b[ni >> 1][(ni & 1) * 2], b[ni >> 1][(ni & 1) * 2 + 1]);
    }
    // the MMAs above consumed every ldmatrix result of this stage: hand it back (one arrival per warp)
    __syncwarp();
    if (lane == 0) mbar_arrive(pipe.empty + stage);
    ++pipe.uses;
#ifdef LBX_HEAD_PROFILE
    if (blockIdx.x == 0 && tid == 32) {          // warp 1 lane 0: cycles waiting for data / computing, chunk count
      const long long t_c = clock64();
      atomicAdd(reinterpret_cast<unsigned long long*>(g_head_prof), (unsigned long long)(t_w1 - t_w0));
      atomicAdd(reinterpret_cast<unsigned long long*>(g_head_prof) + 1, (unsigned long long)(t_c - t_w1));
      atomicAdd(reinterpret_cast<unsigned long long*>(g_head_prof) + 2, 1ull);
    }
#endif
  }

#ifdef LBX_HEAD_PROFILE
  const long long t_e2 = clock64();
#endif
  // ---- epilogue.  In the accumulator layout a thread (g = lane / 4, t = lane % 4) holds rows g, g + 8 and columns 2t,
  // 2t + 1 of every 16 x 8 block: stored directly, one warp instruction would touch 8 rows x 32 bytes (measured: 1.8-2.3
  // us per tile, longer than the main loop).  Each warp therefore turns its 32 x 32 block around in a private
  // shared-memory tile (pitch 40 floats: both phases bank-conflict-free) and leaves with 16-byte accesses, a quarter
  // warp per 128-byte row segment.
  const int gq = lane >> 2, tq = lane & 3;
  const int epi = g.epi, relu = g.relu;
  const long long ldo = g.ldo;
  float* const out_f32 = g.out_f32 + split * g.split_stride;
  bf16* const out_bf16 = g.out_bf16;
  const float* const bias = g.bias;
  const bf16* const mask = g.mask;
  float* const colsum = g.colsum;
  float* st = reinterpret_cast<float*>(smem + HT_STAGES * HT_STAGE_BYTES) + warp * (32 * HT_EPI_PITCH);
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
        *reinterpret_cast<float2*>(st + (mi * 16 + gq + h * 8) * HT_EPI_PITCH + ni * 8 + 2 * tq) =
            make_float2(acc[mi][ni][2 * h], acc[mi][ni][2 * h + 1]);
  __syncwarp();
  const int rq = lane >> 3, cq = (lane & 7) * 4;              // this lane: rows rq + 4 it, columns cq .. cq + 3
  const int n = n0 + wn * 32 + cq;
  const bool n_ok = n < N;                                    // N is a multiple of 8: four columns are in range or not as a whole
  float4 bv = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (epi == HEPI_BIAS_ACT_BF16 && bias != nullptr && n_ok) bv = *reinterpret_cast<const float4*>(bias + n);
  float4 cs = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = rq + 4 * it;
    const int m = m0 + wm * 32 + r;
    if (m >= M || !n_ok) continue;
    float4 x = *reinterpret_cast<const float4*>(st + r * HT_EPI_PITCH + cq);
    const long long o = (long long)m * ldo + n;
    if (epi == HEPI_ATOMIC) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out_f32 + o), "f"(x.x), "f"(x.y), "f"(x.z),
                   "f"(x.w)
                   : "memory");
    } else if (epi == HEPI_STORE_F32) {
      *reinterpret_cast<float4*>(out_f32 + o) = x;
    } else {
      if (epi == HEPI_BIAS_ACT_BF16) {
        x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
        if (relu) { x.x = fmaxf(x.x, 0.0f); x.y = fmaxf(x.y, 0.0f); x.z = fmaxf(x.z, 0.0f); x.w = fmaxf(x.w, 0.0f); }
      } else {
        if (mask) {
          const uint2 w = *reinterpret_cast<const uint2*>(mask + o);
          if ((int)(w.x << 16) <= 0) x.x = 0.0f;              // bf16 > 0  <=>  sign clear and magnitude non-zero
          if ((int)(w.x & 0xFFFF0000u) <= 0) x.y = 0.0f;
          if ((int)(w.y << 16) <= 0) x.z = 0.0f;
          if ((int)(w.y & 0xFFFF0000u) <= 0) x.w = 0.0f;
        }
        cs.x += x.x; cs.y += x.y; cs.z += x.z; cs.w += x.w;
      }
      const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&lo);
      pk.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(out_bf16 + o) = pk;
    }
  }
  if (epi == HEPI_MASK_BF16_COLSUM && colsum != nullptr) {
    float v[4] = {cs.x, cs.y, cs.z, cs.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[e] += __shfl_xor_sync(0xffffffffu, v[e], 8);
      v[e] += __shfl_xor_sync(0xffffffffu, v[e], 16);
    }
    if (rq == 0 && n_ok) {
#pragma unroll
      for (int e = 0; e < 4; ++e) atomicAdd(colsum + n + e, v[e]);
    }
  }
  __syncwarp();                                               // the staging tile is rewritten by this warp's next tile
#ifdef LBX_HEAD_PROFILE
  if (blockIdx.x == 0 && tid == 32) {
    const long long t_e3 = clock64();
    unsigned long long* pr = reinterpret_cast<unsigned long long*>(g_head_prof);
    atomicAdd(pr + 4, (unsigned long long)(t_e1 - t_e0));
    atomicAdd(pr + 5, (unsigned long long)(t_e2 - t_e1));
    atomicAdd(pr + 6, (unsigned long long)(t_e3 - t_e2));
    atomicAdd(pr + 7, 1ull);
  }
#endif
}

__global__ void __launch_bounds__(HT_THREADS, 1) head_chain_kernel(const __grid_constant__ HeadChain P) {
  extern __shared__ __align__(128) unsigned char head_smem[];
  __shared__ uint64_t s_full[HT_STAGES], s_empty[HT_STAGES];
  __shared__ unsigned int s_base;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(head_smem) + 1023) & ~(uintptr_t)1023);
  if (threadIdx.x == 0) {
    for (int i = 0; i < P.n_maps; ++i) tma_prefetch_desc(&P.maps[i]);
    for (int i = 0; i < HT_STAGES; ++i) {
      mbar_init(s_full + i, 1);                  // one arrive.expect_tx by the issuing thread + the bytes
      mbar_init(s_empty + i, HT_THREADS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  HeadPipe pipe{s_full, s_empty, 0u, 0u};
  LBX_PDL_SYNC();
  if (P.n_steps > 1) {
    if (threadIdx.x == 0) s_base = *reinterpret_cast<volatile unsigned int*>(P.sync_ws + 1);
    __syncthreads();
  }
  // phase stamps of CTA 0 (globaltimer ns, 64-bit) at sync_ws[8 + 4 s]: step s work done, [10 + 4 s]: barrier passed;
  // sync_ws[4]: kernel start — measurement aid, costs three stores per step
  unsigned long long* stamps = reinterpret_cast<unsigned long long*>(P.sync_ws + 4);
  if (blockIdx.x == 0 && threadIdx.x == 0) stamps[0] = hd_timer_ns();
  for (int s = 0; s < P.n_steps; ++s) {
    const HeadStep& st = P.step[s];
    if (st.kind == HSTEP_GEMM) {
      int items[2] = {0, 0};
      for (int j = 0; j < st.n_gemms; ++j) items[j] = st.g[j].m_tiles * st.g[j].n_tiles * st.g[j].k_splits;
      const int total = items[0] + items[1];
      for (int it = blockIdx.x; it < total; it += gridDim.x) {
        const int j = it < items[0] ? 0 : 1;
        const HeadGemm& g = st.g[j];
        const int r = j == 0 ? it : it - items[0];
        const int tiles = g.m_tiles * g.n_tiles;
        const int split = r / tiles, tile = r - split * tiles;
        const int m_tile = tile / g.n_tiles, n_tile = tile - m_tile * g.n_tiles;
        const int nkc = (g.K + HT_K - 1) / HT_K;
        const int per = (nkc + g.k_splits - 1) / g.k_splits;
        const int kc0 = split * per, kc1 = min(nkc, kc0 + per);
        if (kc0 >= kc1) continue;               // host: every split is non-empty
        if (g.a_trans == 0 && g.b_trans == 1) gemm_tile<0, 1>(g, &P.maps[g.map_a], &P.maps[g.map_b], m_tile, n_tile, split, kc0, kc1, smem, pipe);
        else if (g.a_trans == 0 && g.b_trans == 0) gemm_tile<0, 0>(g, &P.maps[g.map_a], &P.maps[g.map_b], m_tile, n_tile, split, kc0, kc1, smem, pipe);
        else if (g.a_trans == 1 && g.b_trans == 1) gemm_tile<1, 1>(g, &P.maps[g.map_a], &P.maps[g.map_b], m_tile, n_tile, split, kc0, kc1, smem, pipe);
        else gemm_tile<1, 0>(g, &P.maps[g.map_a], &P.maps[g.map_b], m_tile, n_tile, split, kc0, kc1, smem, pipe);
      }
    } else if (st.kind == HSTEP_FINALIZE) {
      const long long n4 = (long long)st.M * st.N / 4;
      for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 z = reinterpret_cast<const float4*>(st.z)[i];
        for (int sl = 1; sl < st.n_slabs; ++sl) {
          const float4 zs = reinterpret_cast<const float4*>(st.z)[(long long)sl * n4 + i];
          z.x += zs.x; z.y += zs.y; z.z += zs.z; z.w += zs.w;
        }
        const long long e = 4 * i;
        const int m = (int)(e / st.N), n = (int)(e - (long long)m * st.N);
        if (st.bias) {
          const float4 b = *reinterpret_cast<const float4*>(st.bias + n);
          z.x += b.x; z.y += b.y; z.z += b.z; z.w += b.w;
        }
        if (st.relu) {
          z.x = fmaxf(z.x, 0.0f); z.y = fmaxf(z.y, 0.0f); z.z = fmaxf(z.z, 0.0f); z.w = fmaxf(z.w, 0.0f);
        }
        const __nv_bfloat162 lo = __floats2bfloat162_rn(z.x, z.y), hi = __floats2bfloat162_rn(z.z, z.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo);
        pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(st.out + (long long)m * st.ldo + n) = pk;
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) stamps[2 + 2 * s] = hd_timer_ns();
    if (s + 1 < P.n_steps) grid_barrier(P.sync_ws, s_base + (unsigned int)(s + 1));
    if (blockIdx.x == 0 && threadIdx.x == 0) stamps[3 + 2 * s] = hd_timer_ns();
  }
  // the barrier count advances once per launch; every CTA has read it (they all passed the first barrier)
  if (P.n_steps > 1 && blockIdx.x == 0 && threadIdx.x == 0) P.sync_ws[1] = s_base + (unsigned int)(P.n_steps - 1);
}

#ifdef LBX_HEAD_PROFILE
extern "C" int lbx_head_profile(unsigned long long* out, int reset) {
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpyFromSymbol(out, g_head_prof, sizeof(z));
  if (reset) cudaMemcpyToSymbol(g_head_prof, z, sizeof(z));
  return 0;
}
#endif
static int g_head_grid = 0;

static int head_init() {
  if (g_head_grid == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    LBX_CUDA(cudaGetDevice(&dev));
    LBX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    LBX_CUDA(cudaFuncSetAttribute(head_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_SMEM));
    LBX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, head_chain_kernel, HT_THREADS, HT_SMEM));
    if (per_sm < 1) return set_error(LBX_ECUDA, "head kernel does not fit an SM");
    g_head_grid = sms;                            // one CTA per SM: all resident at once (grid-wide barriers)
  }
  return LBX_OK;
}

static int head_launch(HeadChain& P, cudaStream_t stream) {
  LBX_LAUNCH_PDL(head_chain_kernel, dim3((unsigned)g_head_grid), dim3(HT_THREADS), (size_t)HT_SMEM, stream, P);
  return LBX_OK;
}

static void set_tiles(HeadGemm& g, int k_splits) {
  g.m_tiles = (g.M + HT_M - 1) / HT_M;
  g.n_tiles = (g.N + HT_N - 1) / HT_N;
  const int nkc = (g.K + HT_K - 1) / HT_K;
  int ks = k_splits < 1 ? 1 : (k_splits > nkc ? nkc : k_splits);
  const int per = (nkc + ks - 1) / ks;
  g.k_splits = (nkc + per - 1) / per;            // no empty split
}

// tensor maps of one GEMM's operands (see gemm_tile for the box shapes)
static int set_operands(HeadChain& P, int& n_maps, HeadGemm& g, const void* A, long long lda, int a_trans, const void* B,
                        long long ldb, int b_trans) {
  g.a_trans = a_trans; g.b_trans = b_trans;
  g.map_a = n_maps++; g.map_b = n_maps++;
  int rc;
  if (a_trans == 0) rc = make_map(&P.maps[g.map_a], A, g.M, g.K, lda, 64, 64);
  else rc = make_map(&P.maps[g.map_a], A, g.K, g.M, lda, 64, 64);
  if (rc) return rc;
  if (b_trans == 0) rc = make_map(&P.maps[g.map_b], B, g.N, g.K, ldb, 64, 128);
  else rc = make_map(&P.maps[g.map_b], B, g.K, g.N, ldb, 64, 64);
  return rc;
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace lbx

using namespace lbx;

extern "C" int lbx_head_fwd(const void* pooled_bf16, long long B, int K1, const void* w1_bf16, int ldw1, const float* b1,
                            int N1, const void* w2_bf16, int ldw2, const float* b2, int N2, void* h1_bf16,
                            void* h2_bf16, float* scratch, long long scratch_floats, unsigned int* sync_ws,
                            void* stream) {
  LBX_CHECK_ARG(B >= 0 && B <= 1 << 20 && K1 >= 8 && N1 >= 8 && N2 >= 8, "bad head shape");
  LBX_CHECK_ARG(K1 % 8 == 0 && N1 % 8 == 0 && N2 % 8 == 0 && ldw1 % 8 == 0 && ldw2 % 8 == 0,
                "layer widths and pitches must be multiples of 8");
  if (B == 0) return LBX_OK;
  LBX_CHECK_ARG(pooled_bf16 && w1_bf16 && w2_bf16 && h1_bf16 && h2_bf16 && scratch && sync_ws, "NULL pointer argument");
  LBX_CHECK_ARG(al16(pooled_bf16) && al16(w1_bf16) && al16(w2_bf16) && al16(h1_bf16) && al16(h2_bf16) && al16(scratch) &&
                    (b1 == nullptr || al16(b1)) && (b2 == nullptr || al16(b2)),
                "buffers must be 16-byte aligned");
  int rc = head_init();
  if (rc) return rc;
  HeadChain P{};
  P.sync_ws = sync_ws;
  P.n_steps = 3;
  // step 0: Z1 += pooled . W1, split-K so that every SM has a tile
  HeadStep& s0 = P.step[0];
  s0.kind = HSTEP_GEMM; s0.n_gemms = 1;
  HeadGemm& g0 = s0.g[0];
  int n_maps = 0;
  g0.M = (int)B; g0.N = N1; g0.K = K1;
  if ((rc = set_operands(P, n_maps, g0, pooled_bf16, K1, 0, w1_bf16, ldw1, 1))) return rc;
  // every split stores its partial sums into its own slab of the scratch buffer; the finalize step adds the slabs in
  // a fixed order: the forward pass is bit-reproducible (no floating-point atomics)
  g0.epi = HEPI_STORE_F32; g0.out_f32 = scratch; g0.ldo = N1; g0.split_stride = B * N1;
  set_tiles(g0, 1);
  {
    const int tiles = g0.m_tiles * g0.n_tiles;
    long long want = tiles >= g_head_grid ? 1 : g_head_grid / tiles;
    const long long room = scratch_floats / (B * N1);
    LBX_CHECK_ARG(room >= 1, "scratch must hold at least B * N1 floats");
    set_tiles(g0, (int)(want < room ? want : room));
  }
  // step 1: H1 = bf16(relu(sum of the slabs + b1))
  HeadStep& s1 = P.step[1];
  s1.kind = HSTEP_FINALIZE; s1.z = scratch; s1.n_slabs = g0.k_splits; s1.bias = b1; s1.out = (bf16*)h1_bf16;
  s1.M = (int)B; s1.N = N1; s1.relu = 1;
  s1.ldo = N1;
  // step 2: H2 = bf16(relu(H1 . W2 + b2))
  HeadStep& s2 = P.step[2];
  s2.kind = HSTEP_GEMM; s2.n_gemms = 1;
  HeadGemm& g2 = s2.g[0];
  g2.M = (int)B; g2.N = N2; g2.K = N1;
  if ((rc = set_operands(P, n_maps, g2, h1_bf16, N1, 0, w2_bf16, ldw2, 1))) return rc;
  g2.epi = HEPI_BIAS_ACT_BF16; g2.out_bf16 = (bf16*)h2_bf16; g2.ldo = N2; g2.bias = b2; g2.relu = 1;
  set_tiles(g2, 1);
  P.n_maps = n_maps;
  return head_launch(P, (cudaStream_t)stream);
}

extern "C" int lbx_head_bwd(const void* dh2_bf16, const void* pooled_bf16, const void* h1_bf16, long long B, int K1,
                            int N1, int N2, const void* w1_bf16, int ldw1, const void* w2_bf16, int ldw2,
                            void* dh1_bf16, float* gpool, float* dw1, float* db1, float* dw2, unsigned int* sync_ws,
                            void* stream) {
  LBX_CHECK_ARG(B >= 0 && B <= 1 << 20 && K1 >= 8 && N1 >= 8 && N2 >= 8, "bad head shape");
  LBX_CHECK_ARG(K1 % 8 == 0 && N1 % 8 == 0 && N2 % 8 == 0 && ldw1 % 8 == 0 && ldw2 % 8 == 0,
                "layer widths and pitches must be multiples of 8");
  if (B == 0) return LBX_OK;
  LBX_CHECK_ARG(dh2_bf16 && pooled_bf16 && h1_bf16 && w1_bf16 && w2_bf16 && dh1_bf16 && gpool && dw1 && dw2 && sync_ws,
                "NULL pointer argument");
  LBX_CHECK_ARG(al16(dh2_bf16) && al16(pooled_bf16) && al16(h1_bf16) && al16(w1_bf16) && al16(w2_bf16) && al16(dh1_bf16) &&
                    al16(gpool) && al16(dw1) && al16(dw2),
                "buffers must be 16-byte aligned");
  int rc = head_init();
  if (rc) return rc;
  HeadChain P{};
  P.sync_ws = sync_ws;
  P.n_steps = 2;
  // step 0: dH1 = (dH2 . W2^T) * (H1 > 0) (+ db1); dW2 += H1^T . dH2
  HeadStep& s0 = P.step[0];
  s0.kind = HSTEP_GEMM; s0.n_gemms = 2;
  HeadGemm& a = s0.g[0];
  int n_maps = 0;
  a.M = (int)B; a.N = N1; a.K = N2;
  if ((rc = set_operands(P, n_maps, a, dh2_bf16, N2, 0, w2_bf16, ldw2, 0))) return rc;   // B[n = input unit][k = output unit] = W2 as stored
  a.epi = HEPI_MASK_BF16_COLSUM; a.out_bf16 = (bf16*)dh1_bf16; a.ldo = N1; a.mask = (const bf16*)h1_bf16; a.colsum = db1;
  set_tiles(a, 1);
  HeadGemm& b = s0.g[1];
  b.M = N1; b.N = N2; b.K = (int)B;
  if ((rc = set_operands(P, n_maps, b, h1_bf16, N1, 1, dh2_bf16, N2, 1))) return rc;     // A^T stored [batch, N1]
  b.epi = HEPI_ATOMIC; b.out_f32 = dw2; b.ldo = ldw2;
  set_tiles(b, 4);
  // step 1: gpool = dH1 . W1^T; dW1 += pooled^T . dH1
  HeadStep& s1 = P.step[1];
  s1.kind = HSTEP_GEMM; s1.n_gemms = 2;
  HeadGemm& c = s1.g[0];
  c.M = (int)B; c.N = K1; c.K = N1;
  if ((rc = set_operands(P, n_maps, c, dh1_bf16, N1, 0, w1_bf16, ldw1, 0))) return rc;
  c.epi = HEPI_STORE_F32; c.out_f32 = gpool; c.ldo = K1;
  set_tiles(c, 1);
  HeadGemm& d = s1.g[1];
  d.M = K1; d.N = N1; d.K = (int)B;
  if ((rc = set_operands(P, n_maps, d, pooled_bf16, K1, 1, dh1_bf16, N1, 1))) return rc;
  d.epi = HEPI_ATOMIC; d.out_f32 = dw1; d.ldo = ldw1;
  set_tiles(d, 1);
  P.n_maps = n_maps;
  return head_launch(P, (cudaStream_t)stream);
}
