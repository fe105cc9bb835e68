// Grouped weight-gradient GEMM (sm_100a): every dW_p += X_p^T . dZ_p of one backward pass in ONE persistent launch.
//
// The weight gradients of the TDNN frame layers (lidbox/models/xvector.py:38-43, 53-57 under Keras `fit`) contract over
// the batch*time rows — tens of thousands — into outputs of a few hundred rows: each is a split-K problem, and as
// separate launches each pays its own prologue / pipeline fill / exposed atomic epilogue / tail (5 launches, 116 us for
// 52 us of tensor work at batch 256 x 2 s).  Here the (problem, k-chunk, output tile, k-block) space of ALL problems is
// laid out on one line and cut into equal contiguous ranges, one per CTA pair ("stream-K"): every pair runs the same
// number of k-blocks back to back, crossing tile and problem boundaries without draining the TMA pipeline, and the
// fp32 `red.global.add.v4` epilogue of a finished segment overlaps the main loop of the next one (two TMEM accumulator
// stages).  Because the outputs are accumulated atomically into the (zeroed) flat gradient buffer, partial tiles need
// no fix-up pass.
//
// Line order inside a problem: k-chunk (length ~ one pair's share) -> output tile -> k-block, so neighbouring pairs work
// on the same rows of the activations at the same time (each activation row is fetched from HBM once and served to the
// other tiles from L2), exactly like the split-major tile order of gemm_bf16_kernel.
//
// Operands are read in place from the activation / gradient buffers as MN-major tcgen05 operands (layout TN of
// gemm.cu): A = X [rows, k*C_in] through the overlapping-row view (pitch stride*C_in), B = dZ [rows, C_out].
// CTA pairs (tcgen05 cta_group::2, M = 256): roles as in gemm.cu — warp 0 TMA producer, warp 1 MMA issuer (leader CTA),
// warps 4..11 epilogue.
#include "tc_ptx.cuh"

namespace lbx {

constexpr int GW_MAX_PROBLEMS = 8;
constexpr int GW_BN = 256;
constexpr int GW_STAGES = 6;
constexpr int GW_A_BYTES = BM * BK * 2;                 // 16 KB: this CTA's 128 output rows x 64 contraction rows
constexpr int GW_B_BYTES = (GW_BN / 2) * BK * 2;        // 16 KB: this CTA's half of the 256 output columns
constexpr int GW_STAGE_BYTES = GW_A_BYTES + GW_B_BYTES;
constexpr int GW_EPI_WARPS = 8;
constexpr int GW_CTRL_WARPS = 4;
constexpr int GW_THREADS = 32 * (GW_CTRL_WARPS + GW_EPI_WARPS);
constexpr int GW_TMEM_COLS = 2 * GW_BN;
constexpr size_t GW_SMEM = (size_t)GW_STAGES * GW_STAGE_BYTES + 1024 + 256;

struct GwProblem {
  int M, N;                // output rows (= k*C_in), output columns (= C_out)
  int kb_total;            // 64-row blocks of the contraction
  int chunk;               // k-blocks per k-chunk
  int m_units, n_tiles;    // 256-row units (two 128-row tiles of a CTA pair), 256-column tiles
  long long line_start;    // first position of this problem on the work line
  long long ldo;           // output pitch (floats)
  float* out;
};

struct GwParams {
  CUtensorMap mapA[GW_MAX_PROBLEMS];
  CUtensorMap mapB[GW_MAX_PROBLEMS];
  GwProblem prob[GW_MAX_PROBLEMS];
  int n_problems;
  long long line_total;
};

struct GwSegment {
  int p, m_unit, n_blk, kb0, kb1;
};

// Decodes line position x (x < x_end) into the segment that starts there and ends at the end of its (tile, k-chunk) or
// at x_end, whichever comes first; returns the position after the segment.
// (n_tiles counts 256-column tiles, or PAIRS of them when two CTA pairs of a cluster share the A tile)
__host__ __device__ __forceinline__ long long gw_decode(const GwParams& P, long long x, long long x_end, GwSegment& s) {
  int p = 0;
  while (p + 1 < P.n_problems && x >= P.prob[p + 1].line_start) ++p;
  const GwProblem& q = P.prob[p];
  const long long units = (long long)q.m_units * q.n_tiles;
  const long long r = x - q.line_start;
  const long long per_chunk = units * q.chunk;
  const int c = (int)(r / per_chunk);
  const long long r2 = r - (long long)c * per_chunk;
  const int len = q.chunk < q.kb_total - c * q.chunk ? q.chunk : q.kb_total - c * q.chunk;
  const int u = (int)(r2 / len);
  const int koff = (int)(r2 - (long long)u * len);
  long long n = len - koff;
  if (n > x_end - x) n = x_end - x;
  s.p = p;
  s.m_unit = u / q.n_tiles;
  s.n_blk = u - s.m_unit * q.n_tiles;
  s.kb0 = c * q.chunk + koff;
  s.kb1 = s.kb0 + (int)n;
  return x + n;
}

// TMA load multicast to the CTAs of `mask`; the completion bytes are counted on the pair leader of every destination
// (the mbarrier operand names the leader's barrier of the issuing CTA's pair: its CTA-relative offset and its "peer
// bit" select the barrier in each destination pair)
__device__ __forceinline__ void tma_load_2d_pair_mc(const CUtensorMap* map, uint32_t mbar_cluster_addr, void* dst, int c0,
                                                    int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mask(uint64_t* bar, uint16_t mask) {   // arrives at this offset in the CTAs of mask
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// QUAD: clusters of 4 = two CTA pairs that work on the SAME rows of A (same 256 output rows) and on neighbouring
// 256-column tiles of B.  Every CTA loads half of its A tile and multicasts it to the CTA of the other pair that needs
// the same rows, so a pair ingests 48 KB instead of 64 KB per k-block: the launch is bound by the L2 -> SM fabric.
template <bool QUAD>
__global__ void __launch_bounds__(GW_THREADS, 1) wgrad_grouped_kernel(const __grid_constant__ GwParams P) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)GW_STAGES * GW_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + GW_STAGES;
  uint64_t* tfull_bar = empty_bar + GW_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  static_assert((2 * GW_STAGES + 4) * 8 + 4 <= 256, "barrier area");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cluster_rank = (int)cluster_ctarank();            // 0..1, QUAD: 0..3
  const int pair = QUAD ? cluster_rank >> 1 : 0;              // which CTA pair of the cluster
  const int cta_rank = cluster_rank & 1;                      // rank inside the pair (0 = leader, issues the MMAs)
  const int lead_rank = cluster_rank & ~1;                    // cluster rank of this pair's leader
  constexpr int CSZ = QUAD ? 4 : 2;
  const long long worker = blockIdx.x / CSZ, n_workers = gridDim.x / CSZ;
  const long long x_begin = P.line_total * worker / n_workers, x_end = P.line_total * (worker + 1) / n_workers;

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.n_problems; ++i) {
      tma_prefetch_desc(&P.mapA[i]);
      tma_prefetch_desc(&P.mapB[i]);
    }
    for (int i = 0; i < GW_STAGES; ++i) {
      mbar_init(full_bar + i, 2);               // the producers of both CTAs arrive on the leader's barrier
      mbar_init(empty_bar + i, QUAD ? 2 : 1);   // QUAD: a stage is refilled from BOTH pairs: both MMA issuers release it
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar + i, 1);
      mbar_init(tempty_bar + i, 2 * GW_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(GW_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  LBX_PDL_SYNC();

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long x = x_begin; x < x_end;) {
        GwSegment s;
        x = gw_decode(P, x, x_end, s);
        const CUtensorMap* mA = &P.mapA[s.p];
        const CUtensorMap* mB = &P.mapB[s.p];
        const int m0 = (2 * s.m_unit + cta_rank) * BM;
        const int n_blk = QUAD ? 2 * s.n_blk + pair : s.n_blk;
        const int n0 = n_blk * GW_BN + cta_rank * (GW_BN / 2);
        for (int kb = s.kb0; kb < s.kb1; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          unsigned char* sA = smem + (size_t)stage * GW_STAGE_BYTES;
          unsigned char* sB = sA + GW_A_BYTES;
          const uint32_t lead_full = mapa_u32(smem_u32(full_bar + stage), lead_rank);
          if (cta_rank == 0) mbar_expect_tx(full_bar + stage, 2u * (uint32_t)GW_STAGE_BYTES);
          else mbar_arrive_remote(lead_full);
          if (QUAD) {
            // this CTA fetches 64 of its 128 A columns (box `pair`) for itself AND for the CTA with the same rank in the
            // other pair; the other 64 arrive from there
            tma_load_2d_pair_mc(mA, lead_full, sA + pair * 8192, m0 + pair * 64, kb * BK,
                                (uint16_t)((1u << cta_rank) | (1u << (cta_rank + 2))));
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d_pair(mA, lead_full, sA + i * 8192, m0 + i * 64, kb * BK);
          }
#pragma unroll
          for (int i = 0; i < GW_BN / 2 / 64; ++i) tma_load_2d_pair(mB, lead_full, sB + i * 8192, n0 + i * 64, kb * BK);
          if (++stage == GW_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA) =====================================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc<GW_BN, 256>(1, 1);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (long long x = x_begin; x < x_end;) {
        GwSegment s;
        x = gw_decode(P, x, x_end, s);
        const int iters = s.kb1 - s.kb0;
        mbar_wait(tempty_bar + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * GW_BN;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + (size_t)stage * GW_STAGE_BYTES);
          const uint32_t sB = sA + GW_A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // MN-major operands: 64-element blocks are 8192 B apart (LBO), 8 k-rows = 1024 B (SBO), 2048 B per UMMA_K
            const uint64_t da = make_smem_desc(sA + k * (UMMA_K * 128), 8192, 1024);
            const uint64_t db = make_smem_desc(sB + k * (UMMA_K * 128), 8192, 1024);
            tc_mma_bf16_pair(tmem_d, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          if (QUAD) tc_commit_mask(empty_bar + stage, (uint16_t)0xF);   // the stage is free in all four CTAs (for this pair)
          else tc_commit_pair(empty_bar + stage);
          if (++stage == GW_STAGES) { stage = 0; phase ^= 1; }
        }
        if (QUAD) tc_commit_mask(tfull_bar + acc, (uint16_t)(3u << (2 * pair)));
        else tc_commit_pair(tfull_bar + acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= GW_CTRL_WARPS) {
    // ===================================== epilogue: fp32 vector reductions into the gradient buffer ==========
    const int sub = warp & 3;                                 // TMEM sub-partition: lanes [32*sub, 32*sub + 32)
    const int chalf = (warp - GW_CTRL_WARPS) >> 2;            // which 128 of the tile's 256 columns
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long x = x_begin; x < x_end;) {
      GwSegment s;
      x = gw_decode(P, x, x_end, s);
      const GwProblem& q = P.prob[s.p];
      const int m = (2 * s.m_unit + cta_rank) * BM + sub * 32 + lane;
      const int col_base = (QUAD ? 2 * s.n_blk + pair : s.n_blk) * GW_BN + chalf * (GW_BN / 2);
      const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * GW_BN + chalf * (GW_BN / 2));
      float* orow = q.out + (long long)m * q.ldo;
      const bool row_ok = m < q.M;
      const bool al16 = ((reinterpret_cast<uintptr_t>(q.out) & 15) == 0) && ((q.ldo & 3) == 0);
      mbar_wait(tfull_bar + acc, acc_phase);
      tc_fence_after();
      uint32_t v[2][32];
      tc_ld32(taddr, v[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tc_wait_ld();
        if (c + 1 < 4) {
          tc_ld32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
        } else {
          tc_fence_before();                                  // all TMEM reads of this warp have landed
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(tempty_bar + acc), lead_rank));
        }
        const int n0 = col_base + c * 32;
        if (row_ok && n0 < q.N) {
          float* o = orow + n0;
          const uint32_t* w = v[c & 1];
          if (al16 && n0 + 32 <= q.N) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * g),
                           "f"(__uint_as_float(w[4 * g])), "f"(__uint_as_float(w[4 * g + 1])),
                           "f"(__uint_as_float(w[4 * g + 2])), "f"(__uint_as_float(w[4 * g + 3]))
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < q.N) atomicAdd(o + j, __uint_as_float(w[j]));
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(GW_TMEM_COLS) : "memory");
  }
}

}  // namespace lbx

using namespace lbx;

static int g_wgrad_quad = 0;

// max co-resident clusters of `csz` CTAs of kernel `k`
template <typename K>
static int gw_max_clusters(K k, int csz, int sms) {
  cudaLaunchConfig_t occ{};
  occ.gridDim = dim3((unsigned)(sms / csz * csz));
  occ.blockDim = dim3(GW_THREADS);
  occ.dynamicSmemBytes = GW_SMEM;
  cudaLaunchAttribute oa[1];
  oa[0].id = cudaLaunchAttributeClusterDimension;
  oa[0].val.clusterDim.x = (unsigned)csz; oa[0].val.clusterDim.y = 1; oa[0].val.clusterDim.z = 1;
  occ.attrs = oa; occ.numAttrs = 1;
  int clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&clusters, k, &occ) != cudaSuccess || clusters < 1) {
    (void)cudaGetLastError();
    clusters = sms / csz;
  }
  return clusters < sms / csz ? clusters : sms / csz;
}

// Lays the problems on the work line (tile counts, k-chunk lengths, line offsets) and, when with_maps, encodes the
// tensor maps.  The k-chunk length is the worker's share of the line, so that neighbouring workers run the same
// k-blocks of neighbouring tiles at the same time.
static int gw_plan(const lbx_wgrad_t* problems, int n, int workers_max, bool quad, GwParams& P, bool with_maps) {
  long long total = 0;
  int np = 0;
  for (int i = 0; i < n; ++i) {
    const lbx_wgrad_t& g = problems[i];
    LBX_CHECK_ARG(g.rows >= 0 && g.rows <= 2147483647LL && g.a_cols >= 0 && g.b_cols >= 0, "bad extent in problem %d", i);
    if (with_maps) {
      LBX_CHECK_ARG(g.a && g.b && g.out, "NULL operand in problem %d", i);
      LBX_CHECK_ARG(g.lda % 8 == 0 && g.ldb % 8 == 0, "operand pitches must be multiples of 8 elements (16 bytes)");
      LBX_CHECK_ARG((reinterpret_cast<uintptr_t>(g.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.b) & 15) == 0,
                    "operands must be 16-byte aligned");
    }
    if (g.rows == 0 || g.a_cols == 0 || g.b_cols == 0) continue;
    GwProblem& q = P.prob[np];
    q.M = g.a_cols; q.N = g.b_cols;
    q.kb_total = (int)((g.rows + BK - 1) / BK);
    q.m_units = ((q.M + BM - 1) / BM + 1) / 2;
    q.n_tiles = (q.N + GW_BN - 1) / GW_BN;
    if (quad) q.n_tiles /= 2;                        // pairs of tiles: one per CTA pair of the cluster
    q.ldo = g.ldo; q.out = g.out;
    if (with_maps) {
      int rc;
      if ((rc = make_map(&P.mapA[np], g.a, g.rows, g.a_cols, g.lda, 64, 64))) return rc;
      if ((rc = make_map(&P.mapB[np], g.b, g.rows, g.b_cols, g.ldb, 64, 64))) return rc;
    }
    total += (long long)q.m_units * q.n_tiles * q.kb_total;
    ++np;
  }
  P.n_problems = np;
  P.line_total = total;
  if (np == 0) return LBX_OK;
  const long long share = (total + workers_max - 1) / workers_max;
  long long pos = 0;
  for (int i = 0; i < np; ++i) {
    GwProblem& q = P.prob[i];
    long long splits = (q.kb_total + share / 2) / (share > 0 ? share : 1);      // round(kb_total / share)
    if (splits < 1) splits = 1;
    if (splits > q.kb_total) splits = q.kb_total;
    q.chunk = (int)((q.kb_total + splits - 1) / splits);
    q.line_start = pos;
    pos += (long long)q.m_units * q.n_tiles * q.kb_total;
  }
  return LBX_OK;
}

extern "C" int lbx_wgrad_grouped_plan(const lbx_wgrad_t* problems, int n, int workers, int quad, int* segments,
                                      int max_segments, int* n_segments) {
  LBX_CHECK_ARG(problems != nullptr && n >= 1 && n <= GW_MAX_PROBLEMS && workers >= 1 && segments && n_segments,
                "bad arguments");
  GwParams P{};
  int rc = gw_plan(problems, n, workers, quad != 0, P, false);
  if (rc) return rc;
  const long long total = P.line_total;
  const long long nw = total < workers ? total : workers;
  int count = 0;
  for (long long w = 0; w < nw; ++w) {
    const long long x_end = total * (w + 1) / nw;
    for (long long x = total * w / nw; x < x_end;) {
      GwSegment sg;
      x = gw_decode(P, x, x_end, sg);
      LBX_CHECK_ARG(count < max_segments, "segment buffer too small");
      int* o = segments + 6 * count++;
      o[0] = (int)w; o[1] = sg.p; o[2] = sg.m_unit; o[3] = sg.n_blk; o[4] = sg.kb0; o[5] = sg.kb1;
    }
  }
  *n_segments = count;
  return LBX_OK;
}

extern "C" int lbx_set_wgrad_quad(int enabled) {
  g_wgrad_quad = enabled ? 1 : 0;
  return LBX_OK;
}

extern "C" int lbx_wgrad_grouped(const lbx_wgrad_t* problems, int n, void* stream) {
  LBX_CHECK_ARG(problems != nullptr && n >= 1 && n <= GW_MAX_PROBLEMS, "1..%d problems per launch", GW_MAX_PROBLEMS);
  static int max_pairs = 0, max_quads = 0;
  if (max_pairs == 0) {
    int dev = 0, sms = 0;
    LBX_CUDA(cudaGetDevice(&dev));
    LBX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    LBX_CUDA(cudaFuncSetAttribute(wgrad_grouped_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GW_SMEM));
    LBX_CUDA(cudaFuncSetAttribute(wgrad_grouped_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GW_SMEM));
    // the line is cut into one range per cluster: every cluster must be resident at once, or the last ones would run alone
    max_pairs = gw_max_clusters(wgrad_grouped_kernel<false>, 2, sms);
    max_quads = gw_max_clusters(wgrad_grouped_kernel<true>, 4, sms);
  }
  // QUAD (clusters of two pairs sharing the A tile): every problem must have an even number of 256-column tiles
  bool quad = g_wgrad_quad != 0 && max_quads >= 1;
  for (int i = 0; i < n && quad; ++i)
    if (problems[i].rows > 0 && problems[i].a_cols > 0 && problems[i].b_cols > 0 &&
        ((problems[i].b_cols + GW_BN - 1) / GW_BN) % 2 != 0)
      quad = false;
  const int workers_max = quad ? max_quads : max_pairs;
  GwParams P{};
  int rc = gw_plan(problems, n, workers_max, quad, P, true);
  if (rc) return rc;
  const int np = P.n_problems;
  const long long total = P.line_total;
  if (np == 0) return LBX_OK;
  const int workers = (int)(total < workers_max ? total : workers_max);
  const int csz = quad ? 4 : 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(csz * workers));
  cfg.blockDim = dim3(GW_THREADS);
  cfg.dynamicSmemBytes = GW_SMEM;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g_use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = (unsigned)csz;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t le = quad ? cudaLaunchKernelEx(&cfg, wgrad_grouped_kernel<true>, P)
                        : cudaLaunchKernelEx(&cfg, wgrad_grouped_kernel<false>, P);
  if (le != cudaSuccess) return set_error(LBX_ECUDA, "grouped weight-gradient launch failed: %s", cudaGetErrorString(le));
  LBX_LAUNCH_CHECK();
  return LBX_OK;
}
