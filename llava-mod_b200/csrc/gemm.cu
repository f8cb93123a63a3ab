// gemm.cu -- hand-written tcgen05 + TMA GEMM for sm_100a:  D[M,N] (+)= A[M,K] * B[N,K]^T  (bf16 in, fp32 TMEM accumulate)
//
// Replaces the nn.Linear call sites of the path (modeling_qwen2.py:199-200 gate/up/down, :678-680 q/k/v, :726 o_proj, :1176
// lm_head; CLIP / projector linears) and, in its grouped form, DeepSpeed's per-expert loop (Experts.forward) on COMPACT
// expert rows -- no capacity padding.
//
// Structure (one persistent CTA per SM, 256 threads, warp-specialised):
//   warp 0 lane 0 : TMA producer  -- cp.async.bulk.tensor.2d into a 4-stage 128B-swizzled shared-memory ring (A 128x64, B 256x64)
//   warp 1 lane 0 : MMA issuer    -- tcgen05.mma.cta_group::1.kind::f16, M=128 N=256 K=16, accumulators in TMEM (2 x 256 columns,
//                                    double buffered so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warp 2        : TMEM allocator
//   warps 4..7    : epilogue      -- tcgen05.ld 32x32b.x32 -> (+bias) (+D_old) -> bf16 -> 64-byte row segments to global
//   mbarriers     : full[stage] (TMA complete_tx) / empty[stage] (tcgen05.commit) / tmem_full[2] / tmem_empty[2]
// Operand "major-ness" is a template parameter, so dgrad (B = W as stored, MN-major) and wgrad (A = dY^T, B = X^T, both MN-major)
// run without transposes; the UMMA shared-memory descriptors and the TMA boxes change, the pipeline does not.
#include <stdlib.h>
#include "tc05.cuh"

namespace {

constexpr int BM = 128, BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;            // 16 KB
// two tile widths: BN = 256 (4-stage ring) for the big GEMMs, BN = 128 (6-stage ring) when a 256-wide tiling would leave SMs idle
template <int BN> struct Cfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024;   // + alignment slack
};
constexpr int GEMM_THREADS = 256;
constexpr int MAX_GROUPS = 8;

struct GemmParams {
  __nv_bfloat16* D;
  const __nv_bfloat16* bias;       // [N] or null
  float* D32;                      // optional fp32 output (accumulated: D32 += acc) instead of D
  int64_t ldd;
  int M, N, K;                     // dense problem (per group for grouped: M is the slab bound)
  int bn;                          // tile width (256 or 128)
  int beta;                        // 1: D = bf16(D + acc)
  int splits;                      // split-K factor (>1: fp32 atomic accumulation into D32)
  int swiglu;                      // CTA-pair kernel only: tile = [128 gate | 128 up] columns -> D[:, N/2] = bf16(bf16(silu(g)) * u)
  int dbg_nostore;                 // timing experiments only (LMOD_GEMM_NOSTORE=1): epilogue drains TMEM but does not write D
  // grouped (experts): row ranges from `offsets` (device), B / D32 advance per group
  const int32_t* offsets;          // [groups+1] or null
  int groups;
  int64_t b_group_rows;            // rows of B per group (N for K-major B) -- B coordinate offset
  int64_t d_group_stride;          // wgrad grouped: element stride of D per group
  int wgrad_grouped;               // 1: groups split the REDUCTION (K) range via offsets; M,N dense per group
  // dense only: extents read from DEVICE memory at kernel start (the loss head works on the batch's supervised rows, whose count is
  // data-dependent, inside a CUDA graph): M_eff = min(M, *m_dev), K_eff = min(K, round_up(*k_dev, BK)); null = static
  const int32_t* m_dev;
  const int32_t* k_dev;
};

__device__ __forceinline__ GemmParams effective_extents(const GemmParams& in) {
  GemmParams p = in;
  if (in.m_dev) p.M = min(in.M, max(0, *in.m_dev));
  if (in.k_dev) p.K = min(in.K, (max(0, *in.k_dev) + BK - 1) / BK * BK);
  return p;
}

// K-major 128B-swizzled tile (rows x 64 bf16, 128 B per row, 8-row atoms of 1024 B): LBO unused (=1), SBO = 1024; +32 B per UMMA_K
// MN-major 128B-swizzled tile (64 k-rows x 64 mn per 8 KB block): LBO = 8192 (next 64-wide mn block), SBO = 1024 (next 8 k-rows);
//   +2048 B per UMMA_K (16 k-rows)
template <bool MN>
__device__ __forceinline__ uint64_t operand_desc(uint32_t tile_saddr, int k16) {
  if (!MN) return smem_desc(tile_saddr + k16 * 32, 16, 1024);
  return smem_desc(tile_saddr + k16 * 2048, BK * 128, 1024);
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6) | a_format BF16 (1) [7,10) | b_format BF16 (1) [10,13) | a_major [15] | b_major [16]
// | N>>3 [17,23) | M>>4 [24,29)
template <int BN, bool A_MN, bool B_MN>
__device__ __forceinline__ uint32_t instr_desc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
         ((uint32_t)(BM >> 4) << 24);
}

struct Tile { int m0, n0, kb0, kb1, group, m_end; };

// tile index -> coordinates.  Dense: M fastest (consecutive CTAs share the B tile in L2).  Grouped forward/dgrad: per-group row
// ranges [offsets[g], offsets[g+1]) (128-row aligned by the router), B rows offset by g*b_group_rows.  Grouped wgrad: each group owns
// the reduction range [offsets[g], offsets[g+1]) and its own D.
__device__ __forceinline__ bool get_tile(const GemmParams& p, int t, Tile& o) {
  const int BN = p.bn;
  const int num_n = (p.N + BN - 1) / BN;
  if (p.offsets == nullptr) {
    const int num_m = (p.M + BM - 1) / BM;
    const int mn = num_m * num_n, nk = (p.K + BK - 1) / BK;
    if (t >= mn * p.splits) return false;
    const int ks = t / mn, r = t % mn;                       // output tiles fastest: concurrent CTAs hit different D tiles
    const int per = (nk + p.splits - 1) / p.splits;
    o.m0 = (r % num_m) * BM; o.n0 = (r / num_m) * BN; o.kb0 = ks * per; o.kb1 = min(nk, o.kb0 + per); o.group = 0; o.m_end = p.M;
    return true;
  }
  if (p.wgrad_grouped) {
    const int num_m = (p.M + BM - 1) / BM;
    const int per = num_m * num_n;
    const int g = t / per;
    if (g >= p.groups) return false;
    const int r = t % per;
    o.m0 = (r % num_m) * BM; o.n0 = (r / num_m) * BN; o.group = g; o.m_end = p.M;
    o.kb0 = p.offsets[g] / BK; o.kb1 = (p.offsets[g + 1] + BK - 1) / BK;
    return true;
  }
  int base = 0;
  for (int g = 0; g < p.groups; ++g) {
    const int r0 = p.offsets[g], r1 = p.offsets[g + 1];
    const int nm = (r1 - r0 + BM - 1) / BM;
    const int cnt = nm * num_n;
    if (t < base + cnt) {
      const int r = t - base;
      o.m0 = r0 + (r % nm) * BM; o.n0 = (r / nm) * BN; o.kb0 = 0; o.kb1 = (p.K + BK - 1) / BK; o.group = g; o.m_end = r1;
      return true;
    }
    base += cnt;
  }
  return false;
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p_in) {
  const GemmParams p = effective_extents(p_in);
  constexpr int STAGES = Cfg<BN>::STAGES, STAGE_BYTES = Cfg<BN>::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // SW128 needs 1024 B
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" :: "l"(&tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&tma_b) : "memory");
  }
  if (warp == 2) tmem_alloc(&tmem_base_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    uint32_t stage = 0, phase = 0;
    Tile t;
    for (int ti = blockIdx.x; get_tile(p, ti, t); ti += gridDim.x) {
      const int b_row0 = (int)(t.group * p.b_group_rows);
      for (int kb = t.kb0; kb < t.kb1; ++kb) {
        mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        uint8_t* sb = sa + A_STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
        if (!A_MN) {
          tma_load_2d(sa, &tma_a, kb * BK, t.m0, &full_bar[stage]);                      // box (64 k, 128 rows)
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * (BK * 128), &tma_a, t.m0 + 64 * j, kb * BK, &full_bar[stage]);   // box (64 mn, 64 k)
        }
        if (!B_MN) {
          tma_load_2d(sb, &tma_b, kb * BK, b_row0 + t.n0, &full_bar[stage]);             // box (64 k, 256 rows)
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * (BK * 128), &tma_b, t.n0 + 64 * j, b_row0 + kb * BK, &full_bar[stage]);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = instr_desc<BN, A_MN, B_MN>();
    uint32_t stage = 0, phase = 0, it = 0;
    Tile t;
    for (int ti = blockIdx.x; get_tile(p, ti, t); ti += gridDim.x, ++it) {
      const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
      mbar_wait_bounded(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      uint32_t first = 1;
      for (int kb = t.kb0; kb < t.kb1; ++kb) {
        mbar_wait_bounded(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          umma_f16(d_tmem, operand_desc<A_MN>(sa, k), operand_desc<B_MN>(sb, k), idesc, first ? 0u : 1u);
          first = 0;
        }
        umma_commit(&empty_bar[stage]);                    // frees the smem slot once these MMAs have read it
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tmem_full[acc]);                        // accumulator ready for the epilogue
    }
  } else if (warp >= 4) {
    // ===================== epilogue (TMEM -> registers -> global) =====================
    const int q = warp & 3;                                // TMEM lane quarter this warp may access
    uint32_t it = 0;
    Tile t;
    for (int ti = blockIdx.x; get_tile(p, ti, t); ti += gridDim.x, ++it) {
      const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
      mbar_wait_warp(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = t.m0 + q * 32 + lane;
      const bool row_ok = row < t.m_end;
      const bool empty_k = t.kb1 <= t.kb0;                 // grouped wgrad of an expert with no rows: contributes zero
      __nv_bfloat16* drow = p.D ? p.D + (p.wgrad_grouped ? t.group * p.d_group_stride : 0) + (int64_t)row * p.ldd : nullptr;
      float* d32row = p.D32 ? p.D32 + (p.wgrad_grouped ? t.group * p.d_group_stride : 0) + (int64_t)row * p.ldd : nullptr;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + acc * BN + c * 32 + ((uint32_t)(q * 32) << 16), r);
        const int col0 = t.n0 + c * 32;
        if (!row_ok || col0 >= p.N || p.dbg_nostore) continue;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int col = col0 + v * 8;
          if (col >= p.N) break;
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = empty_k ? 0.f : __uint_as_float(r[v * 8 + j]);
          if (p.bias) {
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.bias + col));
            f[0] += bf16lo(b.x); f[1] += bf16hi(b.x); f[2] += bf16lo(b.y); f[3] += bf16hi(b.y);
            f[4] += bf16lo(b.z); f[5] += bf16hi(b.z); f[6] += bf16lo(b.w); f[7] += bf16hi(b.w);
          }
          if (d32row && p.splits > 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(d32row + col + j, f[j]);
          } else if (d32row) {
            float4* o = reinterpret_cast<float4*>(d32row + col);
            float4 a = o[0], b2 = o[1];
            a.x += f[0]; a.y += f[1]; a.z += f[2]; a.w += f[3]; b2.x += f[4]; b2.y += f[5]; b2.z += f[6]; b2.w += f[7];
            o[0] = a; o[1] = b2;
          } else {
            uint4* o = reinterpret_cast<uint4*>(drow + col);
            if (p.beta) {
              const uint4 old = *o;
              f[0] += bf16lo(old.x); f[1] += bf16hi(old.x); f[2] += bf16lo(old.y); f[3] += bf16hi(old.y);
              f[4] += bf16lo(old.z); f[5] += bf16hi(old.z); f[6] += bf16lo(old.w); f[7] += bf16hi(old.w);
            }
            uint4 w;
            w.x = pack_bf16x2(f[0], f[1]); w.y = pack_bf16x2(f[2], f[3]); w.z = pack_bf16x2(f[4], f[5]); w.w = pack_bf16x2(f[6], f[7]);
            *o = w;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * BN);
}

// =================================================================================================================================
// 2-CTA variant (cta_group::2): a CTA PAIR (cluster of 2, same TPC) owns a 256 x 256 output tile.  Each CTA stages its own 128 rows
// of A and only HALF of B (128 of the 256 columns) -- the UMMA reads B from both CTAs' shared memory -- so shared-memory and L2
// traffic per FLOP halve.  The leader CTA (rank 0) issues tcgen05.mma.cta_group::2 (M = 256); both CTAs run a TMA producer (their
// loads complete_tx on the LEADER's full barrier) and an epilogue over their own 128 TMEM lanes.  tcgen05.commit multicasts the
// "stage free" / "accumulator ready" arrivals to both CTAs; the peer's epilogue warps arrive remotely on the leader's tmem_empty.
constexpr int STAGES2 = 6;
constexpr int B2_STAGE_BYTES = 128 * BK * 2;                    // half of the 256-wide B tile per CTA
constexpr int STAGE2_BYTES = A_STAGE_BYTES + B2_STAGE_BYTES;    // 32 KB
constexpr int GEMM2_SMEM = STAGES2 * STAGE2_BYTES + 1024;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;                     // clears the CTA-pair rank bit of a shared::cluster address -> leader CTA

__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* leader_bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               :: "r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(leader_bar) & PEER_MASK) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {       // arrive on the same-offset barrier of BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {     // arrive on the leader CTA's copy of `bar`
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(smem_u32(bar) & PEER_MASK) : "memory");
}

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p_in) {
  const GemmParams p = effective_extents(p_in);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES2], empty_bar[STAGES2], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int num_m = (p.M + 255) / 256, num_n = (p.N + 255) / 256, num_k = (p.K + BK - 1) / BK;
  const int ntiles = num_m * num_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }     // 4 epilogue warps x 2 CTAs (leader's copy)
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" :: "l"(&tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&tma_b) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // both CTAs' barriers initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer (both CTAs: own half of A rows and of B rows) =====================
    uint32_t stage = 0, phase = 0;
    for (int t = cluster; t < ntiles; t += nclusters) {
      const int m0 = (t % num_m) * 256 + 128 * (int)rank, n0 = (t / num_m) * 256 + 128 * (int)rank;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE2_BYTES;
        uint8_t* sb = sa + A_STAGE_BYTES;
        if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE2_BYTES);          // bytes of BOTH CTAs land on the leader's barrier
        if (!A_MN) {
          tma_load_2d_2sm(sa, &tma_a, kb * BK, m0, &full_bar[stage]);
        } else {
#pragma unroll
          for (int j = 0; j < 2; ++j) tma_load_2d_2sm(sa + j * (BK * 128), &tma_a, m0 + 64 * j, kb * BK, &full_bar[stage]);
        }
        if (!B_MN) {
          tma_load_2d_2sm(sb, &tma_b, kb * BK, n0, &full_bar[stage]);
        } else {
#pragma unroll
          for (int j = 0; j < 2; ++j) tma_load_2d_2sm(sb + j * (BK * 128), &tma_b, n0 + 64 * j, kb * BK, &full_bar[stage]);
        }
        if (++stage == STAGES2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ===================== MMA issuer (leader CTA only) =====================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(256 >> 3) << 17) |
                           ((uint32_t)(256 >> 4) << 24);
    uint32_t stage = 0, phase = 0, it = 0;
    for (int t = cluster; t < ntiles; t += nclusters, ++it) {
      const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
      mbar_wait_bounded(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      uint32_t first = 1;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait_bounded(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE2_BYTES), sb = sa + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          umma2_f16(d_tmem, operand_desc<A_MN>(sa, k), operand_desc<B_MN>(sb, k), idesc, first ? 0u : 1u);
          first = 0;
        }
        umma2_commit_mc(&empty_bar[stage]);
        if (++stage == STAGES2) { stage = 0; phase ^= 1; }
      }
      umma2_commit_mc(&tmem_full[acc]);
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs: own 128 rows x 256 columns) =====================
    const int q = warp & 3;
    uint32_t it = 0;
    for (int t = cluster; t < ntiles; t += nclusters, ++it) {
      const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
      mbar_wait_warp(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = (t % num_m) * 256 + 128 * (int)rank + q * 32 + lane;
      const int n00 = (t / num_m) * 256;
      const bool row_ok = row < p.M;
      __nv_bfloat16* drow = p.D ? p.D + (int64_t)row * p.ldd : nullptr;
      float* d32row = p.D32 ? p.D32 + (int64_t)row * p.ldd : nullptr;
      if (p.swiglu) {
        // fused SwiGLU epilogue (Qwen2MLP act_fn(gate_proj(x)) * up_proj(x), modeling_qwen2.py:199-200): the weight rows of this
        // 256-wide tile are [128 gate rows | 128 matching up rows], so gate and up of the same output column sit in this thread's lane
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t g[32], u[32];
          tmem_ld32(tmem_base + acc * 256 + c * 32 + ((uint32_t)(q * 32) << 16), g);
          tmem_ld32(tmem_base + acc * 256 + 128 + c * 32 + ((uint32_t)(q * 32) << 16), u);
          const int col0 = (n00 >> 1) + c * 32;
          if (!row_ok || col0 >= (p.N >> 1)) continue;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float gb = bf16_round(__uint_as_float(g[v * 8 + j])), ub = bf16_round(__uint_as_float(u[v * 8 + j]));   // reference rounds both GEMM outputs to bf16
              f[j] = bf16_round(gb * (1.f / (1.f + __expf(-gb)))) * ub;      // same expression as silu_mul_fwd_kernel -> bit-identical
            }
            uint4 w;
            w.x = pack_bf16x2(f[0], f[1]); w.y = pack_bf16x2(f[2], f[3]); w.z = pack_bf16x2(f[4], f[5]); w.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(drow + col0 + v * 8) = w;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + acc * 256 + c * 32 + ((uint32_t)(q * 32) << 16), r);
        const int col0 = n00 + c * 32;
        if (!row_ok || col0 >= p.N || p.dbg_nostore) continue;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int col = col0 + v * 8;
          if (col >= p.N) break;
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = (num_k == 0) ? 0.f : __uint_as_float(r[v * 8 + j]);   // empty (dynamic) reduction: nothing was accumulated
          if (p.bias) {
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.bias + col));
            f[0] += bf16lo(b.x); f[1] += bf16hi(b.x); f[2] += bf16lo(b.y); f[3] += bf16hi(b.y);
            f[4] += bf16lo(b.z); f[5] += bf16hi(b.z); f[6] += bf16lo(b.w); f[7] += bf16hi(b.w);
          }
          if (d32row) {
            float4* o = reinterpret_cast<float4*>(d32row + col);
            float4 a = o[0], b2 = o[1];
            a.x += f[0]; a.y += f[1]; a.z += f[2]; a.w += f[3]; b2.x += f[4]; b2.y += f[5]; b2.z += f[6]; b2.w += f[7];
            o[0] = a; o[1] = b2;
          } else {
            uint4* o = reinterpret_cast<uint4*>(drow + col);
            if (p.beta) {
              const uint4 old = *o;
              f[0] += bf16lo(old.x); f[1] += bf16hi(old.x); f[2] += bf16lo(old.y); f[3] += bf16hi(old.y);
              f[4] += bf16lo(old.z); f[5] += bf16hi(old.z); f[6] += bf16lo(old.w); f[7] += bf16hi(old.w);
            }
            uint4 w;
            w.x = pack_bf16x2(f[0], f[1]); w.y = pack_bf16x2(f[2], f[3]); w.z = pack_bf16x2(f[4], f[5]); w.w = pack_bf16x2(f[6], f[7]);
            *o = w;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // the pair frees its tensor memory together
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512) : "memory");
}

template <bool A_MN, bool B_MN>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    LMOD_CUDA_OK(cudaFuncSetAttribute(gemm2_tcgen05_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM2_SMEM));
    attr = true;
  }
  const int tiles = ((p.M + 255) / 256) * ((p.N + 255) / 256);
  int clusters = lmod_num_sms() / 2;
  if (tiles < clusters) clusters = tiles;
  gemm2_tcgen05_kernel<A_MN, B_MN><<<2 * clusters, GEMM_THREADS, GEMM2_SMEM, st>>>(ta, tb, p);
  LMOD_LAUNCH_OK();
  return LMOD_OK;
}
int dispatch2(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch2<false, false>(ta, tb, p, st);
  if (!a_mn && b_mn) return launch2<false, true>(ta, tb, p, st);
  if (a_mn && b_mn) return launch2<true, true>(ta, tb, p, st);
  return launch2<true, false>(ta, tb, p, st);
}

// ---------------------------------------------------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int tiles_upper, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    LMOD_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM));
    attr = true;
  }
  int grid = lmod_num_sms();
  if (tiles_upper < grid) grid = tiles_upper;
  if (grid < 1) grid = 1;
  gemm_tcgen05_kernel<BN, A_MN, B_MN><<<grid, GEMM_THREADS, Cfg<BN>::SMEM, st>>>(ta, tb, p);
  LMOD_LAUNCH_OK();
  return LMOD_OK;
}

template <int BN>
int dispatch_bn(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int tiles, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch<BN, false, false>(ta, tb, p, tiles, st);
  if (!a_mn && b_mn) return launch<BN, false, true>(ta, tb, p, tiles, st);
  if (a_mn && b_mn) return launch<BN, true, true>(ta, tb, p, tiles, st);
  return launch<BN, true, false>(ta, tb, p, tiles, st);
}
int dispatch(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int tiles, cudaStream_t st) {
  return p.bn == 256 ? dispatch_bn<256>(a_mn, b_mn, ta, tb, p, tiles, st) : dispatch_bn<128>(a_mn, b_mn, ta, tb, p, tiles, st);
}
// 256-wide tiles unless that tiling cannot fill the machine ~1.5 times over
int pick_bn(int64_t m_tiles, int64_t N) {
  const int64_t t256 = m_tiles * ((N + 255) / 256);
  return (t256 >= (int64_t)lmod_num_sms() * 3 / 2 || N <= 128) ? 256 : 128;
}

}  // namespace

// D[M,N] = A * B^T.  a_mn_major = 0: A stored [M,K] (row stride lda) ; 1: A stored [K,M].  b_mn_major = 0: B stored [N,K] ; 1: B stored [K,N].
// epilogue bit 0: D = bf16(D + acc) ;  d_f32_accum != null: fp32 D32 += acc (ldd applies to it) instead of the bf16 output;
// epilogue bit 1: fused SwiGLU (B rows tile-interleaved [128 gate | 128 up] per 256; D has N/2 columns; CTA-pair kernel only);
// epilogue bits 8..: split-K factor (fp32 atomic accumulation into D32, which the caller zero-initialises).
extern "C" int lmod_gemm_bf16_dyn(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major, void* D, int64_t ldd,
                                  int64_t M, int64_t N, int64_t K, const void* bias, int epilogue, float* d_f32_accum,
                                  const int32_t* m_rows_dev, const int32_t* k_rows_dev, void* stream) {
  LMOD_CHECK_ARG(A && B && (D || d_f32_accum) && M > 0 && N > 0 && K > 0, "lmod_gemm_bf16: null pointer or empty problem");
  LMOD_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && ldd % 8 == 0 && N % 8 == 0 && ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) &&
                 (!D || (uintptr_t)D % 16 == 0), "lmod_gemm_bf16: strides / N must be multiples of 8 elements and pointers 16-byte aligned (TMA)");
  CUtensorMap ta, tb;
  int rc;
  const int splits_req = (epilogue >> 8) > 1 ? (epilogue >> 8) : 1;
  static const int two_cta_env = getenv("LMOD_GEMM_2CTA") ? atoi(getenv("LMOD_GEMM_2CTA")) : 1;
  const int64_t tiles256 = ((M + 255) / 256) * ((N + 255) / 256);
  const bool pair = two_cta_env && splits_req == 1 && tiles256 >= (int64_t)(lmod_num_sms() / 2) * 3 / 2;
  if (epilogue & 2) LMOD_CHECK_ARG(pair && N % 256 == 0 && D && !d_f32_accum && !bias && !(epilogue & 1),
                                   "lmod_gemm_bf16: the SwiGLU epilogue needs the CTA-pair kernel (>= 111 256x256 tiles) and N %% 256 == 0");
  if (pair) {
    // CTA-pair kernel: 256 x 256 tiles, each CTA stages 128 rows of A and 128 rows of B
    if (!a_mn_major) rc = make_map(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, 128);
    else rc = make_map(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK);
    if (rc) return rc;
    if (!b_mn_major) rc = make_map(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, 128);
    else rc = make_map(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK);
    if (rc) return rc;
    GemmParams p2 = {};
    p2.D = (__nv_bfloat16*)D; p2.bias = (const __nv_bfloat16*)bias; p2.D32 = d_f32_accum; p2.ldd = ldd;
    p2.M = (int)M; p2.N = (int)N; p2.K = (int)K; p2.beta = epilogue & 1; p2.splits = 1; p2.groups = 1; p2.bn = 256;
    p2.dbg_nostore = getenv("LMOD_GEMM_NOSTORE") ? 1 : 0;
    p2.swiglu = (epilogue & 2) ? 1 : 0;
    p2.m_dev = m_rows_dev; p2.k_dev = k_rows_dev;
    return dispatch2(a_mn_major != 0, b_mn_major != 0, ta, tb, p2, (cudaStream_t)stream);
  }
  const int BN = pick_bn(((M + BM - 1) / BM) * splits_req, N);
  if (!a_mn_major) rc = make_map(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM);
  else rc = make_map(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK);
  if (rc) return rc;
  if (!b_mn_major) rc = make_map(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, BN);
  else rc = make_map(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK);
  if (rc) return rc;
  GemmParams p = {};
  p.D = (__nv_bfloat16*)D; p.bias = (const __nv_bfloat16*)bias; p.D32 = d_f32_accum; p.ldd = ldd;
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.beta = epilogue & 1; p.offsets = nullptr; p.groups = 1; p.bn = BN;
  p.dbg_nostore = getenv("LMOD_GEMM_NOSTORE") ? 1 : 0;
  p.splits = (epilogue >> 8) > 1 ? (epilogue >> 8) : 1;
  p.m_dev = m_rows_dev; p.k_dev = k_rows_dev;
  LMOD_CHECK_ARG(p.splits == 1 || d_f32_accum, "lmod_gemm_bf16: split-K needs the fp32 accumulate output");
  const int tiles = (int)(((M + BM - 1) / BM) * ((N + BN - 1) / BN)) * p.splits;
  return dispatch(a_mn_major != 0, b_mn_major != 0, ta, tb, p, tiles, (cudaStream_t)stream);
}

extern "C" int lmod_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major, void* D, int64_t ldd,
                              int64_t M, int64_t N, int64_t K, const void* bias, int epilogue, float* d_f32_accum, void* stream) {
  return lmod_gemm_bf16_dyn(A, lda, a_mn_major, B, ldb, b_mn_major, D, ldd, M, N, K, bias, epilogue, d_f32_accum, nullptr, nullptr, stream);
}

// 1 when lmod_gemm_bf16 would run this problem on the CTA-pair kernel (which is the one that offers the fused SwiGLU epilogue)
extern "C" int lmod_gemm_swiglu_ok(int64_t M, int64_t N) {
  static const int two_cta_env = getenv("LMOD_GEMM_2CTA") ? atoi(getenv("LMOD_GEMM_2CTA")) : 1;
  const int64_t tiles256 = ((M + 255) / 256) * ((N + 255) / 256);
  return (two_cta_env && N % 256 == 0 && tiles256 >= (int64_t)(lmod_num_sms() / 2) * 3 / 2) ? 1 : 0;
}

// Grouped (per-expert) GEMM on rows [offsets[g], offsets[g+1]) (device array; boundaries must be multiples of 128 for mode 0/1, of 64 for
// mode 2).   mode 0 (forward):  D[rows_g, N] = A[rows_g, K] * B[g][N,K]^T            (A,D [R,*] ; B [G,N,K])
//            mode 1 (dgrad)  :  D[rows_g, N] = A[rows_g, K] * B[g][K,N]              (B stored [G,K,N], MN-major)
//            mode 2 (wgrad)  :  D[g][M,N] (+)= A[rows_g, M]^T * B[rows_g, N]         (A stored [R,M], B stored [R,N], both MN-major; reduction
//                                                                                      over the group's rows)
extern "C" int lmod_grouped_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd, const int32_t* offsets, int G,
                                      int64_t max_rows, int64_t M, int64_t N, int64_t K, int mode, int epilogue, void* stream) {
  LMOD_CHECK_ARG(A && B && D && offsets && G >= 1 && G <= MAX_GROUPS && max_rows > 0, "lmod_grouped_gemm_bf16: bad arguments");
  LMOD_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && ldd % 8 == 0 && N % 8 == 0, "lmod_grouped_gemm_bf16: strides / N must be multiples of 8");
  CUtensorMap ta, tb;
  GemmParams p = {};
  p.D = (__nv_bfloat16*)D; p.ldd = ldd; p.beta = epilogue & 1; p.offsets = offsets; p.groups = G; p.splits = 1;
  int rc, tiles;
  const int BN = (mode == 2) ? pick_bn(G * ((M + BM - 1) / BM), N) : pick_bn((max_rows + BM - 1) / BM, N);
  p.bn = BN;
  if (mode == 0 || mode == 1) {
    rc = make_map(&ta, A, (uint64_t)K, (uint64_t)max_rows, (uint64_t)lda, BK, BM);
    if (rc) return rc;
    if (mode == 0) { rc = make_map(&tb, B, (uint64_t)K, (uint64_t)(G * N), (uint64_t)ldb, BK, BN); p.b_group_rows = N; }
    else { rc = make_map(&tb, B, (uint64_t)N, (uint64_t)(G * K), (uint64_t)ldb, 64, BK); p.b_group_rows = K; }
    if (rc) return rc;
    p.M = (int)max_rows; p.N = (int)N; p.K = (int)K;
    tiles = (int)(((max_rows + BM - 1) / BM + G) * ((N + BN - 1) / BN));
    return dispatch(false, mode == 1, ta, tb, p, tiles, (cudaStream_t)stream);
  }
  LMOD_CHECK_ARG(mode == 2, "lmod_grouped_gemm_bf16: mode must be 0, 1 or 2");
  rc = make_map(&ta, A, (uint64_t)M, (uint64_t)max_rows, (uint64_t)lda, 64, BK);
  if (rc) return rc;
  rc = make_map(&tb, B, (uint64_t)N, (uint64_t)max_rows, (uint64_t)ldb, 64, BK);
  if (rc) return rc;
  p.M = (int)M; p.N = (int)N; p.K = (int)max_rows; p.wgrad_grouped = 1; p.d_group_stride = M * ldd; p.b_group_rows = 0;
  tiles = (int)(G * ((M + BM - 1) / BM) * ((N + BN - 1) / BN));
  return dispatch(true, true, ta, tb, p, tiles, (cudaStream_t)stream);
}
