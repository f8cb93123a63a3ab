// attn_bwd.cu -- flash-attention BACKWARD on tcgen05 / TMEM / TMA (sm_100a).
//
// Backward of Qwen2SdpaAttention's scaled_dot_product_attention (modeling_qwen2.py:713-721) for the sparse student: given the fused
// RoPE'd QKV buffer, the forward output O, dO and the log-sum-exp of the forward kernel, produces dQ|dK|dV in one fused buffer.
//
// One CTA owns a block of 128 keys of one (batch, kv-head) and sweeps the 64-query blocks (and the query heads of its GQA group) that can
// see it.  Five tensor-core products per (key block, query block), all issued by one thread, accumulators in TMEM:
//     S^T  = K_j Q_i^T            (SS, M=128 N=64)        dP^T = V_j dO_i^T        (SS, M=128 N=64)
//     dV_j += P^T  dO_i           (TS: A = P^T  bf16 in TMEM, B = dO_i MN-major)   accumulates over the whole sweep
//     dK_j += dS^T Q_i            (TS: A = dS^T bf16 in TMEM, B = Q_i  MN-major)   accumulates over the whole sweep
//     dQ_i  = dS K_j              (SS, M=64: A = dS written to 128B-swizzled smem by the softmax threads, MN-major; B = K_j re-read MN-major)
// 256 softmax threads, two per key row (TMEM lane = key, 32 query columns each): P = exp2(S*c - lse), dS = P (dP - D) * scale with
// lse / D broadcast per query column.  Four more warps drain dQ_i to an fp32 workspace with red.global.add.v4 (a key block only holds
// a partial dQ) off the critical path; for head_dim 64 the score accumulators are double-buffered so that the tensor core computes
// S/dP of the next query block while the softmax threads work on the current one.
#include "tc05.cuh"

namespace {

constexpr int BKVB = 128, BQB = 64;
constexpr int ATB_THREADS = 448;     // TMA, MMA, 8 softmax warps, 4 dQ-drain warps

struct AttnBwdParams {
  const float* lse;       // [B, nh, T]
  const float* dsum;      // [B, nh, T]  D = rowsum(dO * O)
  float* dq32;            // [B*T, nh*hd] fp32 workspace (zero-initialised)
  __nv_bfloat16* dqkv;    // fused gradient buffer; this kernel writes the k and v columns
  int64_t ld_dqkv, ld_dq32;
  int B, T, nh, nkv;
  int causal;
  float scale, scale_log2;
  const int32_t* kv_lo;   // padded batches: real key range per batch row (see attn.cu); NULL = no padding
  const int32_t* kv_hi;
};

__device__ __forceinline__ uint32_t bwd_idesc(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void red_add_v4(float* dst, const uint32_t* r) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(dst), "f"(__uint_as_float(r[0])), "f"(__uint_as_float(r[1])),
               "f"(__uint_as_float(r[2])), "f"(__uint_as_float(r[3])) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

template <int HD>
__global__ void __launch_bounds__(ATB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tma_kv, const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_do,
                const AttnBwdParams p) {
  constexpr int KSUB = HD / 64;
  constexpr int NQ = 3;                                // Q_i / dO_i ring
  constexpr int NBUF = (HD == 64) ? 2 : 1;             // S^T/dP^T (and dS smem) double-buffered when TMEM has room: 2*HD + NBUF*128 + 64 <= 512
  constexpr int KV_BYTES = BKVB * HD * 2;              // one of K_j / V_j
  constexpr int Q_BYTES = BQB * HD * 2;                // one of Q_i / dO_i
  constexpr int DS_BYTES = BKVB * BQB * 2;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t kv_full, q_full[NQ], q_empty[NQ], s_full[2], ds_ready[2], dq_full, dq_empty, acc_full;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_lse2[2][BQB], s_dsum[2][BQB];      // per query block: lse*log2(e) and D, double buffered
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;
  uint8_t* sV = sK + KV_BYTES;
  uint8_t* sQ = sV + KV_BYTES;                         // stage s: Q at sQ + s*2*Q_BYTES, dO right after
  uint8_t* sDS = sQ + NQ * 2 * Q_BYTES;                // NBUF buffers of DS_BYTES
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // heavy key blocks (small j under the causal mask) are scheduled first: j is the slowest grid index
  const int j = blockIdx.z, hk = blockIdx.x, b = blockIdx.y;
  const int group = p.nh / p.nkv;
  const int kv0 = j * BKVB;
  const int nq = (p.T + BQB - 1) / BQB;
  int lo = 0, hi = p.T;
  if (p.kv_lo) { lo = p.kv_lo[b]; hi = p.kv_hi[b]; }
  const bool all_pad = hi <= lo;
  const bool padded = (lo > 0) || (hi < p.T);
  // un-masked query rows (rows in front of a left-padded sequence see every key) make every query block a partner of this key block
  const int i_start = (p.causal && lo == 0 && !all_pad) ? (kv0 / BQB) : 0;
  const int n_i = nq - i_start;                        // query blocks per head for this key block
  const int n_it = n_i * group;
  const int row_base = b * p.T;
  const int col_k = (p.nh + hk) * HD, col_v = (p.nh + p.nkv + hk) * HD;

  if (threadIdx.x == 0) {
    mbar_init(&kv_full, 1);
    for (int s = 0; s < NQ; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&ds_ready[s], 8); }
    mbar_init(&dq_full, 1); mbar_init(&dq_empty, 4); mbar_init(&acc_full, 1);
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" :: "l"(&tma_kv) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&tma_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&tma_do) : "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tDK = tmem, tDV = tmem + HD, tS0 = tmem + 2 * HD, tDQ = tS0 + NBUF * 2 * BQB;   // buffer b: S^T at tS0 + b*128, dP^T 64 columns later; dQ: HD columns

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    mbar_expect_tx(&kv_full, 2 * KV_BYTES);
#pragma unroll
    for (int i = 0; i < KSUB; ++i) {
      tma_load_2d(sK + i * (BKVB * 128), &tma_kv, col_k + 64 * i, row_base + kv0, &kv_full);
      tma_load_2d(sV + i * (BKVB * 128), &tma_kv, col_v + 64 * i, row_base + kv0, &kv_full);
    }
    for (int it = 0; it < n_it; ++it) {
      const int s = it % NQ;
      const int h = hk * group + it / n_i, qi = i_start + it % n_i;
      mbar_wait_bounded(&q_empty[s], ((it / NQ) & 1) ^ 1);
      uint8_t* q = sQ + s * 2 * Q_BYTES;
      uint8_t* d = q + Q_BYTES;
      mbar_expect_tx(&q_full[s], 2 * Q_BYTES);
#pragma unroll
      for (int i = 0; i < KSUB; ++i) {
        tma_load_2d(q + i * (BQB * 128), &tma_q, h * HD + 64 * i, row_base + qi * BQB, &q_full[s]);
        tma_load_2d(d + i * (BQB * 128), &tma_do, h * HD + 64 * i, row_base + qi * BQB, &q_full[s]);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    const uint32_t id_s = bwd_idesc(128, BQB, false, false);       // S^T, dP^T
    const uint32_t id_acc = bwd_idesc(128, HD, false, true);       // dV, dK   (A from TMEM, B MN-major)
    const uint32_t id_dq = bwd_idesc(BQB, HD, true, true);         // dQ       (A = dS MN-major: 64 queries contiguous per key row, B = K_j MN-major)
    const uint32_t aK = smem_u32(sK), aV = smem_u32(sV);
    // S^T = K Q^T and dP^T = V dO^T of iteration `it` into buffer it % NBUF
    auto issue_scores = [&](int it) {
      const int s = it % NQ, buf = it % NBUF;
      const uint32_t aQ = smem_u32(sQ + s * 2 * Q_BYTES), aDO = aQ + Q_BYTES;
      const uint32_t tS = tS0 + buf * 2 * BQB, tDP = tS + BQB;
      mbar_wait_bounded(&q_full[s], (it / NQ) & 1);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        const uint32_t ko = (k / 4), ki = (k % 4) * 32;
        umma_f16(tS, smem_desc(aK + ko * (BKVB * 128) + ki, 16, 1024), smem_desc(aQ + ko * (BQB * 128) + ki, 16, 1024), id_s, k > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        const uint32_t ko = (k / 4), ki = (k % 4) * 32;
        umma_f16(tDP, smem_desc(aV + ko * (BKVB * 128) + ki, 16, 1024), smem_desc(aDO + ko * (BQB * 128) + ki, 16, 1024), id_s, k > 0 ? 1u : 0u);
      }
      umma_commit(&s_full[buf]);
    };
    // dV += P^T dO, dK += dS^T Q, dQ = dS K of iteration `it`
    auto issue_grads = [&](int it) {
      const int s = it % NQ, buf = it % NBUF;
      const uint32_t aQ = smem_u32(sQ + s * 2 * Q_BYTES), aDO = aQ + Q_BYTES, aDS = smem_u32(sDS + buf * DS_BYTES);
      const uint32_t tS = tS0 + buf * 2 * BQB, tDP = tS + BQB;
      mbar_wait_bounded(&ds_ready[buf], (it / NBUF) & 1);
      if (it > 0) mbar_wait_bounded(&dq_empty, (it - 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < BQB / 16; ++k) {         // reduction over the 64 queries of the block
        const uint32_t ka = (k >> 1) * 32 + (k & 1) * 8;       // 16 queries = 8 bf16x2 columns; the second 32 queries start at column 32
        umma_f16_ts(tDV, tS + ka, smem_desc(aDO + k * 2048, BQB * 128, 1024), id_acc, (it > 0 || k > 0) ? 1u : 0u);
        umma_f16_ts(tDK, tDP + ka, smem_desc(aQ + k * 2048, BQB * 128, 1024), id_acc, (it > 0 || k > 0) ? 1u : 0u);
      }
#pragma unroll
      for (int k = 0; k < BKVB / 16; ++k) {        // reduction over the 128 keys
        umma_f16(tDQ, smem_desc(aDS + k * 2048, 16, 1024), smem_desc(aK + k * 2048, BKVB * 128, 1024), id_dq, k > 0 ? 1u : 0u);
      }
      umma_commit(&q_empty[s]);
      umma_commit(&dq_full);
    };
    mbar_wait_bounded(&kv_full, 0);
    if (NBUF == 2) {
      // scores of it+1 are in flight (other S buffer, other Q stage) while the softmax threads work on it
      issue_scores(0);
      for (int it = 0; it < n_it; ++it) {
        if (it + 1 < n_it) issue_scores(it + 1);
        issue_grads(it);
      }
    } else {
      for (int it = 0; it < n_it; ++it) { issue_scores(it); issue_grads(it); }
    }
    umma_commit(&acc_full);
  } else if (warp >= 2 && warp < 10) {
    // ===================== softmax / dS: two threads per key row (32 of the 64 query columns each) =====================
    const int q = warp & 3;
    const int g = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int kv = kv0 + r;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    // lse*log2(e) and D of a query block are staged one iteration ahead: the first 128 threads (g == 0) load them from global memory
    // before waiting on the scores and publish them after their own math, so the L2 latency is off the critical path
    auto stage_load = [&](int it) -> float {
      const int h = hk * group + it / n_i, q0 = (i_start + it % n_i) * BQB;
      const int qc = min(q0 + (r & 63), p.T - 1);                       // rows >= T are masked below, clamp the address
      const int64_t base = ((int64_t)b * p.nh + h) * p.T + qc;
      return (r < 64) ? __ldg(p.lse + base) * LOG2E_F : __ldg(p.dsum + base);
    };
    auto stage_store = [&](int it, float v) {
      if (r < 64) s_lse2[it & 1][r & 63] = v; else s_dsum[it & 1][r & 63] = v;
    };
    if (g == 0 && n_it > 0) stage_store(0, stage_load(0));
    asm volatile("bar.sync 1, 256;" ::: "memory");
    for (int it = 0; it < n_it; ++it) {
      const int qi = i_start + it % n_i;
      const int q0 = qi * BQB;
      const int buf = it % NBUF;
      const uint32_t tS = tS0 + buf * 2 * BQB, tDP = tS + BQB;
      uint8_t* ds_row = sDS + buf * DS_BYTES + r * 128;
      float staged = 0.f;
      const bool prefetch = (g == 0) && (it + 1 < n_it);
      if (prefetch) staged = stage_load(it + 1);
      const float* lse2 = s_lse2[it & 1] + g * 32;
      const float* dsm_s = s_dsum[it & 1] + g * 32;
      // only the blocks on the causal diagonal and the ragged tail need per-element masks
      const bool masked = padded || (q0 + BQB > p.T) || (kv0 + BKVB > p.T) || (p.causal && q0 < kv0 + BKVB - 1);
      mbar_wait_warp(&s_full[buf], (it / NBUF) & 1);
      tc_fence_after();
      uint32_t sv[32], dv[32];
      tmem_ld32(tS + g * 32 + lane_off, sv);
      tmem_ld32(tDP + g * 32 + lane_off, dv);
      uint32_t pk[16], dk[16];
      if (masked) {
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float pp[2], dd[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int qidx = q0 + g * 32 + c + e;
            const bool seen = (all_pad || qidx < lo) ? true : (kv >= lo && kv < hi && (!p.causal || kv <= qidx));
            const bool ok = (qidx < p.T) && (kv < p.T) && seen;
            const float pv = ok ? ex2f(fmaf(__uint_as_float(sv[c + e]), p.scale_log2, -lse2[c + e])) : 0.f;
            pp[e] = pv;
            dd[e] = pv * (__uint_as_float(dv[c + e]) - dsm_s[c + e]) * p.scale;
          }
          pk[c >> 1] = pack_bf16x2(pp[0], pp[1]);
          dk[c >> 1] = pack_bf16x2(dd[0], dd[1]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float pp[2], dd[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float pv = ex2f(fmaf(__uint_as_float(sv[c + e]), p.scale_log2, -lse2[c + e]));
            pp[e] = pv;
            dd[e] = pv * (__uint_as_float(dv[c + e]) - dsm_s[c + e]) * p.scale;
          }
          pk[c >> 1] = pack_bf16x2(pp[0], pp[1]);
          dk[c >> 1] = pack_bf16x2(dd[0], dd[1]);
        }
      }
      // P^T over S^T, dS^T over dP^T (bf16x2: this thread's 32 queries = 16 columns)
      // each thread overwrites only columns it has read itself (its partner on the same row runs unsynchronised): queries 32g..32g+31
      // land in columns 32g..32g+15
      tmem_st8(tS + g * 32 + lane_off, pk);
      tmem_st8(tS + g * 32 + 8 + lane_off, pk + 8);
      tmem_st8(tDP + g * 32 + lane_off, dk);
      tmem_st8(tDP + g * 32 + 8 + lane_off, dk + 8);
      // dS for dQ = dS K: row = key, 64 queries contiguous (MN-major A operand), 128B swizzle (16-byte chunk ^ (row & 7))
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int cidx = g * 4 + ch;
        uint4 w = make_uint4(dk[ch * 4], dk[ch * 4 + 1], dk[ch * 4 + 2], dk[ch * 4 + 3]);
        *reinterpret_cast<uint4*>(ds_row + ((cidx ^ (r & 7)) << 4)) = w;
      }
      tmem_st_wait();
      fence_proxy_async();                               // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_ready[buf]);
      if (prefetch) stage_store(it + 1, staged);
      asm volatile("bar.sync 1, 256;" ::: "memory");     // next block's lse / D visible; this block's are free to be overwritten
    }
    // ---- dK_j, dV_j epilogue: the two threads of a row split the 32-column chunks ----
    mbar_wait_warp(&acc_full, 0);
    tc_fence_after();
    {
      // tcgen05.ld is warp-collective (.sync.aligned): every lane executes the loads, only the stores are predicated on the row
      const bool row_ok = kv < p.T;
      __nv_bfloat16* dkrow = p.dqkv + (int64_t)(row_base + kv) * p.ld_dqkv + col_k;
      __nv_bfloat16* dvrow = p.dqkv + (int64_t)(row_base + kv) * p.ld_dqkv + col_v;
#pragma unroll
      for (int cc = 0; cc < HD / 64; ++cc) {
        const int c = cc * 2 + g;
        uint32_t o[32];
        tmem_ld32(tDK + c * 32 + lane_off, o);
        if (row_ok) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[v * 8 + 0]), __uint_as_float(o[v * 8 + 1]));
            w.y = pack_bf16x2(__uint_as_float(o[v * 8 + 2]), __uint_as_float(o[v * 8 + 3]));
            w.z = pack_bf16x2(__uint_as_float(o[v * 8 + 4]), __uint_as_float(o[v * 8 + 5]));
            w.w = pack_bf16x2(__uint_as_float(o[v * 8 + 6]), __uint_as_float(o[v * 8 + 7]));
            *reinterpret_cast<uint4*>(dkrow + c * 32 + v * 8) = w;
          }
        }
        tmem_ld32(tDV + c * 32 + lane_off, o);
        if (row_ok) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[v * 8 + 0]), __uint_as_float(o[v * 8 + 1]));
            w.y = pack_bf16x2(__uint_as_float(o[v * 8 + 2]), __uint_as_float(o[v * 8 + 3]));
            w.z = pack_bf16x2(__uint_as_float(o[v * 8 + 4]), __uint_as_float(o[v * 8 + 5]));
            w.w = pack_bf16x2(__uint_as_float(o[v * 8 + 6]), __uint_as_float(o[v * 8 + 7]));
            *reinterpret_cast<uint4*>(dvrow + c * 32 + v * 8) = w;
          }
        }
      }
    }
  } else if (warp >= 10) {
    // ===================== dQ_i drain: M = 64 accumulator, query row (16*quarter + lane) in lanes 0..15 of every quarter, columns = head dim;
    // vector fp32 red.add into the workspace (a key block only holds a partial dQ) =====================
    const int q = warp & 3;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int row = q * 16 + lane;
    for (int it = 0; it < n_it; ++it) {
      const int h = hk * group + it / n_i, qi = i_start + it % n_i;
      const int q0 = qi * BQB;
      const bool row_ok = (lane < 16) && (q0 + row < p.T);
      float* dst = p.dq32 + (int64_t)(row_base + q0 + row) * p.ld_dq32 + h * HD;
      mbar_wait_warp(&dq_full, it & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < HD / 64; ++c) {
        uint32_t o0[32], o1[32];
        tmem_ld32(tDQ + c * 64 + lane_off, o0);
        tmem_ld32(tDQ + c * 64 + 32 + lane_off, o1);
        if (c == HD / 64 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&dq_empty);                        // registers hold the tile: the tensor core may overwrite dQ now
        }
        if (row_ok) {
#pragma unroll
          for (int v = 0; v < 8; ++v) red_add_v4(dst + c * 64 + v * 4, o0 + v * 4);
#pragma unroll
          for (int v = 0; v < 8; ++v) red_add_v4(dst + c * 64 + 32 + v * 4, o1 + v * 4);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// D[b,h,t] = sum_d dO[t,h,d] * O[t,h,d]   (one warp per (t,h))
template <int HD>
__global__ void __launch_bounds__(256) attn_dsum_kernel(const __nv_bfloat16* __restrict__ o, int64_t ld_o, const __nv_bfloat16* __restrict__ dout,
                                                       int64_t ld_do, int B, int T, int nh, float* __restrict__ dsum) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (int64_t)B * T * nh) return;
  const int h = (int)(w % nh);
  const int64_t row = w / nh;                      // b*T + t
  const uint32_t* po = reinterpret_cast<const uint32_t*>(o + row * ld_o + h * HD);
  const uint32_t* pd = reinterpret_cast<const uint32_t*>(dout + row * ld_do + h * HD);
  float a = 0.f;
#pragma unroll
  for (int i = lane; i < HD / 2; i += 32) {
    const uint32_t x = __ldg(po + i), y = __ldg(pd + i);
    a = fmaf(bf16lo(x), bf16lo(y), a);
    a = fmaf(bf16hi(x), bf16hi(y), a);
  }
  a = warp_sum(a);
  if (lane == 0) dsum[((row / T) * nh + h) * (int64_t)T + (row % T)] = a;
}

// dq32 [rows, nh*hd] fp32 -> q columns of the fused bf16 gradient buffer
__global__ void attn_dq_convert_kernel(const float* __restrict__ dq32, int64_t ld32, int64_t rows, int cols, __nv_bfloat16* __restrict__ dqkv, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per 8 columns
  const int cv = cols >> 3;
  if (i >= rows * cv) return;
  const int64_t row = i / cv;
  const int c = (int)(i % cv) * 8;
  const float4 a = *reinterpret_cast<const float4*>(dq32 + row * ld32 + c), b2 = *reinterpret_cast<const float4*>(dq32 + row * ld32 + c + 4);
  uint4 w;
  w.x = pack_bf16x2(a.x, a.y); w.y = pack_bf16x2(a.z, a.w); w.z = pack_bf16x2(b2.x, b2.y); w.w = pack_bf16x2(b2.z, b2.w);
  *reinterpret_cast<uint4*>(dqkv + row * ld + c) = w;
}

template <int HD>
int launch_attn_bwd(const CUtensorMap& tkv, const CUtensorMap& tq, const CUtensorMap& tdo, const AttnBwdParams& p, cudaStream_t st) {
  constexpr int SMEM = 2 * BKVB * HD * 2 + 6 * BQB * HD * 2 + ((HD == 64) ? 2 : 1) * BKVB * BQB * 2 + 1024;
  static bool attr = false;
  if (!attr) {
    LMOD_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr = true;
  }
  dim3 grid(p.nkv, p.B, (p.T + BKVB - 1) / BKVB);
  attn_bwd_kernel<HD><<<grid, ATB_THREADS, SMEM, st>>>(tkv, tq, tdo, p);
  LMOD_LAUNCH_OK();
  return LMOD_OK;
}

}  // namespace

// qkv / dqkv: fused [batch*seq, (nh+2nkv)*hd]; out, dout: [batch*seq, nh*hd]; lse [batch, nh, seq] from lmod_attn_fwd.
// dq32_ws: fp32 [batch*seq, nh*hd] workspace, dsum_ws: fp32 [batch, nh, seq] workspace (both written here; dq32 is zeroed inside).
// kv_lo / kv_hi: as in lmod_attn_fwd (padded batches), or NULL.
extern "C" int lmod_attn_bwd(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_o, const void* dout, int64_t ld_do, const float* lse,
                             int64_t batch, int64_t seq, int nh, int nkv, int hd, int causal, float softmax_scale, void* dqkv, int64_t ld_dqkv,
                             float* dq32_ws, float* dsum_ws, const int32_t* kv_lo, const int32_t* kv_hi, void* stream) {
  LMOD_CHECK_ARG((kv_lo == nullptr) == (kv_hi == nullptr), "lmod_attn_bwd: kv_lo and kv_hi come together");
  LMOD_CHECK_ARG(qkv && out && dout && lse && dqkv && dq32_ws && dsum_ws && batch > 0 && seq > 0 && nh % nkv == 0, "lmod_attn_bwd: bad arguments");
  LMOD_CHECK_ARG(hd == 64 || hd == 128, "lmod_attn_bwd: head_dim %d not built (64 and 128 are)", hd);
  LMOD_CHECK_ARG(ld_qkv % 8 == 0 && ld_o % 8 == 0 && ld_do % 8 == 0 && ld_dqkv % 8 == 0, "lmod_attn_bwd: strides must be multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = batch * seq;
  const int qcols = nh * hd;
  LMOD_CUDA_OK(cudaMemsetAsync(dq32_ws, 0, (size_t)rows * qcols * sizeof(float), st));
  {
    const int64_t warps = rows * nh;
    const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
    if (hd == 128) attn_dsum_kernel<128><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)out, ld_o, (const __nv_bfloat16*)dout, ld_do, (int)batch, (int)seq, nh, dsum_ws);
    else attn_dsum_kernel<64><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)out, ld_o, (const __nv_bfloat16*)dout, ld_do, (int)batch, (int)seq, nh, dsum_ws);
    LMOD_LAUNCH_OK();
  }
  CUtensorMap tkv, tq, tdo;
  const uint64_t cols = (uint64_t)(nh + 2 * nkv) * hd;
  int rc = make_map(&tkv, qkv, cols, (uint64_t)rows, (uint64_t)ld_qkv, 64, BKVB);
  if (rc) return rc;
  rc = make_map(&tq, qkv, cols, (uint64_t)rows, (uint64_t)ld_qkv, 64, BQB);
  if (rc) return rc;
  rc = make_map(&tdo, dout, (uint64_t)qcols, (uint64_t)rows, (uint64_t)ld_do, 64, BQB);
  if (rc) return rc;
  AttnBwdParams p;
  p.lse = lse; p.dsum = dsum_ws; p.dq32 = dq32_ws; p.dqkv = (__nv_bfloat16*)dqkv; p.ld_dqkv = ld_dqkv; p.ld_dq32 = qcols;
  p.B = (int)batch; p.T = (int)seq; p.nh = nh; p.nkv = nkv; p.causal = causal; p.scale = softmax_scale; p.scale_log2 = softmax_scale * LOG2E_F;
  p.kv_lo = kv_lo; p.kv_hi = kv_hi;
  rc = (hd == 128) ? launch_attn_bwd<128>(tkv, tq, tdo, p, st) : launch_attn_bwd<64>(tkv, tq, tdo, p, st);
  if (rc) return rc;
  const int64_t n = rows * (qcols / 8);
  attn_dq_convert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dq32_ws, qcols, rows, qcols, (__nv_bfloat16*)dqkv, ld_dqkv);
  LMOD_LAUNCH_OK();
  return LMOD_OK;
}
