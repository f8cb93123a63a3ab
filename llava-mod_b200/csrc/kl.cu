// kl.cu -- fused mimic-KL (+ shifted CE) forward+backward over the vocabulary.
//
// Replaces (reference) llavamod/train/align_trainer.py:473-475 (teacher softmax fp32), :497-499
// (student log_softmax fp32), :509-526 (product / masked_fill / vocab sum / masked mean) and
// llava_qwen1_5_moe.py:413-421 (shifted CE), i.e. >= 8 full [N,V] passes, by ONE sweep:
//
//   * a row (V = 151936 bf16 logits, student + teacher = 608 KB) does not fit one SM's shared
//     memory, so a thread-block CLUSTER of 8 CTAs owns a row: each CTA pulls its 1/8 slice of both
//     rows into shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx),
//     reduces max / sum-exp / sum p_T*s locally, exchanges 6 floats through distributed shared
//     memory (one barrier.cluster per row), then produces the gradient slice from the SAME shared
//     memory copy.  HBM traffic is therefore exactly the algorithmic 4*V bytes read + 2*V written
//     per token; nothing is re-read.
//   * two CTAs are resident per SM (<= 2 x 76 KB smem) so one CTA's loads overlap the other's math.
//   * rows whose KD mask and CE mask are both 0 contribute nothing (reference multiplies by 0):
//     their loads are skipped and their gradient slice is zero-filled.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace {

constexpr int KL_CHUNKS = 4;       // mbarrier-tracked load chunks per slice
constexpr int KL_MAX_CS = 8;
constexpr int KL_MIN_CTAS = 2;       // two CTAs per SM (2 x 76 KB of shared memory)
constexpr int KL_DEFAULT_MODE = 5;      // sb128.  measured (profiles/kl_modes_r1.txt, all rows active): sb128 0.60 ms, sb256 0.64, sb384 0.72, sb512 0.90, db256 1.10, db512 1.24

struct KlParams {
  const __nv_bfloat16* s;
  const __nv_bfloat16* t;
  const int64_t* labels;
  const float* counts;   // {n_kd, n_ce}
  float* row_out;        // [N,4]
  __nv_bfloat16* d;      // may be null / alias s
  int64_t ld_s, ld_t, ld_d;
  int64_t n_rows, seq_len;
  int vocab, slice;      // slice: elements per CTA (multiple of 8)
  int distill_all;
  float w_kd, w_ce;
  // compact mode (rows.cu): s / t / d hold only the batch's active rows, row j of them is original row perm[j] (labels, row_out);
  // the row count comes from device memory.  null = dense layout, n_rows rows.
  const int32_t* perm;
  const int32_t* count;
};

struct Xchg {            // per-CTA partials published to the cluster
  float ms, mt, zs, zt, a, slab, pad0, pad1;
};

__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t hmin2_u32(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmin2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

template <bool CHECK_INF>
__device__ __forceinline__ void accum_pair(uint32_t sw, uint32_t tw, float nms, float nmt, float& zs,
                                           float& zt, float& a, float& zk) {
  float s0 = bf16lo(sw), s1 = bf16hi(sw), t0 = bf16lo(tw), t1 = bf16hi(tw);
  float es0 = ex2f(fmaf(s0, LOG2E_F, nms)), es1 = ex2f(fmaf(s1, LOG2E_F, nms));
  float et0 = ex2f(fmaf(t0, LOG2E_F, nmt)), et1 = ex2f(fmaf(t1, LOG2E_F, nmt));
  zs += es0; zs += es1;
  zt += et0; zt += et1;
  if (CHECK_INF) {   // align_trainer.py:509-510: terms where log q_S is +-inf are dropped
    if (isinf(s0)) { s0 = 0.f; et0 = 0.f; }
    if (isinf(s1)) { s1 = 0.f; et1 = 0.f; }
    zk += et0; zk += et1;      // teacher mass of the KEPT terms: x = sum_kept p_T*(s - lse_S)
  }
  a = fmaf(et0, s0, a);
  a = fmaf(et1, s1, a);
}

// NBUF = 1 (default): one slice buffer, two CTAs per SM hide each other's loads AND each other's barrier waits (33 co-resident clusters).
// NBUF = 2 (LMOD_KL_MODE=db256/db512, experiment kept for the record): one CTA per SM with two slice buffers, the loads of the cluster's
// NEXT active row issued before the math of the current one.  Measured 1.6-1.8x SLOWER: with one CTA per SM nothing fills the SM while
// the CTA sits in __syncthreads / barrier.cluster (22 % of warp samples), and only 15 clusters of 8 single-CTA SMs fit the GPCs.
template <int KL_THREADS, int NBUF>
__global__ void __launch_bounds__(KL_THREADS, NBUF == 2 ? 1 : KL_MIN_CTAS) kl_fused_kernel(const KlParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(16) Xchg xchg[2];
  __shared__ __align__(8) uint64_t bars_all[NBUF][KL_CHUNKS];
  __shared__ float red[7][KL_THREADS / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank(), cs = cluster_nctarank();
  const uint32_t cid = cluster_id_x(), ncl = cluster_nclusters_x();

  const size_t buf_bytes = (size_t)p.slice * 4;                 // student + teacher slice

  // this CTA's slice of the vocabulary
  const int v0 = (int)rank * p.slice;
  int len = p.vocab - v0;
  len = len < 0 ? 0 : (len > p.slice ? p.slice : len);
  const int nvec = len >> 3;                                   // 16-byte vectors (8 bf16)
  const int cvec = ((nvec + KL_CHUNKS - 1) / KL_CHUNKS);        // vectors per chunk

  if (tid == 0) {
    for (int b = 0; b < NBUF; ++b)
      for (int c = 0; c < KL_CHUNKS; ++c) mbar_init(&bars_all[b][c], 1);
    mbar_fence_init();
  }
  __syncthreads();
  cluster_sync_all();

  const float n_kd = p.counts[0], n_ce = p.counts[1];
  uint32_t it_active = 0;

  const int64_t n_rows = p.count ? (int64_t)*p.count : p.n_rows;
  for (int64_t row = cid; row < n_rows; row += ncl) {
    // ---- masks (uniform over the cluster) ----
    const int64_t orow = p.perm ? (int64_t)p.perm[row] : row;     // where this row sits in the batch: labels and row_out are indexed by it
    const int64_t lab_here = p.labels[orow];
    const int64_t tpos = orow % p.seq_len;
    int64_t lab_next = LMOD_IGNORE_INDEX;
    if (tpos + 1 < p.seq_len) lab_next = p.labels[orow + 1];
    const bool m_kd = p.distill_all ? true : (lab_here != LMOD_IGNORE_INDEX);
    const bool m_ce = (lab_next != LMOD_IGNORE_INDEX);   // nll is always reported (loss/lm metric); w_ce only scales its gradient
    const bool active = m_kd || m_ce;

    if (!active) {
      if (p.d != nullptr) {
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4* dst = reinterpret_cast<uint4*>(p.d + row * p.ld_d + v0);
        for (int i = tid; i < nvec; i += KL_THREADS) stg_v4(dst + i, z);
      }
      if (rank == 0 && tid == 0) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(p.row_out + orow * 4) = o;
      }
      continue;
    }
    const uint32_t xi = it_active & 1u;                                        // exchange slot
    const uint32_t bsel = (NBUF == 2) ? (it_active & 1u) : 0u;                  // slice buffer
    const uint32_t par = (NBUF == 2) ? ((it_active >> 1) & 1u) : (it_active & 1u);
    uint64_t* bars = bars_all[bsel];
    uint4* s_buf = reinterpret_cast<uint4*>(smem_raw + bsel * buf_bytes);
    uint4* t_buf = reinterpret_cast<uint4*>(smem_raw + bsel * buf_bytes + (size_t)p.slice * 2);

    // ---- issue the slice loads (1-D TMA bulk copies), chunked so math can start early ----
    auto issue_row = [&](int64_t r, uint32_t b) {
      if (nvec <= 0) return;
      const __nv_bfloat16* srow = p.s + r * p.ld_s + v0;
      const __nv_bfloat16* trow = p.t + r * p.ld_t + v0;
      uint4* sb = reinterpret_cast<uint4*>(smem_raw + b * buf_bytes);
      uint4* tb = reinterpret_cast<uint4*>(smem_raw + b * buf_bytes + (size_t)p.slice * 2);
      for (int c = 0; c < KL_CHUNKS; ++c) {
        int cb = c * cvec, ce = min(nvec, cb + cvec);
        if (ce <= cb) { mbar_arrive(&bars_all[b][c]); continue; }
        uint32_t bytes = (uint32_t)(ce - cb) * 16u;
        mbar_expect_tx(&bars_all[b][c], 2 * bytes);
        bulk_g2s(sb + cb, srow + (size_t)cb * 8, bytes, &bars_all[b][c]);
        bulk_g2s(tb + cb, trow + (size_t)cb * 8, bytes, &bars_all[b][c]);
      }
    };
    if (NBUF == 1) {
      if (tid == 0) issue_row(row, 0);
    } else if (warp == 0) {
      // the first active row loads itself; every active row then looks ahead (32 candidate rows per ballot) for the cluster's next
      // active row and starts ITS loads into the other buffer, which the previous row released at its closing __syncthreads
      if (it_active == 0 && lane == 0) issue_row(row, 0);
      int64_t nxt = -1;
      for (int64_t base = row; base + ncl < n_rows; base += 32 * (int64_t)ncl) {
        const int64_t r = base + (int64_t)(lane + 1) * ncl;
        bool act = false;
        if (r < n_rows) {
          const int64_t ro = p.perm ? (int64_t)p.perm[r] : r;
          act = p.distill_all || (p.labels[ro] != LMOD_IGNORE_INDEX);
          if (!act && (ro % p.seq_len) + 1 < p.seq_len) act = p.labels[ro + 1] != LMOD_IGNORE_INDEX;
        }
        const unsigned m = __ballot_sync(0xffffffffu, act);
        if (m) { nxt = base + (int64_t)__ffs(m) * ncl; break; }
      }
      if (lane == 0 && nxt >= 0) issue_row(nxt, bsel ^ 1u);
    }
    ++it_active;

    // ---- pass A+B (one sweep, chunk by chunk as the bulk copies land): running maxima with rescaled partial sums -------------------
    // per thread: (m_s, zs = sum e^{s-m_s}) and (m_t, zt = sum e^{t-m_t}, acc = sum e^{t-m_t} s, zk = kept teacher mass).  The running
    // maxima settle after a few vectors, so the rescale branch is cold.  Infinite logits never enter a running maximum: -inf terms vanish
    // on their own, +inf keeps the old degenerate outcome (sum-exp = inf); vectors that hold an infinite student logit take the checked path
    // (align_trainer.py:509-510 drops those terms).
    float m_s = -INFINITY, m_t = -INFINITY, zs = 0.f, zt = 0.f, acc = 0.f, zk = 0.f;
    if (nvec > 0) {
      for (int c = 0; c < KL_CHUNKS; ++c) {
        const int b0 = c * cvec, e0 = min(nvec, b0 + cvec);
        mbar_wait(&bars[c], par);
        for (int i = b0 + tid; i < e0; i += KL_THREADS) {
          const uint4 sv = s_buf[i], tv = t_buf[i];
          const uint32_t pmx_s = hmax2_u32(hmax2_u32(sv.x, sv.y), hmax2_u32(sv.z, sv.w));
          const uint32_t pmn_s = hmin2_u32(hmin2_u32(sv.x, sv.y), hmin2_u32(sv.z, sv.w));
          const uint32_t pmx_t = hmax2_u32(hmax2_u32(tv.x, tv.y), hmax2_u32(tv.z, tv.w));
          const float vmax_s = fmaxf(bf16lo(pmx_s), bf16hi(pmx_s)), vmin_s = fminf(bf16lo(pmn_s), bf16hi(pmn_s));
          const float vmax_t = fmaxf(bf16lo(pmx_t), bf16hi(pmx_t));
          if (vmax_s > m_s && !isinf(vmax_s)) {
            zs *= ex2f((m_s - vmax_s) * LOG2E_F);                      // m_s = -inf the first time: zs = 0 * 0
            m_s = vmax_s;
          }
          if (vmax_t > m_t && !isinf(vmax_t)) {
            const float f = ex2f((m_t - vmax_t) * LOG2E_F);
            zt *= f; acc *= f; zk *= f;
            m_t = vmax_t;
          }
          const float nms = isinf(m_s) ? 0.f : -m_s * LOG2E_F, nmt = isinf(m_t) ? 0.f : -m_t * LOG2E_F;
          if (!(isinf(vmin_s) || isinf(vmax_s))) {
            float zk_unused = 0.f;
            accum_pair<false>(sv.x, tv.x, nms, nmt, zs, zt, acc, zk_unused);
            accum_pair<false>(sv.y, tv.y, nms, nmt, zs, zt, acc, zk_unused);
            accum_pair<false>(sv.z, tv.z, nms, nmt, zs, zt, acc, zk_unused);
            accum_pair<false>(sv.w, tv.w, nms, nmt, zs, zt, acc, zk_unused);
            // all terms kept: the kept teacher mass of this vector equals its share of zt (added below through dzt)
          } else {
            const float zt0 = zt;
            float zk_v = 0.f;
            accum_pair<true>(sv.x, tv.x, nms, nmt, zs, zt, acc, zk_v);
            accum_pair<true>(sv.y, tv.y, nms, nmt, zs, zt, acc, zk_v);
            accum_pair<true>(sv.z, tv.z, nms, nmt, zs, zt, acc, zk_v);
            accum_pair<true>(sv.w, tv.w, nms, nmt, zs, zt, acc, zk_v);
            zk += zk_v - (zt - zt0);                                   // zk tracks (kept - all) teacher mass; the final zk = zt + this
          }
        }
      }
    }
    zk += zt;                                                         // kept teacher mass on this thread's scale m_t
    // ---- block reduction with the (max, scaled sums) combine; one __syncthreads ----
    {
      float Ms = warp_max(m_s), Mt = warp_max(m_t);
      const float fs = isinf(m_s) ? 0.f : ex2f((m_s - Ms) * LOG2E_F);   // Ms finite whenever some lane's m_s is
      const float ft = isinf(m_t) ? 0.f : ex2f((m_t - Mt) * LOG2E_F);
      zs = warp_sum(isinf(zs) ? zs : zs * fs);                          // an infinite sum-exp (+inf logit) stays infinite
      zt = warp_sum(isinf(zt) ? zt : zt * ft);
      acc = warp_sum(acc * ft); zk = warp_sum(zk * ft);
      if (lane == 0) { red[0][warp] = Ms; red[1][warp] = Mt; red[3][warp] = zs; red[4][warp] = zt; red[5][warp] = acc; red[6][warp] = zk; }
    }
    __syncthreads();
    float ms = -INFINITY, mt = -INFINITY;
    if (warp == 0) {
      constexpr int NW = KL_THREADS / 32;
      const float wms = (lane < NW) ? red[0][lane] : -INFINITY, wmt = (lane < NW) ? red[1][lane] : -INFINITY;
      ms = warp_max(wms); mt = warp_max(wmt);
      const float fs = isinf(wms) ? 0.f : ex2f((wms - ms) * LOG2E_F);
      const float ft = isinf(wmt) ? 0.f : ex2f((wmt - mt) * LOG2E_F);
      float a = (lane < NW) ? red[3][lane] : 0.f, b = (lane < NW) ? red[4][lane] : 0.f;
      float c = (lane < NW) ? red[5][lane] : 0.f, k = (lane < NW) ? red[6][lane] : 0.f;
      a = warp_sum(isinf(a) ? a : a * fs); b = warp_sum(isinf(b) ? b : b * ft);
      c = warp_sum(c * ft); k = warp_sum(k * ft);
      if (lane == 0) {
        float slab = 0.f;
        if (m_ce) {
          int64_t off = lab_next - v0;
          if (off >= 0 && off < len)
            slab = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(s_buf)[off]);
        }
        Xchg x;
        x.ms = (nvec > 0) ? ms : -INFINITY; x.mt = (nvec > 0) ? mt : -INFINITY;
        x.zs = a; x.zt = b; x.a = c; x.slab = slab; x.pad0 = k; x.pad1 = 0.f;
        xchg[xi] = x;
      }
    }
    // ---- cluster exchange through distributed shared memory ----
    cluster_sync_all();
    float lse_s, lse_t, xrow, slab;
    {
      float r_ms = -INFINITY, r_mt = -INFINITY, r_zs = 0.f, r_zt = 0.f, r_a = 0.f, r_sl = 0.f, r_zk = 0.f;
      if ((uint32_t)lane < cs) {
        const float* base = reinterpret_cast<const float*>(&xchg[xi]);
        r_ms = dsmem_ld_f32(base + 0, lane); r_mt = dsmem_ld_f32(base + 1, lane);
        r_zs = dsmem_ld_f32(base + 2, lane); r_zt = dsmem_ld_f32(base + 3, lane);
        r_a = dsmem_ld_f32(base + 4, lane);  r_sl = dsmem_ld_f32(base + 5, lane); r_zk = dsmem_ld_f32(base + 6, lane);
      }
      float Ms = warp_max(r_ms), Mt = warp_max(r_mt);
      float Ms_u = isinf(Ms) ? 0.f : Ms, Mt_u = isinf(Mt) ? 0.f : Mt;
      float fs = isinf(r_ms) ? 0.f : ex2f((r_ms - Ms_u) * LOG2E_F);
      float ft = isinf(r_mt) ? 0.f : ex2f((r_mt - Mt_u) * LOG2E_F);
      float Zs = warp_sum(r_zs * fs), Zt = warp_sum(r_zt * ft), A = warp_sum(r_a * ft), Zk = warp_sum(r_zk * ft);
      slab = warp_sum(r_sl);
      lse_s = Ms_u + lg2f(Zs) * LN2_F;
      lse_t = Mt_u + lg2f(Zt) * LN2_F;
      xrow = (A - lse_s * Zk) / Zt;
    }
    if (rank == 0 && tid == 0) {
      float4 o = make_float4(xrow, m_ce ? (lse_s - slab) : 0.f, lse_s, lse_t);
      *reinterpret_cast<float4*>(p.row_out + orow * 4) = o;
    }

    // ---- pass C: gradient slice straight from shared memory ----
    if (p.d != nullptr) {
      const float ckd = m_kd ? (p.w_kd / n_kd) : 0.f;
      const float cce = m_ce ? (p.w_ce / n_ce) : 0.f;
      const float ca = ckd + cce, cb = ckd;
      const float es = -lse_s * LOG2E_F, et = -lse_t * LOG2E_F;
      const int lab_local = m_ce ? (int)(lab_next - v0) : -1;
      uint4* dst = reinterpret_cast<uint4*>(p.d + row * p.ld_d + v0);
      for (int i = tid; i < nvec; i += KL_THREADS) {
        uint4 sv = s_buf[i], tv = t_buf[i];
        float g[8];
        const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
        const uint32_t tw[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float q0 = ex2f(fmaf(bf16lo(sw[j]), LOG2E_F, es)), q1 = ex2f(fmaf(bf16hi(sw[j]), LOG2E_F, es));
          float p0 = ex2f(fmaf(bf16lo(tw[j]), LOG2E_F, et)), p1 = ex2f(fmaf(bf16hi(tw[j]), LOG2E_F, et));
          g[2 * j] = fmaf(ca, q0, -cb * p0);
          g[2 * j + 1] = fmaf(ca, q1, -cb * p1);
        }
        const unsigned rel = (unsigned)(lab_local - i * 8);
        if (rel < 8u) {
#pragma unroll
          for (int j = 0; j < 8; ++j) if (rel == (unsigned)j) g[j] -= cce;
        }
        uint4 o;
        o.x = pack_bf16x2(g[0], g[1]); o.y = pack_bf16x2(g[2], g[3]);
        o.z = pack_bf16x2(g[4], g[5]); o.w = pack_bf16x2(g[6], g[7]);
        stg_v4(dst + i, o);
      }
    }
    __syncthreads();   // all generic-proxy reads of the slice are done before the next bulk load lands
  }
  cluster_sync_all();  // keep this CTA's shared memory alive until every peer finished its DSMEM reads
}

__global__ void kl_counts_kernel(const int64_t* __restrict__ labels, int64_t n, int64_t T, int distill_all,
                                 float* __restrict__ out) {
  __shared__ float red[32];
  float a = 0.f, b = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    a += (distill_all || labels[i] != LMOD_IGNORE_INDEX) ? 1.f : 0.f;
    if ((i % T) + 1 < T) b += (labels[i + 1] != LMOD_IGNORE_INDEX) ? 1.f : 0.f;
  }
  a = block_sum(a, red);
  b = block_sum(b, red);
  if (threadIdx.x == 0) { out[0] = a; out[1] = b; }
}

__global__ void kl_finalize_kernel(const float* __restrict__ row_out, const int64_t* __restrict__ labels,
                                   int64_t n, int64_t T, int distill_all, float* __restrict__ out) {
  __shared__ float red[32];
  float sx = 0.f, sn = 0.f, ck = 0.f, cc = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    float4 r = *reinterpret_cast<const float4*>(row_out + i * 4);
    bool m = distill_all || labels[i] != LMOD_IGNORE_INDEX;
    bool c = ((i % T) + 1 < T) && labels[i + 1] != LMOD_IGNORE_INDEX;
    if (m) { sx += r.x; ck += 1.f; }
    if (c) { sn += r.y; cc += 1.f; }
  }
  sx = block_sum(sx, red); sn = block_sum(sn, red); ck = block_sum(ck, red); cc = block_sum(cc, red);
  if (threadIdx.x == 0) {
    out[0] = -sx / ck;     // 0/0 -> NaN like align_trainer.py:526
    out[1] = sn / cc;
    out[2] = ck;
    out[3] = cc;
  }
}

}  // namespace

extern "C" int lmod_kl_counts(const int64_t* labels, int64_t n_rows, int64_t seq_len, int distill_all,
                              float* counts2, void* stream) {
  LMOD_CHECK_ARG(labels && counts2 && n_rows > 0 && seq_len > 0 && n_rows % seq_len == 0,
                 "lmod_kl_counts: bad arguments (n_rows=%lld seq_len=%lld)", (long long)n_rows, (long long)seq_len);
  kl_counts_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(labels, n_rows, seq_len, distill_all, counts2);
  LMOD_LAUNCH_OK();
  return LMOD_OK;
}

extern "C" int lmod_kl_finalize(const float* row_out, const int64_t* labels, int64_t n_rows, int64_t seq_len,
                                int distill_all, float* out4, void* stream) {
  LMOD_CHECK_ARG(row_out && labels && out4 && n_rows > 0 && seq_len > 0, "lmod_kl_finalize: bad arguments");
  kl_finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(row_out, labels, n_rows, seq_len, distill_all, out4);
  LMOD_LAUNCH_OK();
  return LMOD_OK;
}

template <int T, int NB>
int kl_launch(KlParams p, int cs, size_t smem, int64_t n_rows, cudaStream_t stream) {
  static bool attr_done = false;
  if (!attr_done) {
    LMOD_CUDA_OK(cudaFuncSetAttribute(kl_fused_kernel<T, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_done = true;
  }
  smem *= NB;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(T); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cfg.gridDim = dim3(cs);
  static int cached_clusters[2] = {0, 0};        // per cluster size (1 / 8): queried once, outside any stream capture
  int& max_clusters = cached_clusters[cs == 1 ? 0 : 1];   // (one multi-CTA cluster size per process)
  if (max_clusters <= 0) {
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kl_fused_kernel<T, NB>, &cfg);
    if (e != cudaSuccess || max_clusters <= 0) { (void)cudaGetLastError(); max_clusters = lmod_num_sms() / cs; }
    if (getenv("LMOD_KL_VERBOSE")) fprintf(stderr, "[lmod] kl_fused_kernel<%d,%d>: cluster %d, %zu B smem, %d co-resident clusters\n", T, NB, cs, smem, max_clusters);
  }
  int64_t ncl = n_rows < max_clusters ? n_rows : max_clusters;
  cfg.gridDim = dim3((unsigned)(ncl * cs));
  LMOD_CUDA_OK(cudaLaunchKernelEx(&cfg, kl_fused_kernel<T, NB>, p));
  lmod_count_launch();
  return LMOD_OK;
}

extern "C" int lmod_kl_fwd_bwd_rows(const void* s_logits, int64_t ld_s, const void* t_logits, int64_t ld_t,
                                    const int64_t* labels, int64_t n_rows, int64_t seq_len, int64_t vocab,
                                    int distill_all, float w_kd, float w_ce, const float* counts2,
                                    float* row_out, void* dlogits, int64_t ld_d, const int32_t* perm, const int32_t* count, void* stream) {
  LMOD_CHECK_ARG(s_logits && t_logits && labels && counts2 && row_out, "lmod_kl_fwd_bwd: null pointer");
  LMOD_CHECK_ARG(n_rows > 0 && seq_len > 0 && n_rows % seq_len == 0, "lmod_kl_fwd_bwd: n_rows %% seq_len != 0");
  LMOD_CHECK_ARG(vocab >= 8 && vocab % 8 == 0 && ld_s % 8 == 0 && ld_t % 8 == 0 && ld_s >= vocab && ld_t >= vocab,
                 "lmod_kl_fwd_bwd: vocab and row strides must be multiples of 8 elements (16-byte TMA bulk copies)");
  LMOD_CHECK_ARG(((uintptr_t)s_logits % 16 == 0) && ((uintptr_t)t_logits % 16 == 0), "lmod_kl_fwd_bwd: pointers must be 16B aligned");
  if (dlogits) LMOD_CHECK_ARG(ld_d % 8 == 0 && ld_d >= vocab && ((uintptr_t)dlogits % 16 == 0), "lmod_kl_fwd_bwd: bad dlogits stride");

  int cs = (vocab >= 512) ? KL_MAX_CS : 1;      // (a 16-CTA cluster with 38 KB slices, 4 CTAs/SM, measured 2.1x slower)
  int64_t per = (vocab + cs - 1) / cs;
  int slice = (int)((per + 7) / 8 * 8);
  size_t smem = (size_t)slice * 2 * 2;
  LMOD_CHECK_ARG(smem <= 220 * 1024, "lmod_kl_fwd_bwd: vocab %lld too large for the 8-CTA cluster layout", (long long)vocab);

  KlParams p;
  p.s = (const __nv_bfloat16*)s_logits; p.t = (const __nv_bfloat16*)t_logits; p.labels = labels;
  p.counts = counts2; p.row_out = row_out; p.d = (__nv_bfloat16*)dlogits;
  p.ld_s = ld_s; p.ld_t = ld_t; p.ld_d = ld_d; p.n_rows = n_rows; p.seq_len = seq_len;
  p.vocab = (int)vocab; p.slice = slice; p.distill_all = distill_all; p.w_kd = w_kd; p.w_ce = w_ce;
  LMOD_CHECK_ARG((perm == nullptr) == (count == nullptr), "lmod_kl_fwd_bwd_rows: perm and count go together");
  p.perm = perm; p.count = count;

  // LMOD_KL_MODE: "sb256" one slice buffer, 2 CTAs/SM (round-1 first version); "db256"/"db512" double-buffered slice, 1 CTA/SM
  static const char* mode_env = getenv("LMOD_KL_MODE");
  static const int mode = !mode_env ? KL_DEFAULT_MODE
                          : (!strcmp(mode_env, "sb256") ? 0 : (!strcmp(mode_env, "db256") ? 1 : (!strcmp(mode_env, "sb512") ? 3 : (!strcmp(mode_env, "sb384") ? 4 : (!strcmp(mode_env, "sb128") ? 5 : 2)))));
  const bool db_fits = smem * 2 <= 220 * 1024;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 3) return kl_launch<512, 1>(p, cs, smem, n_rows, st);
  if (mode == 4) return kl_launch<384, 1>(p, cs, smem, n_rows, st);
  if (mode == 5) return kl_launch<128, 1>(p, cs, smem, n_rows, st);
  if (mode == 0 || !db_fits) return kl_launch<256, 1>(p, cs, smem, n_rows, st);
  return (mode == 1) ? kl_launch<256, 2>(p, cs, smem, n_rows, st) : kl_launch<512, 2>(p, cs, smem, n_rows, st);
}

extern "C" int lmod_kl_fwd_bwd(const void* s_logits, int64_t ld_s, const void* t_logits, int64_t ld_t,
                               const int64_t* labels, int64_t n_rows, int64_t seq_len, int64_t vocab,
                               int distill_all, float w_kd, float w_ce, const float* counts2,
                               float* row_out, void* dlogits, int64_t ld_d, void* stream) {
  return lmod_kl_fwd_bwd_rows(s_logits, ld_s, t_logits, ld_t, labels, n_rows, seq_len, vocab, distill_all, w_kd, w_ce, counts2, row_out,
                              dlogits, ld_d, nullptr, nullptr, stream);
}
