// moe.cu -- DeepSpeed-0.9.5 top-2 MoE routing on B200: gate + capacity + token scatter fused in one
// cooperative kernel; weighted gather/combine; and their backward kernels.
//
// Replaces (third-party, call site llavamod/model/language_model/llava_qwen1_5_moe.py:536-546)
// deepspeed.moe.sharded_moe.TopKGate / top2gating / MOELayer dispatch+combine einsums
// (SURVEY.md Appendix A): ~40 small ATen kernels, a D2H sync (exp_counts.to('cpu')) and two one-hot
// einsums per MoE layer become: route_scatter (1 launch) + expert GEMMs + gather_combine (1 launch).
//
// Semantics kept bit-exact for the integer record (idx1, idx2, slot, kept) given the same fp32 logits
// and the same Gumbel noise tensor: first choices are seated before any second choice, positions are
// a STABLE prefix count over the flattened [B*T] token order (warp-ballot + popc, no atomics races),
// capacity C = ceil(S/E * cf * 2), drops by position >= C.
#include <cooperative_groups.h>
#include <float.h>
#include "common.cuh"

namespace {

constexpr int RT_THREADS = 256;
constexpr int RT_WARPS = RT_THREADS / 32;
constexpr int MAXE = 8;

struct RouteParams {
  const __nv_bfloat16* x;
  const float* wg;
  const float* noise;
  int S, H, E;
  int capacity;
  int layout;            // 0: compact rows ; 1: offsets[e] = e*capacity (capacity-padded slabs) ; 2: compact, groups aligned to 128 rows
  int stage_cap;         // tokens per block whose rows are kept in shared memory across the grid barrier
  float* logits; float* gates; int32_t* idx; int32_t* row; float* w;
  int32_t* offsets; float* meta; __nv_bfloat16* xp;
  unsigned int* sync_ws;
};

__device__ __forceinline__ void grid_barrier(unsigned int* ws, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&ws[0], 1u);
    while (*reinterpret_cast<volatile unsigned int*>(&ws[0]) < nblocks) { __nanosleep(64); }
    __threadfence();
    unsigned int left = atomicAdd(&ws[1], 1u);
    if (left == nblocks - 1) { ws[0] = 0u; ws[1] = 0u; __threadfence(); }   // last one out re-arms the barrier
  }
  __syncthreads();
}

__global__ void __launch_bounds__(RT_THREADS, 1) moe_route_scatter_kernel(const RouteParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  // layout: wg fp32 [E,H] | stage bf16 [stage_cap,H] | idx bytes [2*S] | small arrays
  float* s_wg = reinterpret_cast<float*>(smem);
  __nv_bfloat16* s_stage = reinterpret_cast<__nv_bfloat16*>(smem + (size_t)p.E * p.H * 4);
  uint8_t* s_idx = smem + (size_t)p.E * p.H * 4 + (size_t)p.stage_cap * p.H * 2;
  __shared__ int s_tot1[MAXE], s_tot2[MAXE], s_pre1[MAXE], s_pre2[MAXE], s_off[MAXE + 1];
  __shared__ float s_red[32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  const int tpb = (p.S + G - 1) / G;
  const int b0 = min(p.S, (int)blockIdx.x * tpb), b1 = min(p.S, b0 + tpb);
  const int E = p.E, H = p.H, hv = H >> 3;

  for (int i = tid; i < E * H / 4; i += RT_THREADS)
    reinterpret_cast<float4*>(s_wg)[i] = reinterpret_cast<const float4*>(p.wg)[i];
  if (tid < MAXE) { s_tot1[tid] = 0; s_tot2[tid] = 0; s_pre1[tid] = 0; s_pre2[tid] = 0; }
  __syncthreads();

  // ---------------- phase 1: gate (warp per token), stage the token rows in shared memory ----------------
  for (int tok = b0 + warp; tok < b1; tok += RT_WARPS) {
    const int loc = tok - b0;
    const uint4* xr = reinterpret_cast<const uint4*>(p.x + (size_t)tok * H);
    float acc[MAXE];
#pragma unroll
    for (int e = 0; e < MAXE; ++e) acc[e] = 0.f;
    for (int v = lane; v < hv; v += 32) {
      uint4 u = ldg_nc_v4(xr + v);
      if (loc < p.stage_cap) reinterpret_cast<uint4*>(s_stage + (size_t)loc * H)[v] = u;
      float xf[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y), bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        if (e < E) {
          const float4 w0 = *reinterpret_cast<const float4*>(s_wg + (size_t)e * H + v * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(s_wg + (size_t)e * H + v * 8 + 4);
          acc[e] = fmaf(xf[0], w0.x, acc[e]); acc[e] = fmaf(xf[1], w0.y, acc[e]);
          acc[e] = fmaf(xf[2], w0.z, acc[e]); acc[e] = fmaf(xf[3], w0.w, acc[e]);
          acc[e] = fmaf(xf[4], w1.x, acc[e]); acc[e] = fmaf(xf[5], w1.y, acc[e]);
          acc[e] = fmaf(xf[6], w1.z, acc[e]); acc[e] = fmaf(xf[7], w1.w, acc[e]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < MAXE; ++e) acc[e] = warp_sum(acc[e]);
    if (lane == 0) {
      float mx = -INFINITY;
      int i1 = 0;
#pragma unroll
      for (int e = 0; e < MAXE; ++e) if (e < E && acc[e] > mx) { mx = acc[e]; i1 = e; }   // first max wins ties
      float ex[MAXE], den = 0.f;
#pragma unroll
      for (int e = 0; e < MAXE; ++e) { ex[e] = (e < E) ? expf(acc[e] - mx) : 0.f; den += ex[e]; }
      float best = -INFINITY;
      int i2 = (i1 == 0) ? 1 : 0;
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        if (e < E && e != i1) {
          float v = acc[e] + p.noise[(size_t)tok * E + e];
          if (v > best) { best = v; i2 = e; }
        }
      }
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        if (e < E) { p.logits[(size_t)tok * E + e] = acc[e]; p.gates[(size_t)tok * E + e] = ex[e] / den; }
      }
      p.idx[2 * tok] = i1; p.idx[2 * tok + 1] = i2;
    }
  }

  grid_barrier(p.sync_ws, G);

  // ---------------- phase 2: stable positions (every block recomputes the global prefix counts) ----------------
  for (int i = tid; i < p.S; i += RT_THREADS) {
    int2 v = __ldcg(reinterpret_cast<const int2*>(p.idx) + i);
    s_idx[2 * i] = (uint8_t)v.x; s_idx[2 * i + 1] = (uint8_t)v.y;
  }
  __syncthreads();
  {
    int t1[MAXE], t2[MAXE], q1[MAXE], q2[MAXE];
#pragma unroll
    for (int e = 0; e < MAXE; ++e) { t1[e] = t2[e] = q1[e] = q2[e] = 0; }
    const int ngroups = (p.S + 31) >> 5;
    for (int g = warp; g < ngroups; g += RT_WARPS) {
      const int tok = (g << 5) + lane;
      const bool valid = tok < p.S;
      const int e1 = valid ? s_idx[2 * tok] : -1, e2 = valid ? s_idx[2 * tok + 1] : -1;
      const bool before = tok < b0;
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        if (e < E) {
          unsigned a = __ballot_sync(0xffffffffu, e1 == e), b = __ballot_sync(0xffffffffu, e2 == e);
          unsigned ab = __ballot_sync(0xffffffffu, before && e1 == e), bb = __ballot_sync(0xffffffffu, before && e2 == e);
          t1[e] += __popc(a); t2[e] += __popc(b); q1[e] += __popc(ab); q2[e] += __popc(bb);
        }
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        if (e < E) { atomicAdd(&s_tot1[e], t1[e]); atomicAdd(&s_tot2[e], t2[e]); atomicAdd(&s_pre1[e], q1[e]); atomicAdd(&s_pre2[e], q2[e]); }
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    int o = 0, used = 0;
    for (int e = 0; e < E; ++e) {
      s_off[e] = (p.layout == 1) ? e * p.capacity : o;
      int rows_e = min(s_tot1[e] + s_tot2[e], p.capacity);
      used += rows_e;
      o += (p.layout == 2) ? ((rows_e + 127) & ~127) : rows_e;
    }
    s_off[E] = (p.layout == 1) ? E * p.capacity : o;
    if (blockIdx.x == 0) {
      for (int e = 0; e <= E; ++e) p.offsets[e] = s_off[e];
      p.meta[1] = (float)p.capacity; p.meta[2] = (float)used; p.meta[3] = (float)s_off[E];
      for (int e = 0; e < E; ++e) p.meta[4 + e] = (float)s_tot1[e];
    }
  }
  __syncthreads();
  // warp 0 seats this block's tokens in order (32 at a time, ballot + popc prefix)
  if (warp == 0) {
    int c1[MAXE], c2[MAXE];
#pragma unroll
    for (int e = 0; e < MAXE; ++e) { c1[e] = (e < E) ? s_pre1[e] : 0; c2[e] = (e < E) ? s_pre2[e] : 0; }
    const unsigned lt = (1u << lane) - 1u;
    for (int base = b0; base < b1; base += 32) {
      const int tok = base + lane;
      const bool valid = tok < b1;
      const int e1 = valid ? s_idx[2 * tok] : -1, e2 = valid ? s_idx[2 * tok + 1] : -1;
      int loc1 = 0, loc2 = 0;
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        if (e < E) {
          unsigned a = __ballot_sync(0xffffffffu, e1 == e), b = __ballot_sync(0xffffffffu, e2 == e);
          if (e1 == e) loc1 = c1[e] + __popc(a & lt);
          if (e2 == e) loc2 = s_tot1[e] + c2[e] + __popc(b & lt);
          c1[e] += __popc(a); c2[e] += __popc(b);
        }
      }
      if (valid) {
        const bool k1 = loc1 < p.capacity, k2 = loc2 < p.capacity;
        const float g1 = k1 ? __ldcg(p.gates + (size_t)tok * E + e1) : 0.f;
        const float g2 = k2 ? __ldcg(p.gates + (size_t)tok * E + e2) : 0.f;
        const float den = fmaxf(g1 + g2, FLT_EPSILON);
        p.row[2 * tok] = k1 ? s_off[e1] + loc1 : -1;
        p.row[2 * tok + 1] = k2 ? s_off[e2] + loc2 : -1;
        p.w[2 * tok] = g1 / den; p.w[2 * tok + 1] = g2 / den;
      }
    }
  }
  // l_aux = E * sum_e mean_s(gates[:,e]) * mean_s(mask1[:,e])   (before capacity drops)
  if (blockIdx.x == 0) {
    float laux = 0.f;
    for (int e = 0; e < E; ++e) {
      float a = 0.f;
      for (int s = tid; s < p.S; s += RT_THREADS) a += __ldcg(p.gates + (size_t)s * E + e);
      a = block_sum(a, s_red);
      laux += (a / (float)p.S) * ((float)s_tot1[e] / (float)p.S);
    }
    if (tid == 0) p.meta[0] = laux * (float)E;
  }
  __syncthreads();   // rows of this block's tokens (global, written by warp 0) visible to the block

  // ---------------- phase 3: scatter the token rows to their expert rows ----------------
  for (int tok = b0 + warp; tok < b1; tok += RT_WARPS) {
    const int loc = tok - b0;
    const int r1 = p.row[2 * tok], r2 = p.row[2 * tok + 1];
    const uint4* src = (loc < p.stage_cap) ? reinterpret_cast<const uint4*>(s_stage + (size_t)loc * H)
                                           : reinterpret_cast<const uint4*>(p.x + (size_t)tok * H);
    for (int v = lane; v < hv; v += 32) {
      uint4 u = src[v];
      if (r1 >= 0) stg_v4(reinterpret_cast<uint4*>(p.xp + (size_t)r1 * H) + v, u);
      if (r2 >= 0) stg_v4(reinterpret_cast<uint4*>(p.xp + (size_t)r2 * H) + v, u);
    }
  }
}

// out[s] = bf16( bf16(w1)*y[row1] + bf16(w2)*y[row2] ) (+ residual, rounded again like the reference's bf16 add)
__global__ void __launch_bounds__(256) moe_gather_combine_kernel(const __nv_bfloat16* __restrict__ y, const int32_t* __restrict__ row,
                                                                const float* __restrict__ w, const __nv_bfloat16* __restrict__ res,
                                                                int S, int H, __nv_bfloat16* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= S) return;
  const int r1 = row[2 * warp], r2 = row[2 * warp + 1];
  const float w1 = bf16_round(w[2 * warp]), w2 = bf16_round(w[2 * warp + 1]);
  const int hv = H >> 3;
  for (int v = lane; v < hv; v += 32) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    if (r1 >= 0) {
      uint4 u = ldg_nc_v4(reinterpret_cast<const uint4*>(y + (size_t)r1 * H) + v);
      a[0] = w1 * bf16lo(u.x); a[1] = w1 * bf16hi(u.x); a[2] = w1 * bf16lo(u.y); a[3] = w1 * bf16hi(u.y);
      a[4] = w1 * bf16lo(u.z); a[5] = w1 * bf16hi(u.z); a[6] = w1 * bf16lo(u.w); a[7] = w1 * bf16hi(u.w);
    }
    if (r2 >= 0) {
      uint4 u = ldg_nc_v4(reinterpret_cast<const uint4*>(y + (size_t)r2 * H) + v);
      a[0] = fmaf(w2, bf16lo(u.x), a[0]); a[1] = fmaf(w2, bf16hi(u.x), a[1]); a[2] = fmaf(w2, bf16lo(u.y), a[2]);
      a[3] = fmaf(w2, bf16hi(u.y), a[3]); a[4] = fmaf(w2, bf16lo(u.z), a[4]); a[5] = fmaf(w2, bf16hi(u.z), a[5]);
      a[6] = fmaf(w2, bf16lo(u.w), a[6]); a[7] = fmaf(w2, bf16hi(u.w), a[7]);
    }
    if (res) {
      uint4 u = ldg_nc_v4(reinterpret_cast<const uint4*>(res + (size_t)warp * H) + v);
      float rr[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y), bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = rr[j] + bf16_round(a[j]);
    }
    uint4 o;
    o.x = pack_bf16x2(a[0], a[1]); o.y = pack_bf16x2(a[2], a[3]); o.z = pack_bf16x2(a[4], a[5]); o.w = pack_bf16x2(a[6], a[7]);
    stg_v4(reinterpret_cast<uint4*>(out + (size_t)warp * H) + v, o);
  }
}

// dY[row_k] = bf16(w_k) * dout[s] ; dw_k = <dout[s], y[row_k]>
__global__ void __launch_bounds__(256) moe_combine_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ y,
                                                             const int32_t* __restrict__ row, const float* __restrict__ w, int S, int H,
                                                             __nv_bfloat16* __restrict__ dy, float* __restrict__ dw) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= S) return;
  const int r1 = row[2 * warp], r2 = row[2 * warp + 1];
  const float w1 = bf16_round(w[2 * warp]), w2 = bf16_round(w[2 * warp + 1]);
  const int hv = H >> 3;
  float d1 = 0.f, d2 = 0.f;
  for (int v = lane; v < hv; v += 32) {
    uint4 g = ldg_nc_v4(reinterpret_cast<const uint4*>(dout + (size_t)warp * H) + v);
    float gf[8] = {bf16lo(g.x), bf16hi(g.x), bf16lo(g.y), bf16hi(g.y), bf16lo(g.z), bf16hi(g.z), bf16lo(g.w), bf16hi(g.w)};
    if (r1 >= 0) {
      uint4 u = ldg_nc_v4(reinterpret_cast<const uint4*>(y + (size_t)r1 * H) + v);
      float yf[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y), bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
      uint4 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) d1 = fmaf(gf[j], yf[j], d1);
      o.x = pack_bf16x2(w1 * gf[0], w1 * gf[1]); o.y = pack_bf16x2(w1 * gf[2], w1 * gf[3]);
      o.z = pack_bf16x2(w1 * gf[4], w1 * gf[5]); o.w = pack_bf16x2(w1 * gf[6], w1 * gf[7]);
      stg_v4(reinterpret_cast<uint4*>(dy + (size_t)r1 * H) + v, o);
    }
    if (r2 >= 0) {
      uint4 u = ldg_nc_v4(reinterpret_cast<const uint4*>(y + (size_t)r2 * H) + v);
      float yf[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y), bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
      uint4 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) d2 = fmaf(gf[j], yf[j], d2);
      o.x = pack_bf16x2(w2 * gf[0], w2 * gf[1]); o.y = pack_bf16x2(w2 * gf[2], w2 * gf[3]);
      o.z = pack_bf16x2(w2 * gf[4], w2 * gf[5]); o.w = pack_bf16x2(w2 * gf[6], w2 * gf[7]);
      stg_v4(reinterpret_cast<uint4*>(dy + (size_t)r2 * H) + v, o);
    }
  }
  d1 = warp_sum(d1); d2 = warp_sum(d2);
  if (lane == 0) { dw[2 * warp] = (r1 >= 0) ? d1 : 0.f; dw[2 * warp + 1] = (r2 >= 0) ? d2 : 0.f; }
}

// gate backward: (dw1, dw2, d l_aux) -> dlogits through renormalisation + softmax
__global__ void moe_gate_bwd_kernel(const float* __restrict__ gates, const int32_t* __restrict__ idx, const int32_t* __restrict__ row,
                                    const float* __restrict__ dw, const float* __restrict__ meta, const float* __restrict__ g_laux,
                                    int S, int E, float* __restrict__ dlogits) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float g[MAXE], dg[MAXE];
  const float gl = g_laux ? g_laux[0] : 0.f;
#pragma unroll
  for (int e = 0; e < MAXE; ++e) {
    g[e] = (e < E) ? gates[(size_t)s * E + e] : 0.f;
    // d l_aux / d gates[s,e] = E * ce[e] / S , ce[e] = exp_counts[e] / S
    dg[e] = (e < E) ? gl * (float)E * (meta[4 + e] / (float)S) / (float)S : 0.f;
  }
  const int e1 = idx[2 * s], e2 = idx[2 * s + 1];
  const bool k1 = row[2 * s] >= 0, k2 = row[2 * s + 1] >= 0;
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int e = 0; e < MAXE; ++e) { if (e == e1 && k1) a = g[e]; if (e == e2 && k2) b = g[e]; }
  const float sum = a + b;
  if (sum > FLT_EPSILON) {      // clamp(min=eps) inactive -> w1 = a/(a+b), w2 = b/(a+b)
    const float inv2 = 1.f / (sum * sum);
    const float da = (dw[2 * s] - dw[2 * s + 1]) * b * inv2, db = (dw[2 * s + 1] - dw[2 * s]) * a * inv2;
#pragma unroll
    for (int e = 0; e < MAXE; ++e) { if (e == e1 && k1) dg[e] += da; if (e == e2 && k2) dg[e] += db; }
  } else {                      // clamped: w = g / eps
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
      if (e == e1 && k1) dg[e] += dw[2 * s] / FLT_EPSILON;
      if (e == e2 && k2) dg[e] += dw[2 * s + 1] / FLT_EPSILON;
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int e = 0; e < MAXE; ++e) dot = fmaf(g[e], dg[e], dot);
#pragma unroll
  for (int e = 0; e < MAXE; ++e) if (e < E) dlogits[(size_t)s * E + e] = g[e] * (dg[e] - dot);
}

// dx[s] = dxp[row1] + dxp[row2] + sum_e dlogits[s,e] * wg[e,:] (+ dres[s])
__global__ void __launch_bounds__(256) moe_scatter_bwd_kernel(const __nv_bfloat16* __restrict__ dxp, const int32_t* __restrict__ row,
                                                             const float* __restrict__ dlogits, const float* __restrict__ wg,
                                                             const __nv_bfloat16* __restrict__ dres, int S, int H, int E,
                                                             __nv_bfloat16* __restrict__ dx) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= S) return;
  const int r1 = row[2 * warp], r2 = row[2 * warp + 1];
  float dl[MAXE];
#pragma unroll
  for (int e = 0; e < MAXE; ++e) dl[e] = (dlogits && e < E) ? dlogits[(size_t)warp * E + e] : 0.f;
  const int hv = H >> 3;
  for (int v = lane; v < hv; v += 32) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    if (r1 >= 0) {
      uint4 u = ldg_nc_v4(reinterpret_cast<const uint4*>(dxp + (size_t)r1 * H) + v);
      a[0] += bf16lo(u.x); a[1] += bf16hi(u.x); a[2] += bf16lo(u.y); a[3] += bf16hi(u.y);
      a[4] += bf16lo(u.z); a[5] += bf16hi(u.z); a[6] += bf16lo(u.w); a[7] += bf16hi(u.w);
    }
    if (r2 >= 0) {
      uint4 u = ldg_nc_v4(reinterpret_cast<const uint4*>(dxp + (size_t)r2 * H) + v);
      a[0] += bf16lo(u.x); a[1] += bf16hi(u.x); a[2] += bf16lo(u.y); a[3] += bf16hi(u.y);
      a[4] += bf16lo(u.z); a[5] += bf16hi(u.z); a[6] += bf16lo(u.w); a[7] += bf16hi(u.w);
    }
    if (dres) {
      uint4 u = ldg_nc_v4(reinterpret_cast<const uint4*>(dres + (size_t)warp * H) + v);
      a[0] += bf16lo(u.x); a[1] += bf16hi(u.x); a[2] += bf16lo(u.y); a[3] += bf16hi(u.y);
      a[4] += bf16lo(u.z); a[5] += bf16hi(u.z); a[6] += bf16lo(u.w); a[7] += bf16hi(u.w);
    }
    if (dlogits) {
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        if (e < E) {
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(wg + (size_t)e * H + v * 8));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(wg + (size_t)e * H + v * 8 + 4));
          a[0] = fmaf(dl[e], w0.x, a[0]); a[1] = fmaf(dl[e], w0.y, a[1]); a[2] = fmaf(dl[e], w0.z, a[2]); a[3] = fmaf(dl[e], w0.w, a[3]);
          a[4] = fmaf(dl[e], w1.x, a[4]); a[5] = fmaf(dl[e], w1.y, a[5]); a[6] = fmaf(dl[e], w1.z, a[6]); a[7] = fmaf(dl[e], w1.w, a[7]);
        }
      }
    }
    uint4 o;
    o.x = pack_bf16x2(a[0], a[1]); o.y = pack_bf16x2(a[2], a[3]); o.z = pack_bf16x2(a[4], a[5]); o.w = pack_bf16x2(a[6], a[7]);
    stg_v4(reinterpret_cast<uint4*>(dx + (size_t)warp * H) + v, o);
  }
}

constexpr int WG_SPLITS = 32;
// ws[split,e,h] = sum_{s in split} dlogits[s,e] * x[s,h]
__global__ void __launch_bounds__(256) moe_wg_grad_partial_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ dlogits,
                                                                 int S, int H, int E, float* __restrict__ ws) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  const int split = blockIdx.y;
  const int per = (S + WG_SPLITS - 1) / WG_SPLITS;
  const int s0 = split * per, s1 = min(S, s0 + per);
  if (h >= H) return;
  float acc[MAXE];
#pragma unroll
  for (int e = 0; e < MAXE; ++e) acc[e] = 0.f;
  for (int s = s0; s < s1; ++s) {
    const float xv = __bfloat162float(x[(size_t)s * H + h]);
#pragma unroll
    for (int e = 0; e < MAXE; ++e) if (e < E) acc[e] = fmaf(__ldg(dlogits + (size_t)s * E + e), xv, acc[e]);
  }
#pragma unroll
  for (int e = 0; e < MAXE; ++e) if (e < E) ws[((size_t)split * E + e) * H + h] = acc[e];
}
__global__ void moe_wg_grad_reduce_kernel(const float* __restrict__ ws, int H, int E, float* __restrict__ dwg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E * H) return;
  float a = 0.f;
  for (int sp = 0; sp < WG_SPLITS; ++sp) a += ws[(size_t)sp * E * H + i];
  dwg[i] += a;
}

}  // namespace

extern "C" int lmod_moe_capacity(int64_t S, int E, float capacity_factor, int64_t min_capacity) {
  // deepspeed _capacity: ceil(S/E * (cf*2)) ; computed in double like Python, raised to min_capacity
  double c = ceil(((double)S / (double)E) * ((double)capacity_factor * 2.0));
  int64_t ci = (int64_t)c;
  if (ci < min_capacity) ci = min_capacity;
  return (int)ci;
}

extern "C" int lmod_moe_route_scatter(const void* x, const float* wg, const float* noise, int64_t S, int64_t H, int E,
                                      float capacity_factor, int64_t min_capacity, int layout, float* logits, float* gates,
                                      int32_t* idx, int32_t* row, float* w, int32_t* offsets, float* meta, void* xp,
                                      int32_t* sync_ws, void* stream) {
  LMOD_CHECK_ARG(x && wg && noise && logits && gates && idx && row && w && offsets && meta && xp && sync_ws,
                 "lmod_moe_route_scatter: null pointer");
  LMOD_CHECK_ARG(E >= 2 && E <= MAXE, "lmod_moe_route_scatter: 2 <= E <= %d required (got %d)", MAXE, E);
  LMOD_CHECK_ARG(S > 0 && H > 0 && H % 8 == 0, "lmod_moe_route_scatter: H must be a multiple of 8");
  LMOD_CHECK_ARG(layout >= 0 && layout <= 2, "lmod_moe_route_scatter: layout must be 0 (compact), 1 (capacity slabs) or 2 (compact, 128-aligned)");
  RouteParams p;
  p.x = (const __nv_bfloat16*)x; p.wg = wg; p.noise = noise; p.S = (int)S; p.H = (int)H; p.E = E;
  p.capacity = lmod_moe_capacity(S, E, capacity_factor, min_capacity);
  p.layout = layout;
  p.logits = logits; p.gates = gates; p.idx = idx; p.row = row; p.w = w; p.offsets = offsets; p.meta = meta;
  p.xp = (__nv_bfloat16*)xp; p.sync_ws = (unsigned int*)sync_ws;

  int grid = lmod_num_sms();
  if (grid > (S + 3) / 4) grid = (int)((S + 3) / 4);
  if (grid < 1) grid = 1;
  const int tpb = (int)((S + grid - 1) / grid);
  size_t fixed = (size_t)E * H * 4 + (size_t)2 * S + 64;
  LMOD_CHECK_ARG(fixed < 180 * 1024, "lmod_moe_route_scatter: S*2 + E*H*4 exceeds the shared-memory budget");
  size_t budget = 200 * 1024 - fixed;
  int cap = (int)(budget / ((size_t)H * 2));
  if (cap > tpb) cap = tpb;
  p.stage_cap = cap;
  size_t smem = (size_t)E * H * 4 + (size_t)cap * H * 2 + (size_t)2 * S + 16;
  static bool attr_done = false;
  if (!attr_done) {
    LMOD_CUDA_OK(cudaFuncSetAttribute(moe_route_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    attr_done = true;
  }
  void* args[] = {(void*)&p};
  LMOD_CUDA_OK(cudaLaunchCooperativeKernel((const void*)moe_route_scatter_kernel, dim3(grid), dim3(RT_THREADS), args, smem,
                                           (cudaStream_t)stream));
  lmod_count_launch();
  return LMOD_OK;
}

extern "C" int lmod_moe_gather_combine(const void* y, const int32_t* row, const float* w, const void* residual, int64_t S,
                                       int64_t H, void* out, void* stream) {
  LMOD_CHECK_ARG(y && row && w && out && S > 0 && H % 8 == 0, "lmod_moe_gather_combine: bad arguments");
  unsigned blocks = (unsigned)((S * 32 + 255) / 256);
  moe_gather_combine_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)y, row, w, (const __nv_bfloat16*)residual,
                                                                      (int)S, (int)H, (__nv_bfloat16*)out);
  LMOD_LAUNCH_OK();
  return LMOD_OK;
}

extern "C" int lmod_moe_combine_bwd(const void* dout, const void* y, const int32_t* row, const float* w, int64_t S, int64_t H,
                                    void* dy, float* dw, void* stream) {
  LMOD_CHECK_ARG(dout && y && row && w && dy && dw && S > 0 && H % 8 == 0, "lmod_moe_combine_bwd: bad arguments");
  unsigned blocks = (unsigned)((S * 32 + 255) / 256);
  moe_combine_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dout, (const __nv_bfloat16*)y, row, w, (int)S,
                                                                   (int)H, (__nv_bfloat16*)dy, dw);
  LMOD_LAUNCH_OK();
  return LMOD_OK;
}

extern "C" int lmod_moe_gate_bwd(const float* gates, const int32_t* idx, const int32_t* row, const float* dw, const float* meta,
                                 const float* g_laux, int64_t S, int E, float* dlogits, void* stream) {
  LMOD_CHECK_ARG(gates && idx && row && dw && meta && dlogits && S > 0 && E >= 2 && E <= MAXE, "lmod_moe_gate_bwd: bad arguments");
  moe_gate_bwd_kernel<<<(unsigned)((S + 127) / 128), 128, 0, (cudaStream_t)stream>>>(gates, idx, row, dw, meta, g_laux, (int)S, E, dlogits);
  LMOD_LAUNCH_OK();
  return LMOD_OK;
}

extern "C" int lmod_moe_scatter_bwd(const void* dxp, const int32_t* row, const float* dlogits, const float* wg, const void* dres,
                                    int64_t S, int64_t H, int E, void* dx, void* stream) {
  LMOD_CHECK_ARG(dxp && row && dx && S > 0 && H % 8 == 0 && E <= MAXE, "lmod_moe_scatter_bwd: bad arguments");
  LMOD_CHECK_ARG(!dlogits || wg, "lmod_moe_scatter_bwd: wg required with dlogits");
  unsigned blocks = (unsigned)((S * 32 + 255) / 256);
  moe_scatter_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dxp, row, dlogits, wg, (const __nv_bfloat16*)dres,
                                                                   (int)S, (int)H, E, (__nv_bfloat16*)dx);
  LMOD_LAUNCH_OK();
  return LMOD_OK;
}

extern "C" int lmod_moe_wg_grad(const void* x, const float* dlogits, int64_t S, int64_t H, int E, float* ws, float* dwg, void* stream) {
  LMOD_CHECK_ARG(x && dlogits && ws && dwg && S > 0 && H > 0 && E <= MAXE, "lmod_moe_wg_grad: bad arguments");
  dim3 grid((unsigned)((H + 255) / 256), WG_SPLITS);
  moe_wg_grad_partial_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, dlogits, (int)S, (int)H, E, ws);
  LMOD_LAUNCH_OK();
  moe_wg_grad_reduce_kernel<<<(unsigned)((E * H + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ws, (int)H, E, dwg);
  LMOD_LAUNCH_OK();
  return LMOD_OK;
}
