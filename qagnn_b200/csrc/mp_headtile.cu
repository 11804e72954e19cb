// Shared-memory-tiled message passing for batches of small sub-graphs (the QA-GNN case: <= 200
// nodes per (question, choice) graph).  One launch per GATConvE layer does
//     logits -> per-SOURCE softmax -> out-degree rescale -> per-TARGET weighted sum
// (modeling/modeling_qagnn.py:442,455-484) with every gathered row served from shared memory.
//
// Why this shape (profiles/r1_microbench.txt, profiles/r1_v1_mp_ncu.md): each edge needs four row
// gathers (Kx[tgt], Ke[combo], Mx[src], Me[combo]); served from L2 they cap at ~11 G rows/s, so the
// per-head slice of the edge tables (C x d floats, 127 KB at C=612, d=50) has to live in shared
// memory next to the node tile of the current graph.  Hence:
//   * grid = H x floor(#SM / H) persistent CTAs; CTA (h, slot) owns head h of graphs slot, slot+S, ...
//   * phase 1: Ke_h resident, Kx_h tiles streamed through a 2-deep TMA (cp.async.bulk) ring ->
//     logits, softmax per source node, rescaled weights a'[e] written in BY-TARGET order (L2);
//   * phase 2: Me_h swapped in, Mx_h tiles streamed the same way -> aggr[:, h*d:(h+1)*d];
//   * ONE loader thread runs one graph ahead of the consumers: it issues the TMA bulk copies of the node tile and
//     of the graph's CSR slice (row pointers, degree-sorted node order, packed local ids, by-target positions /
//     phase-2 weights) into shared memory, so the consumers' inner loops touch no global memory (the v2 kernel
//     lost >50 % to L2 latency there);
//   * consumers: 8 lanes per node (quarter-warp), each lane owning float4 chunks l, l+8 of the padded
//     head row: a quarter-warp LDS.128 covers 128 contiguous bytes = one conflict-free wavefront;
//     packed FP32x2 math (FADD2/FFMA2, sm_100a); the dot product needs 3 shuffles; two edges per
//     iteration for ILP; the softmax runs lane-parallel over the node's edges after the loop.
// Node rows come from the head-major padded projection layout [3][H][N][DP] written by the
// projection GEMM, so a tile is one contiguous n*DP*4-byte bulk copy.
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace qagnn {

namespace {

struct HeadTileParams {
  int64_t N, Eps;  // Eps = per-head stride of score/alpha (E' rounded up to 4)
  int n, G, H, D, d, DP, C, S, W, ecap;
  const int32_t *rowptr_src, *rowptr_tgt, *pk_src, *pk_tgt, *tpos, *perm_src, *order_src, *order_tgt;
  const float *qkmh, *keh, *meh;
  float *score, *alpha, *aggr, *alpha_out;
  void *aggr_hi, *aggr_lo;  // optional split-bf16 planes of aggr [N, D] (A operand of the node-MLP GEMM)
  unsigned long long* trace;  // optional [4]: consumer-warp cycles {total, waiting for tiles, waiting for tables, warps}
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_chunked(char* dst, const char* src, uint32_t bytes, uint64_t* bar) {
  mbar_expect_tx(bar, bytes);
  const uint32_t kChunk = 32768;
  for (uint32_t o = 0; o < bytes; o += kChunk) bulk_g2s(dst + o, src + o, min(kChunk, bytes - o), bar);
}

// 16-byte shared-memory load from a 32-bit shared address (the V == 1 consumers add pre-scaled row offsets to it)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }

// shared-memory carve-up (bytes from the start of dynamic smem)
struct SmemMap {
  uint32_t tab, tile0, tile_bytes, rp0, rp_bytes, od0, od_bytes, ia0, ib0, idx_bytes, bars, meta;
};
__host__ __device__ inline SmemMap make_smem_map(int C, int DP, int n, int ecap) {
  SmemMap m;
  m.tab = 0;
  m.tile0 = (uint32_t)C * DP * 4;
  m.tile_bytes = (uint32_t)n * DP * 4;
  m.rp0 = m.tile0 + 2 * m.tile_bytes;
  m.rp_bytes = (uint32_t)((n + 1 + 3 + 3) / 4 * 4) * 4;  // + alignment slack of the slice start
  m.od0 = m.rp0 + 2 * m.rp_bytes;
  m.od_bytes = (uint32_t)((n + 3 + 3) / 4 * 4) * 4;  // degree-sorted local node ids (+ alignment slack)
  m.ia0 = m.od0 + 2 * m.od_bytes;
  m.idx_bytes = (uint32_t)ecap * 4;
  m.ib0 = m.ia0 + 2 * m.idx_bytes;
  m.bars = m.ib0 + 2 * m.idx_bytes;
  m.meta = m.bars + 5 * 8;
  return m;
}
inline size_t smem_total(const SmemMap& m) { return (size_t)m.meta + 4 * 4 + 16; }

// V selects the consumer code: 0 = the round-1 kernel (full GPU suite, sanitizers: the default); 1 = the round-2
// candidate (QAGNN_MP_VARIANT=1; bit-identical results on the B200 and 2 % faster, profiles/r1_mp_stall_breakdown.md —
// the kernel is bound by the LDS/SHFL pipe, not by instruction issue); 2 = 1 + the proxy fence only before the phase
// switch (compiled, not yet run).  What V >= 1 changes:
//   * serpentine quad->warp assignment: with degree-sorted quads, (warp, warp+W) gives warp 0 the two heaviest
//     quads of each half (critical path 1.32x the mean on the cfg2 batch), (warp, 2W-1-warp) gives 1.13x;
//   * the degree-order entry of the NEXT-next graph is fetched one iteration early, so the Q-row prefetch no longer
//     waits on a dependent global load at the top of every graph;
//   * row offsets are pre-multiplied once per edge by the lane that loads the packed ids (one LEA per row chunk
//     in the edge loop instead of two IMADs);
//   * phase 1 reduces the 8 lane-partials of 8 edges with a 7-shuffle transposing tree (lane j ends up with the
//     logit of edge j, the same summation tree as the butterfly) instead of 3 shuffles per edge.
template <int CPL, int QPW, int V>  // float4 chunks per lane; node-quads per consumer warp and graph (1..4); variant
__global__ void __launch_bounds__(QPW == 1 ? 1024 : QPW == 2 ? 832 : QPW == 3 ? 576 : 448, 1) mp_headtile_kernel(const HeadTileParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int NCH = p.DP / 4;
  const SmemMap sm = make_smem_map(p.C, p.DP, p.n, p.ecap);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + sm.bars);
  uint64_t* full = bars;        // [2]  one expect_tx arrival covering tile + CSR slice bytes
  uint64_t* empty = bars + 2;   // [2]  W consumer arrivals
  uint64_t* tabbar = bars + 4;  // [1]

  const int h = blockIdx.x % p.H;
  const int slot = blockIdx.x / p.H;
  const int Gc = slot < p.G ? (p.G - slot + p.S - 1) / p.S : 0;  // graphs slot, slot+S, ...
  if (Gc == 0) return;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&empty[0], p.W);
    mbar_init(&empty[1], p.W);
    mbar_init(tabbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const size_t head_rows = (size_t)p.N * p.DP;  // floats per [N, DP] slab
  const float* Qh = p.qkmh + (size_t)(0 * p.H + h) * head_rows;
  const float* Kh = p.qkmh + (size_t)(1 * p.H + h) * head_rows;
  const float* Mh = p.qkmh + (size_t)(2 * p.H + h) * head_rows;
  const size_t hE = (size_t)h * p.Eps;

  if (warp == p.W) {
    // ===== loader: ONE thread drives the TMA ring — node tile + CSR slice (row pointers, packed ids, =====
    // ===== by-target positions / phase-2 weights) of the next graph, all as cp.async.bulk copies     =====
    if (lane != 0) return;
    bulk_g2s_chunked((char*)(smem_raw + sm.tab), (const char*)(p.keh + (size_t)h * p.C * p.DP), sm.tile0, tabbar);
    int nb = p.rowptr_src[(size_t)slot * p.n], ne = p.rowptr_src[(size_t)slot * p.n + p.n];  // bounds of graph 0
    for (int t = 0; t < 2 * Gc; ++t) {
      const int b = t & 1;
      const bool ph2 = t >= Gc;
      if (t == Gc) {
        // phase switch: every consumer has left phase 1 once the last two tiles are released
        mbar_wait(&empty[(Gc - 1) & 1], ((Gc - 1) >> 1) & 1);
        if (Gc >= 2) mbar_wait(&empty[(Gc - 2) & 1], ((Gc - 2) >> 1) & 1);
        bulk_g2s_chunked((char*)(smem_raw + sm.tab), (const char*)(p.meh + (size_t)h * p.C * p.DP), sm.tile0, tabbar);
      }
      if (t >= 2) mbar_wait(&empty[b], ((t >> 1) - 1) & 1);
      const int g = slot + (ph2 ? t - Gc : t) * p.S;
      const int64_t v0 = (int64_t)g * p.n;
      const int32_t* rowptr = ph2 ? p.rowptr_tgt : p.rowptr_src;
      const int base = nb & ~3, cnt = ne - base;
      const bool staged = cnt <= p.ecap;
      const uint32_t idx_bytes = staged ? (uint32_t)((cnt + 3) & ~3) * 4u : 0u;
      const int64_t rp_base = v0 & ~(int64_t)3;
      const uint32_t rp_bytes = (uint32_t)(((v0 - rp_base) + p.n + 1 + 3) & ~3) * 4u;
      mbar_expect_tx(&full[b], sm.tile_bytes + rp_bytes + (uint32_t)(((v0 - rp_base) + p.n + 3) & ~3) * 4u + 2 * idx_bytes);
      {
        const char* src = (const char*)((ph2 ? Mh : Kh) + (size_t)v0 * p.DP);
        char* dst = (char*)(smem_raw + sm.tile0 + b * sm.tile_bytes);
        for (uint32_t o = 0; o < sm.tile_bytes; o += 32768) bulk_g2s(dst + o, src + o, min(32768u, sm.tile_bytes - o), &full[b]);
      }
      bulk_g2s(smem_raw + sm.rp0 + b * sm.rp_bytes, rowptr + rp_base, rp_bytes, &full[b]);
      const uint32_t od_bytes = (uint32_t)(((v0 - rp_base) + p.n + 3) & ~3) * 4u;
      bulk_g2s(smem_raw + sm.od0 + b * sm.od_bytes, (ph2 ? p.order_tgt : p.order_src) + rp_base, od_bytes, &full[b]);
      if (staged && idx_bytes) {
        bulk_g2s(smem_raw + sm.ia0 + b * sm.idx_bytes, (ph2 ? p.pk_tgt : p.pk_src) + base, idx_bytes, &full[b]);
        if (ph2) bulk_g2s(smem_raw + sm.ib0 + b * sm.idx_bytes, p.alpha + hE + base, idx_bytes, &full[b]);
        else bulk_g2s(smem_raw + sm.ib0 + b * sm.idx_bytes, p.tpos + base, idx_bytes, &full[b]);
      }
      if (t + 1 < 2 * Gc) {  // bounds of the next graph, in flight while the consumers work
        const bool nph2 = t + 1 >= Gc;
        const int64_t nv0 = (int64_t)(slot + (nph2 ? t + 1 - Gc : t + 1) * p.S) * p.n;
        const int32_t* nrp = nph2 ? p.rowptr_tgt : p.rowptr_src;
        nb = nrp[nv0];
        ne = nrp[nv0 + p.n];
      }
    }
    return;
  }

  // ========================= consumer warps: 4 nodes per warp, 8 lanes per node =========================
  const int l8 = lane & 7, qbase = lane & 24, qi = lane >> 3;
  const int nquads = (p.n + 3) / 4;
  int chunk[CPL];
  bool cvalid[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    cvalid[k] = (l8 + 8 * k) < NCH;
    // idle lanes (chunk slot past the row) re-read an in-row chunk against q = 0, so the product is 0 * finite.
    // It must stay INSIDE the row: the bytes after a row's last chunk belong to the next row / the next smem region
    // and 0 * NaN would poison the shuffled sum (found with compute-sanitizer, whose smem fill is not finite).
    // For CPL == 2 the slot is l8 + 8 >= NCH > 8, so l8 % NCH == l8: bytes 16*l8.. do not share banks with chunks 8..
    chunk[k] = cvalid[k] ? l8 + 8 * k : l8 % NCH;
  }
  const float4* tab = reinterpret_cast<const float4*>(smem_raw + sm.tab);

  auto load_q = [&](int g, int quad, float4 (&q)[CPL]) {
    const int slot_i = quad * 4 + qi;  // position in the degree-sorted order of graph g
    const int vl = p.order_src[(int64_t)g * p.n + (slot_i < p.n ? slot_i : 0)];
    const int64_t v = (int64_t)g * p.n + vl;
#pragma unroll
    for (int k = 0; k < CPL; ++k)
      q[k] = (cvalid[k] && quad < nquads) ? __ldg(reinterpret_cast<const float4*>(Qh + v * p.DP) + chunk[k])
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
  };

  // V == 1 helpers: quad owned by this warp in pass u (serpentine), the degree-order entry and the Q rows apart
  auto quad_of = [&](int u) { return (u & 1) ? (u + 1) * p.W - 1 - warp : warp + u * p.W; };
  auto load_order = [&](int g, int quad) {
    const int slot_i = quad * 4 + qi;
    return p.order_src[(int64_t)g * p.n + (slot_i < p.n ? slot_i : 0)];
  };
  auto load_q_at = [&](int g, int quad, int vl, float4 (&q)[CPL]) {
    const int64_t v = (int64_t)g * p.n + vl;
#pragma unroll
    for (int k = 0; k < CPL; ++k)
      q[k] = (cvalid[k] && quad < nquads) ? __ldg(reinterpret_cast<const float4*>(Qh + v * p.DP) + chunk[k])
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  const float4* tabc[CPL];  // table base + this lane's chunk: a row is then one scaled add away
#pragma unroll
  for (int k = 0; k < CPL; ++k) tabc[k] = tab + chunk[k];

  long long t_begin = 0, t_wait_tile = 0, t_wait_tab = 0;
  if (p.trace != nullptr) t_begin = clock64();
  // ---------------------------------- phase 1: attention weights ----------------------------------
  float4 qn[QPW][CPL];  // Q rows of this warp's quads, prefetched one graph ahead
  int vln[QPW];         // V == 1: degree-order entries of the graph whose Q rows are fetched next
  if constexpr (V >= 1) {
#pragma unroll
    for (int u = 0; u < QPW; ++u) load_q_at(slot, quad_of(u), load_order(slot, quad_of(u)), qn[u]);
#pragma unroll
    for (int u = 0; u < QPW; ++u) vln[u] = Gc > 1 ? load_order(slot + p.S, quad_of(u)) : 0;
  } else {
#pragma unroll
    for (int u = 0; u < QPW; ++u) load_q(slot, warp + u * p.W, qn[u]);
  }
  {
    const long long c0 = p.trace ? clock64() : 0;
    mbar_wait(tabbar, 0);
    if (p.trace) t_wait_tab += clock64() - c0;
  }
  for (int t = 0; t < Gc; ++t) {
    const int b = t & 1;
    const int g = slot + t * p.S;
    float4 qc[QPW][CPL];
#pragma unroll
    for (int u = 0; u < QPW; ++u)
#pragma unroll
      for (int k = 0; k < CPL; ++k) qc[u][k] = qn[u][k];
    if constexpr (V >= 1) {
      if (t + 1 < Gc) {
#pragma unroll
        for (int u = 0; u < QPW; ++u) load_q_at(g + p.S, quad_of(u), vln[u], qn[u]);
        if (t + 2 < Gc) {
#pragma unroll
          for (int u = 0; u < QPW; ++u) vln[u] = load_order(g + 2 * p.S, quad_of(u));
        }
      }
    } else {
      if (t + 1 < Gc) {
#pragma unroll
        for (int u = 0; u < QPW; ++u) load_q(g + p.S, warp + u * p.W, qn[u]);
      }
    }
    {
      const long long c0 = p.trace ? clock64() : 0;
      mbar_wait(&full[b], (t >> 1) & 1);
      if (p.trace) t_wait_tile += clock64() - c0;
    }
    const float4* kt = reinterpret_cast<const float4*>(smem_raw + sm.tile0 + b * sm.tile_bytes);
    const int* rp = reinterpret_cast<const int*>(smem_raw + sm.rp0 + b * sm.rp_bytes) + (int)(((int64_t)g * p.n) & 3);
    const int* od = reinterpret_cast<const int*>(smem_raw + sm.od0 + b * sm.od_bytes) + (int)(((int64_t)g * p.n) & 3);
    const int* ia = reinterpret_cast<const int*>(smem_raw + sm.ia0 + b * sm.idx_bytes);
    const int* ib = reinterpret_cast<const int*>(smem_raw + sm.ib0 + b * sm.idx_bytes);
    const int base = rp[0] & ~3;
    const bool staged = rp[p.n] - base <= p.ecap;
#pragma unroll
    for (int u = 0; u < QPW; ++u) {
      const int quad = (V >= 1) ? quad_of(u) : warp + u * p.W;
      if (quad >= nquads) {
        if constexpr (V >= 1) continue; else break;
      }
      const bool nvalid = quad * 4 + qi < p.n;
      const int vl = nvalid ? od[quad * 4 + qi] : 0;  // 4 nodes of similar out-degree per warp
      const int begr = rp[vl] - base;
      const int deg = nvalid ? rp[vl + 1] - base - begr : 0;
      int maxdeg = max(deg, __shfl_xor_sync(0xffffffffu, deg, 8));
      maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, 16));
      float skeep = -INFINITY;  // lane j keeps the logit of edge j (j < 8)
      if constexpr (V >= 1) {
        uint32_t ktc[CPL];  // shared address of tile base + this lane's chunk: a row is one add away
#pragma unroll
        for (int k = 0; k < CPL; ++k) ktc[k] = smem_u32(kt + chunk[k]);
        const bool b4 = (l8 & 4) != 0, b2 = (l8 & 2) != 0, b1 = (l8 & 1) != 0;
        for (int i0 = 0; i0 < maxdeg; i0 += 8) {
          uint32_t pko = 0;  // (tile row offset in bytes << 16) | table row offset in float4 units
          if (i0 + l8 < deg) {
            const uint32_t pkv = (uint32_t)(staged ? ia[begr + i0 + l8] : p.pk_src[base + begr + i0 + l8]);
            pko = (((pkv >> 16) * (uint32_t)(NCH * 16)) << 16) | ((pkv & 0xffffu) * (uint32_t)NCH);
          }
          const int lim = min(8, maxdeg - i0);
          float v[8];  // this lane's partial dot products of the block's 8 edges
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = 0.f;
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            if (j < lim) {  // warp-uniform
              const uint32_t w0 = __shfl_sync(0xffffffffu, pko, qbase + j);
              const uint32_t w1 = __shfl_sync(0xffffffffu, pko, qbase + j + 1);
              float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
              for (int k = 0; k < CPL; ++k) {
                const float4 x0 = lds128(ktc[k] + (w0 >> 16)), y0 = tabc[k][w0 & 0xffffu];
                const float4 x1 = lds128(ktc[k] + (w1 >> 16)), y1 = tabc[k][w1 & 0xffffu];
                a0 = __ffma2_rn(__fadd2_rn(lo2(x0), lo2(y0)), lo2(qc[u][k]), a0);
                a0 = __ffma2_rn(__fadd2_rn(hi2(x0), hi2(y0)), hi2(qc[u][k]), a0);
                a1 = __ffma2_rn(__fadd2_rn(lo2(x1), lo2(y1)), lo2(qc[u][k]), a1);
                a1 = __ffma2_rn(__fadd2_rn(hi2(x1), hi2(y1)), hi2(qc[u][k]), a1);
              }
              v[j] = a0.x + a0.y;
              v[j + 1] = a1.x + a1.y;
            }
          }
          // transposing reduction over the 8 lanes of the node: after the three stages lane j holds
          // ((v_j[l] + v_j[l^4]) + (v_j[l^2] + v_j[l^6])) + (...[l^1]...), l = j — the butterfly's tree at lane j
          float r4[4], r2[2];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float keep = b4 ? v[i + 4] : v[i], send = b4 ? v[i] : v[i + 4];
            r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float keep = b2 ? r4[i + 2] : r4[i], send = b2 ? r4[i] : r4[i + 2];
            r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
          }
          const float keep = b1 ? r2[1] : r2[0], send = b1 ? r2[0] : r2[1];
          const float sj = keep + __shfl_xor_sync(0xffffffffu, send, 1);
          if (i0 == 0) skeep = sj;
          else if (i0 + l8 < deg) p.score[hE + base + begr + i0 + l8] = sj;  // hub nodes: logits past the 8th edge
        }
      } else {
        for (int i0 = 0; i0 < maxdeg; i0 += 8) {
          int pkv = 0;
          if (i0 + l8 < deg) pkv = staged ? ia[begr + i0 + l8] : p.pk_src[base + begr + i0 + l8];
          const int lim = min(8, maxdeg - i0);
          for (int j = 0; j < lim; j += 2) {
            const uint32_t w0 = (uint32_t)__shfl_sync(0xffffffffu, pkv, qbase + j);
            const uint32_t w1 = (uint32_t)__shfl_sync(0xffffffffu, pkv, qbase + j + 1);
            const float4* k0 = kt + (w0 >> 16) * NCH;
            const float4* e0 = tab + (w0 & 0xffffu) * NCH;
            const float4* k1 = kt + (w1 >> 16) * NCH;
            const float4* e1 = tab + (w1 & 0xffffu) * NCH;
            float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
              // (quarter-warps past their node's degree read row 0 against a discarded result; predicating these loads
              //  off saves shared-memory wavefronts but cost 143 -> 177 us in issue slots: measured, reverted)
              const float4 x0 = k0[chunk[k]], y0 = e0[chunk[k]], x1 = k1[chunk[k]], y1 = e1[chunk[k]];
              a0 = __ffma2_rn(__fadd2_rn(lo2(x0), lo2(y0)), lo2(qc[u][k]), a0);
              a0 = __ffma2_rn(__fadd2_rn(hi2(x0), hi2(y0)), hi2(qc[u][k]), a0);
              a1 = __ffma2_rn(__fadd2_rn(lo2(x1), lo2(y1)), lo2(qc[u][k]), a1);
              a1 = __ffma2_rn(__fadd2_rn(hi2(x1), hi2(y1)), hi2(qc[u][k]), a1);
            }
            float s0 = a0.x + a0.y, s1 = a1.x + a1.y;
            s0 += __shfl_xor_sync(0xffffffffu, s0, 4);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 4);
            s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
            s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
            if (i0 == 0) {
              skeep = (l8 == j) ? s0 : skeep;
              skeep = (l8 == j + 1) ? s1 : skeep;
            } else {  // hub nodes (degree > 8): spill the logits to the L2 scratch
              if (l8 == 0 && i0 + j < deg) p.score[hE + base + begr + i0 + j] = s0;
              if (l8 == 1 && i0 + j + 1 < deg) p.score[hE + base + begr + i0 + j + 1] = s1;
            }
          }
        }
      }
      // softmax over this node's out-edges, lane-parallel (lane j <-> edge j, j + 8, ...)
      const bool hub = maxdeg > 8;  // warp-uniform; logits beyond the 8th edge live in the L2 scratch
      if (hub) __syncwarp();
      float m = (l8 < deg) ? skeep : -INFINITY;
      if (hub)
        for (int j = 8 + l8; j < deg; j += 8) m = fmaxf(m, p.score[hE + base + begr + j]);
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      const float ex0 = (l8 < deg) ? __expf(skeep - m) : 0.f;
      float ssum = ex0;
      if (hub)
        for (int j = 8 + l8; j < deg; j += 8) ssum += __expf(p.score[hE + base + begr + j] - m);
      ssum += __shfl_xor_sync(0xffffffffu, ssum, 4);
      ssum += __shfl_xor_sync(0xffffffffu, ssum, 2);
      ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);
      // a = ex / (sum + 1e-16) (torch_geometric.utils.softmax), then * out-degree of the source (:476-481)
      const float rden = __fdividef(1.f, ssum + 1e-16f);
      const float degf = (float)deg;
      if (l8 < deg) {
        const float a = ex0 * rden;
        const int tp = staged ? ib[begr + l8] : p.tpos[base + begr + l8];
        p.alpha[hE + tp] = a * degf;  // stored in by-target order for phase 2
        if (p.alpha_out != nullptr) p.alpha_out[(size_t)p.perm_src[base + begr + l8] * p.H + h] = a;
      }
      if (hub) {
        for (int j = 8 + l8; j < deg; j += 8) {
          const float a = __expf(p.score[hE + base + begr + j] - m) * rden;
          const int tp = staged ? ib[begr + j] : p.tpos[base + begr + j];
          p.alpha[hE + tp] = a * degf;
          if (p.alpha_out != nullptr) p.alpha_out[(size_t)p.perm_src[base + begr + j] * p.H + h] = a;
        }
      }
    }
    // a'[e] is read back by cp.async.bulk (async proxy) in phase 2.  The loader switches phase after the arrivals of
    // the last two graphs, so one fence before each of those covers all of this thread's earlier stores (V >= 2);
    // per graph the fence was 5 % of the warp stall samples (profiles/r1_mp_stall_breakdown.md).
    if (V < 2 || t >= Gc - 2) {
      __threadfence_block();
      asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[b]);
  }

  // ---------------------------------- phase 2: weighted sum by target ----------------------------------
  {
    const long long c0 = p.trace ? clock64() : 0;
    mbar_wait(tabbar, 1);
    if (p.trace) t_wait_tab += clock64() - c0;
  }
  for (int t = Gc; t < 2 * Gc; ++t) {
    const int b = t & 1;
    const int g = slot + (t - Gc) * p.S;
    {
      const long long c0 = p.trace ? clock64() : 0;
      mbar_wait(&full[b], (t >> 1) & 1);
      if (p.trace) t_wait_tile += clock64() - c0;
    }
    const float4* mt = reinterpret_cast<const float4*>(smem_raw + sm.tile0 + b * sm.tile_bytes);
    const int* rp = reinterpret_cast<const int*>(smem_raw + sm.rp0 + b * sm.rp_bytes) + (int)(((int64_t)g * p.n) & 3);
    const int* od = reinterpret_cast<const int*>(smem_raw + sm.od0 + b * sm.od_bytes) + (int)(((int64_t)g * p.n) & 3);
    const int* ia = reinterpret_cast<const int*>(smem_raw + sm.ia0 + b * sm.idx_bytes);
    const float* ib = reinterpret_cast<const float*>(smem_raw + sm.ib0 + b * sm.idx_bytes);
    const int base = rp[0] & ~3;
    const bool staged = rp[p.n] - base <= p.ecap;
#pragma unroll
    for (int u = 0; u < QPW; ++u) {
      const int quad = (V >= 1) ? quad_of(u) : warp + u * p.W;
      if (quad >= nquads) {
        if constexpr (V >= 1) continue; else break;
      }
      const bool nvalid = quad * 4 + qi < p.n;
      const int vl = nvalid ? od[quad * 4 + qi] : 0;  // 4 nodes of similar in-degree per warp
      const int begr = rp[vl] - base;
      const int deg = nvalid ? rp[vl + 1] - base - begr : 0;
      int maxdeg = max(deg, __shfl_xor_sync(0xffffffffu, deg, 8));
      maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, 16));
      float2 acc[CPL][2];
#pragma unroll
      for (int k = 0; k < CPL; ++k) acc[k][0] = acc[k][1] = make_float2(0.f, 0.f);
      if constexpr (V >= 1) {
        uint32_t mtc[CPL];  // shared address of tile base + this lane's chunk
#pragma unroll
        for (int k = 0; k < CPL; ++k) mtc[k] = smem_u32(mt + chunk[k]);
        for (int i0 = 0; i0 < maxdeg; i0 += 8) {
          uint32_t pko = 0;  // (tile row offset in bytes << 16) | table row offset in float4 units
          float wv = 0.f;
          if (i0 + l8 < deg) {
            const uint32_t pkv = (uint32_t)(staged ? ia[begr + i0 + l8] : p.pk_tgt[base + begr + i0 + l8]);
            pko = (((pkv >> 16) * (uint32_t)(NCH * 16)) << 16) | ((pkv & 0xffffu) * (uint32_t)NCH);
            wv = staged ? ib[begr + i0 + l8] : p.alpha[hE + base + begr + i0 + l8];
          }
          const int lim = min(8, maxdeg - i0);
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            if (j < lim) {  // warp-uniform
              const uint32_t w0 = __shfl_sync(0xffffffffu, pko, qbase + j);
              const uint32_t w1 = __shfl_sync(0xffffffffu, pko, qbase + j + 1);
              const float a0 = __shfl_sync(0xffffffffu, wv, qbase + j);      // 0 beyond this node's degree
              const float a1 = __shfl_sync(0xffffffffu, wv, qbase + j + 1);
              const float2 aa0 = make_float2(a0, a0), aa1 = make_float2(a1, a1);
#pragma unroll
              for (int k = 0; k < CPL; ++k) {
                const float4 x0 = lds128(mtc[k] + (w0 >> 16)), y0 = tabc[k][w0 & 0xffffu];
                const float4 x1 = lds128(mtc[k] + (w1 >> 16)), y1 = tabc[k][w1 & 0xffffu];
                acc[k][0] = __ffma2_rn(__fadd2_rn(lo2(x0), lo2(y0)), aa0, acc[k][0]);
                acc[k][1] = __ffma2_rn(__fadd2_rn(hi2(x0), hi2(y0)), aa0, acc[k][1]);
                acc[k][0] = __ffma2_rn(__fadd2_rn(lo2(x1), lo2(y1)), aa1, acc[k][0]);
                acc[k][1] = __ffma2_rn(__fadd2_rn(hi2(x1), hi2(y1)), aa1, acc[k][1]);
              }
            }
          }
        }
      } else {
        for (int i0 = 0; i0 < maxdeg; i0 += 8) {
          int pkv = 0;
          float wv = 0.f;
          if (i0 + l8 < deg) {
            pkv = staged ? ia[begr + i0 + l8] : p.pk_tgt[base + begr + i0 + l8];
            wv = staged ? ib[begr + i0 + l8] : p.alpha[hE + base + begr + i0 + l8];
          }
          const int lim = min(8, maxdeg - i0);
          for (int j = 0; j < lim; j += 2) {
            const uint32_t w0 = (uint32_t)__shfl_sync(0xffffffffu, pkv, qbase + j);
            const uint32_t w1 = (uint32_t)__shfl_sync(0xffffffffu, pkv, qbase + j + 1);
            const float a0 = __shfl_sync(0xffffffffu, wv, qbase + j);      // 0 beyond this node's degree
            const float a1 = __shfl_sync(0xffffffffu, wv, qbase + j + 1);
            const float4* m0 = mt + (w0 >> 16) * NCH;
            const float4* e0 = tab + (w0 & 0xffffu) * NCH;
            const float4* m1 = mt + (w1 >> 16) * NCH;
            const float4* e1 = tab + (w1 & 0xffffu) * NCH;
            const float2 aa0 = make_float2(a0, a0), aa1 = make_float2(a1, a1);
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
              const float4 x0 = m0[chunk[k]], y0 = e0[chunk[k]], x1 = m1[chunk[k]], y1 = e1[chunk[k]];
              acc[k][0] = __ffma2_rn(__fadd2_rn(lo2(x0), lo2(y0)), aa0, acc[k][0]);
              acc[k][1] = __ffma2_rn(__fadd2_rn(hi2(x0), hi2(y0)), aa0, acc[k][1]);
              acc[k][0] = __ffma2_rn(__fadd2_rn(lo2(x1), lo2(y1)), aa1, acc[k][0]);
              acc[k][1] = __ffma2_rn(__fadd2_rn(hi2(x1), hi2(y1)), aa1, acc[k][1]);
            }
          }
        }
      }
      if (nvalid) {
        const size_t obase = ((int64_t)g * p.n + vl) * p.D + (size_t)h * p.d;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const int c0 = 4 * (l8 + 8 * k);
          if (!cvalid[k] || c0 >= p.d) continue;
          const float v0 = acc[k][0].x, v1 = acc[k][0].y, v2 = acc[k][1].x, v3 = acc[k][1].y;
          if (p.aggr != nullptr) {
            float* out = p.aggr + obase;
            if ((p.d & 3) == 0 && (p.D & 3) == 0) {
              *reinterpret_cast<float4*>(out + c0) = make_float4(v0, v1, v2, v3);
            } else if ((p.d & 1) == 0) {
              *reinterpret_cast<float2*>(out + c0) = make_float2(v0, v1);
              if (c0 + 2 < p.d) *reinterpret_cast<float2*>(out + c0 + 2) = make_float2(v2, v3);
            } else {
              out[c0] = v0;
              if (c0 + 1 < p.d) out[c0 + 1] = v1;
              if (c0 + 2 < p.d) out[c0 + 2] = v2;
              if (c0 + 3 < p.d) out[c0 + 3] = v3;
            }
          }
          if (p.aggr_hi != nullptr) {  // d even (D % 8 == 0 on this path): bf16x2 pairs stay 4-byte aligned
            __nv_bfloat16* oh = (__nv_bfloat16*)p.aggr_hi + obase + c0;
            __nv_bfloat16* ol = (__nv_bfloat16*)p.aggr_lo + obase + c0;
            __nv_bfloat162 h01, l01, h23, l23;
            h01.x = __float2bfloat16_rn(v0); h01.y = __float2bfloat16_rn(v1);
            l01.x = __float2bfloat16_rn(v0 - __bfloat162float(h01.x)); l01.y = __float2bfloat16_rn(v1 - __bfloat162float(h01.y));
            *reinterpret_cast<__nv_bfloat162*>(oh) = h01;
            *reinterpret_cast<__nv_bfloat162*>(ol) = l01;
            if (c0 + 2 < p.d) {
              h23.x = __float2bfloat16_rn(v2); h23.y = __float2bfloat16_rn(v3);
              l23.x = __float2bfloat16_rn(v2 - __bfloat162float(h23.x)); l23.y = __float2bfloat16_rn(v3 - __bfloat162float(h23.y));
              *reinterpret_cast<__nv_bfloat162*>(oh + 2) = h23;
              *reinterpret_cast<__nv_bfloat162*>(ol + 2) = l23;
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[b]);
  }
  if (p.trace != nullptr && lane == 0) {
    atomicAdd(p.trace + 0, (unsigned long long)(clock64() - t_begin));
    atomicAdd(p.trace + 1, (unsigned long long)t_wait_tile);
    atomicAdd(p.trace + 2, (unsigned long long)t_wait_tab);
    atomicAdd(p.trace + 3, 1ull);
  }
}

__global__ void zero_head_pads_kernel(int64_t rows, int d, int DP, float* __restrict__ buf) {
  const int pad = DP - d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows * pad; i += (int64_t)gridDim.x * blockDim.x)
    buf[(i / pad) * DP + d + (i % pad)] = 0.f;
}

struct HeadTilePlan {
  bool ok;
  int DP, C, S, W, cpl, qpw, sms, ecap;
  size_t smem;
};

HeadTilePlan make_plan(const qagnn_shape& s) {
  HeadTilePlan pl{};
  pl.ok = false;
  if (s.n_per_graph <= 0 || s.n_per_graph > 65535 || s.N % s.n_per_graph != 0) return pl;
  const int d = s.D / s.H;
  pl.DP = head_dim_padded(d);
  pl.C = s.R * s.T * s.T + s.T;
  if (pl.C > 65536 || pl.DP > 64) return pl;
  pl.cpl = pl.DP <= 32 ? 1 : 2;
  static int sms_c[kMaxDevices] = {0}, smem_c[kMaxDevices] = {0};
  const int dev = current_device();
  if (sms_c[dev] == 0) {
    cudaDeviceGetAttribute(&smem_c[dev], cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&sms_c[dev], cudaDevAttrMultiProcessorCount, dev);
  }
  const int sms = sms_c[dev], max_smem = smem_c[dev];
  if (sms <= 0) return pl;
  pl.sms = sms;
  if (s.H > sms) return pl;
  // index staging capacity: whatever is left after the table and the two tiles, at least a few edges per node
  const SmemMap m0 = make_smem_map(pl.C, pl.DP, s.n_per_graph, 0);
  const long fixed = (long)smem_total(m0);
  long ecap = ((long)max_smem - fixed) / 16;  // 2 buffers x (ia + ib) x 4 bytes
  ecap = ecap / 4 * 4;
  if (ecap < 2 * (long)s.n_per_graph + 8) return pl;
  if (ecap > 65536) ecap = 65536;
  pl.ecap = (int)ecap;
  pl.smem = smem_total(make_smem_map(pl.C, pl.DP, s.n_per_graph, pl.ecap));
  pl.S = sms / s.H;
  const int quads = (s.n_per_graph + 3) / 4;
  int passes = (quads + 30) / 31;  // at most 31 consumer warps + 1 loader warp
  if (passes > 2) return pl;
  static const int forced_qpw = [] { const char* e = getenv("QAGNN_MP_QPW"); return e ? atoi(e) : 0; }();
  if (forced_qpw >= passes && forced_qpw <= 4) passes = forced_qpw;  // fewer, fatter warps (more registers each)
  pl.qpw = passes;
  pl.W = (quads + passes - 1) / passes;
  const int max_w = passes == 1 ? 31 : passes == 2 ? 25 : passes == 3 ? 17 : 13;  // launch bounds of the variants
  if (pl.W > max_w) return pl;
  pl.ok = true;
  return pl;
}

template <int CPL, int QPW, int V>
int32_t launch_t(const HeadTileParams& p, const HeadTilePlan& plan, unsigned grid, unsigned block, cudaStream_t st) {
  static size_t attr_smem[kMaxDevices] = {0};  // the attribute is per device
  const int dev = current_device();
  if (plan.smem > attr_smem[dev]) {
    QAGNN_CHECK_CUDA(cudaFuncSetAttribute(mp_headtile_kernel<CPL, QPW, V>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)plan.smem));
    attr_smem[dev] = plan.smem;
  }
  mp_headtile_kernel<CPL, QPW, V><<<grid, block, plan.smem, st>>>(p);
  QAGNN_CHECK_LAUNCH();
  return QAGNN_OK;
}

}  // namespace

bool headtile_supported(const qagnn_shape& s) { return make_plan(s).ok; }

int32_t zero_head_pads(const qagnn_shape& s, float* qkmh, cudaStream_t st) {
  const int d = s.D / s.H, DP = head_dim_padded(d);
  if (DP == d) return QAGNN_OK;
  const int64_t rows = (int64_t)3 * s.H * s.N;
  int64_t g = (rows * (DP - d) + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  zero_head_pads_kernel<<<(unsigned)g, 256, 0, st>>>(rows, d, DP, qkmh);
  QAGNN_CHECK_LAUNCH();
  return QAGNN_OK;
}

int32_t launch_message_passing_headtile(const qagnn_shape& s, const int32_t* prep_base, const qagnn_prep_layout& L,
                                        const float* qkmh, const float* keh, const float* meh, float* score,
                                        float* alpha, float* aggr, float* alpha_out, void* aggr_hi, void* aggr_lo,
                                        cudaStream_t st) {
  const HeadTilePlan plan = make_plan(s);
  if (!plan.ok) return QAGNN_ERR_UNSUPPORTED;
  auto I = [&](size_t off) { return (const int32_t*)((const char*)prep_base + off); };
  HeadTileParams p;
  p.N = s.N; p.Eps = (s.N + s.E + 3) / 4 * 4;
  p.n = s.n_per_graph; p.G = (int)(s.N / s.n_per_graph); p.H = s.H; p.D = s.D; p.d = s.D / s.H; p.DP = plan.DP;
  p.C = plan.C; p.S = plan.S; p.W = plan.W; p.ecap = plan.ecap;
  p.rowptr_src = I(L.rowptr_src); p.rowptr_tgt = I(L.rowptr_tgt); p.pk_src = I(L.pk_src); p.pk_tgt = I(L.pk_tgt);
  p.tpos = I(L.csr_src_tpos); p.perm_src = I(L.perm_src); p.order_src = I(L.order_src); p.order_tgt = I(L.order_tgt);
  p.qkmh = qkmh; p.keh = keh; p.meh = meh; p.score = score; p.alpha = alpha; p.aggr = aggr; p.alpha_out = alpha_out;
  p.aggr_hi = (s.D % 2 == 0 && (s.D / s.H) % 2 == 0) ? aggr_hi : nullptr; p.aggr_lo = aggr_lo;
  if (aggr_hi != nullptr && p.aggr_hi == nullptr) return QAGNN_ERR_UNSUPPORTED;
  const unsigned grid = (unsigned)(plan.S * s.H), block = (unsigned)(plan.W + 1) * 32;
  // QAGNN_MP_TRACE=1: cycle accounting of the consumer warps (diagnostic; synchronises after every launch)
  static const bool trace_on = [] { const char* e = getenv("QAGNN_MP_TRACE"); return e && atoi(e) != 0; }();
  static unsigned long long* trace_buf = nullptr;
  p.trace = nullptr;
  if (trace_on) {
    if (!trace_buf) QAGNN_CHECK_CUDA(cudaMalloc(&trace_buf, 4 * sizeof(unsigned long long)));
    QAGNN_CHECK_CUDA(cudaMemsetAsync(trace_buf, 0, 4 * sizeof(unsigned long long), st));
    p.trace = trace_buf;
  }
  struct TraceDump {
    bool on; unsigned long long* buf; cudaStream_t st;
    ~TraceDump() {
      if (!on) return;
      unsigned long long h[4];
      cudaMemcpyAsync(h, buf, sizeof(h), cudaMemcpyDeviceToHost, st);
      cudaStreamSynchronize(st);
      if (h[3]) fprintf(stderr, "[qagnn mp trace] consumer warps %llu: avg cycles %.0f, waiting for tiles %.1f %%, for tables %.1f %%\n",
                        h[3], (double)h[0] / h[3], 100.0 * h[1] / h[0], 100.0 * h[2] / h[0]);
    }
  } dump{trace_on, trace_buf, st};
  // QAGNN_MP_VARIANT=1: the round-2 candidate consumer code (see the kernel's header); not the default until it has
  // been through the GPU parity suite and the bench on a B200
  const char* variant_env = getenv("QAGNN_MP_VARIANT");  // read per launch: tools/check_mp_variant.py flips it in-process
  const int variant = variant_env ? atoi(variant_env) : 0;
  if (variant == 2 && (long)p.n * p.DP * 4 < 65536) {  // = variant 1 + the proxy fence only before the phase switch
    if (plan.cpl == 1 && plan.qpw == 1) return launch_t<1, 1, 2>(p, plan, grid, block, st);
    if (plan.cpl == 1 && plan.qpw == 2) return launch_t<1, 2, 2>(p, plan, grid, block, st);
    if (plan.cpl == 1 && plan.qpw == 3) return launch_t<1, 3, 2>(p, plan, grid, block, st);
    if (plan.cpl == 1 && plan.qpw == 4) return launch_t<1, 4, 2>(p, plan, grid, block, st);
    if (plan.cpl == 2 && plan.qpw == 1) return launch_t<2, 1, 2>(p, plan, grid, block, st);
    if (plan.cpl == 2 && plan.qpw == 3) return launch_t<2, 3, 2>(p, plan, grid, block, st);
    if (plan.cpl == 2 && plan.qpw == 4) return launch_t<2, 4, 2>(p, plan, grid, block, st);
    return launch_t<2, 2, 2>(p, plan, grid, block, st);
  }
  if (variant == 1 && (long)p.n * p.DP * 4 < 65536) {  // byte offsets of tile rows are packed into 16 bits
    if (plan.cpl == 1 && plan.qpw == 1) return launch_t<1, 1, 1>(p, plan, grid, block, st);
    if (plan.cpl == 1 && plan.qpw == 2) return launch_t<1, 2, 1>(p, plan, grid, block, st);
    if (plan.cpl == 1 && plan.qpw == 3) return launch_t<1, 3, 1>(p, plan, grid, block, st);
    if (plan.cpl == 1 && plan.qpw == 4) return launch_t<1, 4, 1>(p, plan, grid, block, st);
    if (plan.cpl == 2 && plan.qpw == 1) return launch_t<2, 1, 1>(p, plan, grid, block, st);
    if (plan.cpl == 2 && plan.qpw == 3) return launch_t<2, 3, 1>(p, plan, grid, block, st);
    if (plan.cpl == 2 && plan.qpw == 4) return launch_t<2, 4, 1>(p, plan, grid, block, st);
    return launch_t<2, 2, 1>(p, plan, grid, block, st);
  }
  if (plan.cpl == 1 && plan.qpw == 1) return launch_t<1, 1, 0>(p, plan, grid, block, st);
  if (plan.cpl == 1 && plan.qpw == 2) return launch_t<1, 2, 0>(p, plan, grid, block, st);
  if (plan.cpl == 1 && plan.qpw == 3) return launch_t<1, 3, 0>(p, plan, grid, block, st);
  if (plan.cpl == 1 && plan.qpw == 4) return launch_t<1, 4, 0>(p, plan, grid, block, st);
  if (plan.cpl == 2 && plan.qpw == 1) return launch_t<2, 1, 0>(p, plan, grid, block, st);
  if (plan.cpl == 2 && plan.qpw == 3) return launch_t<2, 3, 0>(p, plan, grid, block, st);
  if (plan.cpl == 2 && plan.qpw == 4) return launch_t<2, 4, 0>(p, plan, grid, block, st);
  return launch_t<2, 2, 0>(p, plan, grid, block, st);
}

}  // namespace qagnn
