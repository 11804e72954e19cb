// Shared-memory-tiled message passing for batches of small sub-graphs (the QA-GNN case: <= 200
// nodes per (question, choice) graph).  One launch per GATConvE layer does
//     logits -> per-SOURCE softmax -> out-degree rescale -> per-TARGET weighted sum
// (modeling/modeling_qagnn.py:442,455-484) with every gathered row served from shared memory.
//
// Why this shape (profiles/r1_microbench.txt, profiles/r1_v1_mp_ncu.md): each edge needs four row
// gathers (Kx[tgt], Ke[combo], Mx[src], Me[combo]); served from L2 they cap at ~11 G rows/s, so the
// per-head slice of the edge tables (C x d floats, 125 KB at C=624, d=50) has to live in shared
// memory next to the node tile of the current graph.  Hence:
//   * grid = H x floor(#SM / H) persistent CTAs; CTA (h, slot) owns head h of graphs slot, slot+S, ...
//   * phase 1: Ke_h resident, Kx_h tiles of its graphs streamed through a 2-deep TMA (cp.async.bulk)
//     ring -> raw logits, online softmax per source node, rescaled weights a'[e] to global (L2);
//   * phase 2: Me_h swapped in, Mx_h tiles streamed the same way -> aggr[:, h*d:(h+1)*d].
//   * 8 lanes per node (quarter-warp), each lane owning float4 chunks l, l+8 of the padded head row:
//     a quarter-warp LDS.128 reads 128 contiguous bytes = one conflict-free wavefront; the dot
//     product needs 3 shuffles; the softmax is carried online in registers (no extra pass).
// Node rows come from the head-major padded projection layout [3][H][N][DP] written by the
// projection GEMM, so a tile is one contiguous n*DP*4-byte bulk copy.
#include "common.cuh"

namespace qagnn {

namespace {

struct HeadTileParams {
  int64_t N, Ep;
  int n, G, H, D, d, DP, C, S, W;
  const int32_t *rowptr_src, *rowptr_tgt, *pk_src, *pk_tgt, *apos, *perm_src;
  const float *qkmh, *keh, *meh;
  float *score, *alpha, *aggr, *alpha_out;
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
// large copies are split so that a single transaction count stays far below the 2^20-1 limit
__device__ __forceinline__ void bulk_g2s_chunked(char* dst, const char* src, uint32_t bytes, uint64_t* bar) {
  mbar_expect_tx(bar, bytes);
  const uint32_t kChunk = 32768;
  for (uint32_t o = 0; o < bytes; o += kChunk) bulk_g2s(dst + o, src + o, min(kChunk, bytes - o), bar);
}

__device__ __forceinline__ float dot4(const float4& q, const float4& a, const float4& b) {
  return q.x * (a.x + b.x) + q.y * (a.y + b.y) + q.z * (a.z + b.z) + q.w * (a.w + b.w);
}

template <int CPL>
__global__ void __launch_bounds__(1024, 1) mp_headtile_kernel(const HeadTileParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int NCH = p.DP / 4;
  float4* tab = reinterpret_cast<float4*>(smem_raw);
  const size_t tab_bytes = (size_t)p.C * p.DP * 4;
  const size_t tile_bytes = (size_t)p.n * p.DP * 4;
  float4* tile[2] = {reinterpret_cast<float4*>(smem_raw + tab_bytes),
                     reinterpret_cast<float4*>(smem_raw + tab_bytes + tile_bytes)};
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + tab_bytes + 2 * tile_bytes);
  uint64_t* full = bars;        // [2]
  uint64_t* empty = bars + 2;   // [2]
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

  if (warp == p.W) {
    // ===================== producer warp: one lane drives the TMA ring =====================
    if (lane == 0) {
      bulk_g2s_chunked((char*)tab, (const char*)(p.keh + (size_t)h * p.C * p.DP), (uint32_t)tab_bytes, tabbar);
      for (int t = 0; t < 2 * Gc; ++t) {
        const int b = t & 1;
        if (t == Gc) {
          // phase switch: every consumer has left phase 1 once the last two tiles are released
          mbar_wait(&empty[(Gc - 1) & 1], ((Gc - 1) >> 1) & 1);
          if (Gc >= 2) mbar_wait(&empty[(Gc - 2) & 1], ((Gc - 2) >> 1) & 1);
          bulk_g2s_chunked((char*)tab, (const char*)(p.meh + (size_t)h * p.C * p.DP), (uint32_t)tab_bytes, tabbar);
        }
        if (t >= 2) mbar_wait(&empty[b], ((t >> 1) - 1) & 1);
        const int g = slot + (t < Gc ? t : t - Gc) * p.S;
        const float* src = (t < Gc ? Kh : Mh) + (size_t)g * p.n * p.DP;
        bulk_g2s_chunked((char*)tile[b], (const char*)src, (uint32_t)tile_bytes, &full[b]);
      }
    }
    return;
  }

  // ========================= consumer warps: 4 nodes per warp, 8 lanes per node =========================
  const int l8 = lane & 7, qbase = lane & 24, qi = lane >> 3;
  const int nquads = (p.n + 3) / 4;
  bool cvalid[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) cvalid[k] = (l8 + 8 * k) < NCH;
  const size_t hEp = (size_t)h * p.Ep;

  // ---------------------------------- phase 1: attention weights ----------------------------------
  mbar_wait(tabbar, 0);
  for (int t = 0; t < Gc; ++t) {
    const int b = t & 1;
    const int g = slot + t * p.S;
    mbar_wait(&full[b], (t >> 1) & 1);
    const float4* kt = tile[b];
    for (int quad = warp; quad < nquads; quad += p.W) {
      const int vl = quad * 4 + qi;
      const bool nvalid = vl < p.n;
      const int64_t v = (int64_t)g * p.n + (nvalid ? vl : 0);
      const int beg = p.rowptr_src[v];
      const int deg = nvalid ? p.rowptr_src[v + 1] - beg : 0;
      int maxdeg = max(deg, __shfl_xor_sync(0xffffffffu, deg, 8));
      maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, 16));
      float4 q[CPL];
#pragma unroll
      for (int k = 0; k < CPL; ++k)
        q[k] = cvalid[k] ? *reinterpret_cast<const float4*>(Qh + v * p.DP + 4 * (l8 + 8 * k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      float m = -INFINITY, ssum = 0.f, skeep = 0.f;
      for (int i0 = 0; i0 < maxdeg; i0 += 8) {
        const int pkv = (i0 + l8 < deg) ? p.pk_src[beg + i0 + l8] : 0;
        const int lim = min(8, maxdeg - i0);
        for (int j = 0; j < lim; ++j) {
          const uint32_t w = (uint32_t)__shfl_sync(0xffffffffu, pkv, qbase + j);
          const float4* kr = kt + (size_t)(w >> 16) * NCH;
          const float4* er = tab + (size_t)(w & 0xffffu) * NCH;
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < CPL; ++k)
            if (cvalid[k]) s += dot4(q[k], kr[l8 + 8 * k], er[l8 + 8 * k]);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          if (i0 + j < deg) {
            // online softmax: one exp per edge
            const float e = __expf(-fabsf(s - m));
            if (s <= m) { ssum += e; } else { ssum = ssum * e + 1.f; m = s; }
            if (l8 == j) {
              if (i0 == 0) skeep = s; else p.score[hEp + beg + i0 + j] = s;
            }
          }
        }
      }
      if (maxdeg > 8) __syncwarp();
      const float denom = ssum + 1e-16f;  // torch_geometric.utils.softmax
      const float degf = (float)deg;
      for (int j = l8; j < deg; j += 8) {
        const float s = j < 8 ? skeep : p.score[hEp + beg + j];
        const float a = expf(s - m) / denom;
        p.alpha[hEp + beg + j] = a * degf;  // rescale by the out-degree of the source (:476-481)
        if (p.alpha_out != nullptr) p.alpha_out[(size_t)p.perm_src[beg + j] * p.H + h] = a;
      }
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[b]);
  }

  // ---------------------------------- phase 2: weighted sum by target ----------------------------------
  mbar_wait(tabbar, 1);
  for (int t = Gc; t < 2 * Gc; ++t) {
    const int b = t & 1;
    const int g = slot + (t - Gc) * p.S;
    mbar_wait(&full[b], (t >> 1) & 1);
    const float4* mt = tile[b];
    for (int quad = warp; quad < nquads; quad += p.W) {
      const int vl = quad * 4 + qi;
      const bool nvalid = vl < p.n;
      const int64_t v = (int64_t)g * p.n + (nvalid ? vl : 0);
      const int beg = p.rowptr_tgt[v];
      const int deg = nvalid ? p.rowptr_tgt[v + 1] - beg : 0;
      int maxdeg = max(deg, __shfl_xor_sync(0xffffffffu, deg, 8));
      maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, 16));
      float4 acc[CPL];
#pragma unroll
      for (int k = 0; k < CPL; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i0 = 0; i0 < maxdeg; i0 += 8) {
        int pkv = 0;
        float wv = 0.f;
        if (i0 + l8 < deg) {
          pkv = p.pk_tgt[beg + i0 + l8];
          wv = p.alpha[hEp + p.apos[beg + i0 + l8]];
        }
        const int lim = min(8, maxdeg - i0);
        for (int j = 0; j < lim; ++j) {
          const uint32_t w = (uint32_t)__shfl_sync(0xffffffffu, pkv, qbase + j);
          const float a = __shfl_sync(0xffffffffu, wv, qbase + j);  // 0 beyond this node's degree
          const float4* mr = mt + (size_t)(w >> 16) * NCH;
          const float4* er = tab + (size_t)(w & 0xffffu) * NCH;
#pragma unroll
          for (int k = 0; k < CPL; ++k) {
            if (cvalid[k]) {
              const float4 x = mr[l8 + 8 * k], y = er[l8 + 8 * k];
              acc[k].x += (x.x + y.x) * a;
              acc[k].y += (x.y + y.y) * a;
              acc[k].z += (x.z + y.z) * a;
              acc[k].w += (x.w + y.w) * a;
            }
          }
        }
      }
      if (nvalid) {
        float* out = p.aggr + v * p.D + (size_t)h * p.d;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const int c0 = 4 * (l8 + 8 * k);
          if (!cvalid[k] || c0 >= p.d) continue;
          if ((p.d & 3) == 0 && (p.D & 3) == 0) {
            *reinterpret_cast<float4*>(out + c0) = acc[k];
          } else if ((p.d & 1) == 0) {
            *reinterpret_cast<float2*>(out + c0) = make_float2(acc[k].x, acc[k].y);
            if (c0 + 2 < p.d) *reinterpret_cast<float2*>(out + c0 + 2) = make_float2(acc[k].z, acc[k].w);
          } else {
            out[c0] = acc[k].x;
            if (c0 + 1 < p.d) out[c0 + 1] = acc[k].y;
            if (c0 + 2 < p.d) out[c0 + 2] = acc[k].z;
            if (c0 + 3 < p.d) out[c0 + 3] = acc[k].w;
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[b]);
  }
}

__global__ void zero_head_pads_kernel(int64_t rows, int d, int DP, float* __restrict__ buf) {
  const int pad = DP - d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows * pad; i += (int64_t)gridDim.x * blockDim.x)
    buf[(i / pad) * DP + d + (i % pad)] = 0.f;
}

struct HeadTilePlan {
  bool ok;
  int DP, C, S, W, cpl, sms;
  size_t smem;
};

HeadTilePlan make_plan(const qagnn_shape& s) {
  HeadTilePlan pl{};
  pl.ok = false;
  if (s.n_per_graph <= 0 || s.n_per_graph > 65535 || s.N % s.n_per_graph != 0) return pl;
  const int d = s.D / s.H;
  pl.DP = head_dim_padded(d);
  pl.C = (s.R + 1) * s.T * s.T;
  if (pl.C > 65536 || pl.DP > 64) return pl;
  pl.cpl = pl.DP <= 32 ? 1 : 2;
  int dev = 0, sms = 0, max_smem = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return pl;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  pl.sms = sms;
  pl.smem = (size_t)pl.C * pl.DP * 4 + 2 * (size_t)s.n_per_graph * pl.DP * 4 + 64;
  if (pl.smem > (size_t)max_smem || s.H > sms) return pl;
  pl.S = sms / s.H;
  const int quads = (s.n_per_graph + 3) / 4;
  const int passes = (quads + 30) / 31;  // at most 31 consumer warps + 1 producer warp
  pl.W = (quads + passes - 1) / passes;
  pl.ok = true;
  return pl;
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
                                        float* alpha, float* aggr, float* alpha_out, cudaStream_t st) {
  const HeadTilePlan plan = make_plan(s);
  if (!plan.ok) return QAGNN_ERR_UNSUPPORTED;
  auto I = [&](size_t off) { return (const int32_t*)((const char*)prep_base + off); };
  HeadTileParams p;
  p.N = s.N; p.Ep = s.N + s.E;
  p.n = s.n_per_graph; p.G = (int)(s.N / s.n_per_graph); p.H = s.H; p.D = s.D; p.d = s.D / s.H; p.DP = plan.DP;
  p.C = plan.C; p.S = plan.S; p.W = plan.W;
  p.rowptr_src = I(L.rowptr_src); p.rowptr_tgt = I(L.rowptr_tgt); p.pk_src = I(L.pk_src); p.pk_tgt = I(L.pk_tgt);
  p.apos = I(L.csr_tgt_apos); p.perm_src = I(L.perm_src);
  p.qkmh = qkmh; p.keh = keh; p.meh = meh; p.score = score; p.alpha = alpha; p.aggr = aggr; p.alpha_out = alpha_out;
  const unsigned grid = (unsigned)(plan.S * s.H), block = (unsigned)(plan.W + 1) * 32;
  if (plan.cpl == 1) {
    QAGNN_CHECK_CUDA(cudaFuncSetAttribute(mp_headtile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    mp_headtile_kernel<1><<<grid, block, plan.smem, st>>>(p);
  } else {
    QAGNN_CHECK_CUDA(cudaFuncSetAttribute(mp_headtile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    mp_headtile_kernel<2><<<grid, block, plan.smem, st>>>(p);
  }
  QAGNN_CHECK_LAUNCH();
  return QAGNN_OK;
}

}  // namespace qagnn
