// Message passing for graphs that do not fit the per-graph shared-memory tiles of mp_headtile.cu (BASELINE.json
// configs[4]: 2000-node / 20000-edge sub-graphs, hidden 1024, 8 heads) — same math as message_passing.cu
// (modeling/modeling_qagnn.py:442,455-484), different data movement.
//
// The basic CSR kernels gather four full D-wide rows per edge from L2 (Kx[tgt], Ke[combo], Mx[src], Me[combo]: 16 KB
// per edge at D = 1024) and spend 40 shuffles per edge on the per-head reduction: 3.9 ms per layer at configs[4], 8 % of
// the HBM roofline.  Here the two edge tables never leave the SM: they are cut into column slices of SL <= 64 floats
// (a slice lies inside one head), and a persistent CTA keeps ONE slice of Ke (scores) or Me (aggregate) in shared
// memory (C x SL floats, 156 KB at C = 612, SL = 64) while it walks a partition of the nodes.  Per edge and slice it then
// reads SL*4 bytes of one node row from L2 (coalesced 64..256-byte segments, four rows in flight per lane group) and
// SL*4 bytes of a table row from shared memory; the L2 traffic halves and the table reads stop competing for it.
//   mp_slice_scores_kernel     CTA = (slice, node partition): partial logits of its slice; a head made of two slices
//                              accumulates with one atomicAdd per slice — two addends commute, so the result stays
//                              bit-reproducible (heads wider than 2 slices are not taken by this path)
//   mp_slice_softmax_kernel    one warp per source node: max / exp / sum over its out-edges, out-degree rescale
//   mp_slice_aggregate_kernel  CTA = (slice, node partition): weighted sum of (Mx[src] + Me[combo]) slices per target
#include "common.cuh"

namespace qagnn {

namespace {

constexpr int kSliceThreads = 1024;
constexpr int kUnroll = 4;  // edges in flight per lane group

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int SL>
__device__ __forceinline__ void load_table_slice(float* tab, const float* __restrict__ table, int C, int D, int col0) {
  constexpr int CH = SL / 4;
  for (int i = threadIdx.x; i < C * CH; i += blockDim.x) {
    const int c = i / CH, k = i - c * CH;
    reinterpret_cast<float4*>(tab)[i] = ldg4(table + (size_t)c * D + col0 + 4 * k);
  }
}

template <int SL>
__global__ void __launch_bounds__(kSliceThreads, 1) mp_slice_scores_kernel(int64_t N, int D, int H, int C, int parts,
                                                                           const int32_t* __restrict__ rowptr_src,
                                                                           const int32_t* __restrict__ csr_src_tgt,
                                                                           const int32_t* __restrict__ csr_src_combo,
                                                                           const float* __restrict__ qkm,
                                                                           const float* __restrict__ ke, float* __restrict__ score,
                                                                           int slices_per_head) {
  extern __shared__ __align__(16) float tab[];  // [C][SL]
  constexpr int LPE = SL / 4;                   // lanes per edge (one float4 each)
  constexpr int EPW = 32 / LPE;                 // edges a warp handles side by side
  const int slice = blockIdx.x % (D / SL), part = blockIdx.x / (D / SL);
  const int col0 = slice * SL, h = col0 / (D / H);
  load_table_slice<SL>(tab, ke, C, D, col0);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane / LPE, lc = lane % LPE;
  const int ld = 3 * D;
  const int64_t u0 = N * part / parts, u1 = N * (part + 1) / parts;
  for (int64_t u = u0 + warp; u < u1; u += nwarps) {
    const float4 q = ldg4(qkm + u * ld + col0 + 4 * lc);
    const int beg = rowptr_src[u], end = rowptr_src[u + 1];
    for (int pb = beg; pb < end; pb += EPW * kUnroll) {  // warp-uniform trip count: the shuffles below need all lanes
      const int p0 = pb + sub;
      float4 kx[kUnroll];
      int cb[kUnroll];
#pragma unroll
      for (int r = 0; r < kUnroll; ++r) {
        const int p = p0 + r * EPW;
        const bool ok = p < end;
        const int t = ok ? csr_src_tgt[p] : 0;
        cb[r] = ok ? csr_src_combo[p] : 0;
        kx[r] = ok ? ldg4(qkm + (int64_t)t * ld + D + col0 + 4 * lc) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int r = 0; r < kUnroll; ++r) {
        const int p = p0 + r * EPW;
        const float4 kt = reinterpret_cast<const float4*>(tab)[cb[r] * LPE + lc];
        float s = (q.x * (kx[r].x + kt.x) + q.y * (kx[r].y + kt.y)) + (q.z * (kx[r].z + kt.z) + q.w * (kx[r].w + kt.w));
#pragma unroll
        for (int o = LPE / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lc == 0 && p < end) {
          if (slices_per_head == 1) score[(int64_t)p * H + h] = s;
          else atomicAdd(score + (int64_t)p * H + h, s);  // two addends per address: order-independent
        }
      }
    }
  }
}

// logits [E', H] (by-source order) -> a' = softmax over the source's out-edges * out-degree, in place; optional un-scaled copy
// in edge_index' order (what return_attention_weights exposes).  One warp per source node; lane l owns head l % H.
__global__ void __launch_bounds__(256) mp_slice_softmax_kernel(int64_t N, int H, const int32_t* __restrict__ rowptr_src,
                                                               const int32_t* __restrict__ perm_src, float* __restrict__ score,
                                                               float* __restrict__ alpha_out) {
  const int lane = threadIdx.x & 31;
  const int64_t v = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (v >= N) return;
  const int beg = rowptr_src[v], deg = rowptr_src[v + 1] - beg;
  float* sc = score + (int64_t)beg * H;
  float m = -INFINITY;
  for (int j = lane; j < deg * H; j += 32) m = fmaxf(m, sc[j]);  // 32 % H == 0: lane keeps its head
  for (int o = 16; o >= H; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
  for (int j = lane; j < deg * H; j += 32) {
    const float ex = expf(sc[j] - m);
    sc[j] = ex;
    sum += ex;
  }
  for (int o = 16; o >= H; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float degf = (float)deg;
  for (int j = lane; j < deg * H; j += 32) {
    const float a = sc[j] / (sum + 1e-16f);  // torch_geometric.utils.softmax
    sc[j] = a * degf;                        // :476-481
    if (alpha_out != nullptr) alpha_out[(int64_t)perm_src[beg + j / H] * H + (lane % H)] = a;
  }
}

template <int SL>
__global__ void __launch_bounds__(kSliceThreads, 1) mp_slice_aggregate_kernel(int64_t N, int D, int H, int C, int parts,
                                                                              const int32_t* __restrict__ rowptr_tgt,
                                                                              const int32_t* __restrict__ csr_tgt_src,
                                                                              const int32_t* __restrict__ csr_tgt_combo,
                                                                              const int32_t* __restrict__ csr_tgt_apos,
                                                                              const float* __restrict__ qkm,
                                                                              const float* __restrict__ me,
                                                                              const float* __restrict__ alpha,
                                                                              float* __restrict__ aggr) {
  extern __shared__ __align__(16) float tab[];  // [C][SL]
  constexpr int LPE = SL / 4, EPW = 32 / LPE;
  const int slice = blockIdx.x % (D / SL), part = blockIdx.x / (D / SL);
  const int col0 = slice * SL, h = col0 / (D / H);
  load_table_slice<SL>(tab, me, C, D, col0);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane / LPE, lc = lane % LPE;
  const int ld = 3 * D;
  const int64_t v0 = N * part / parts, v1 = N * (part + 1) / parts;
  for (int64_t v = v0 + warp; v < v1; v += nwarps) {
    const int beg = rowptr_tgt[v], end = rowptr_tgt[v + 1];
    // lane group `sub` sums edges beg+sub, beg+sub+EPW, ... in that order; the groups are combined in a fixed tree below,
    // so the result is run-to-run identical
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int pb = beg; pb < end; pb += EPW * kUnroll) {
      const int p0 = pb + sub;
      float4 mx[kUnroll];
      int cb[kUnroll];
      float w[kUnroll];
#pragma unroll
      for (int r = 0; r < kUnroll; ++r) {
        const int p = p0 + r * EPW;
        const bool ok = p < end;
        const int s = ok ? csr_tgt_src[p] : 0;
        cb[r] = ok ? csr_tgt_combo[p] : 0;
        w[r] = ok ? alpha[(int64_t)csr_tgt_apos[p] * H + h] : 0.f;
        mx[r] = ok ? ldg4(qkm + (int64_t)s * ld + 2 * D + col0 + 4 * lc) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int r = 0; r < kUnroll; ++r) {
        const float4 mt = reinterpret_cast<const float4*>(tab)[cb[r] * LPE + lc];
        acc.x += (mx[r].x + mt.x) * w[r];
        acc.y += (mx[r].y + mt.y) * w[r];
        acc.z += (mx[r].z + mt.z) * w[r];
        acc.w += (mx[r].w + mt.w) * w[r];
      }
    }
#pragma unroll
    for (int o = 16; o >= LPE; o >>= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (sub == 0) *reinterpret_cast<float4*>(aggr + v * D + col0 + 4 * lc) = acc;
  }
}

struct SlicePlan {
  bool ok;
  int SL, slices, parts, per_head;
  size_t smem;
};

SlicePlan make_slice_plan(const qagnn_shape& s) {
  SlicePlan pl{};
  pl.ok = false;
  if (s.H <= 0 || s.D % s.H != 0 || 32 % s.H != 0) return pl;
  const int d = s.D / s.H;
  int SL = d >= 64 ? 64 : d;
  if (SL != 16 && SL != 32 && SL != 64) return pl;
  if (d % SL != 0 || d / SL > 2) return pl;  // at most two addends per logit: the atomic accumulation stays deterministic
  const int C = s.R * s.T * s.T + s.T;
  static int sms_c[kMaxDevices] = {0}, smem_c[kMaxDevices] = {0};
  const int dev = current_device();
  if (sms_c[dev] == 0) {
    cudaDeviceGetAttribute(&smem_c[dev], cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&sms_c[dev], cudaDevAttrMultiProcessorCount, dev);
  }
  pl.SL = SL;
  pl.slices = s.D / SL;
  pl.per_head = d / SL;
  pl.smem = (size_t)C * SL * sizeof(float);
  if (pl.smem > (size_t)smem_c[dev] || pl.slices > sms_c[dev]) return pl;
  pl.parts = sms_c[dev] / pl.slices;  // one CTA per SM: the table slice fills most of its shared memory
  if (pl.parts < 1) return pl;
  pl.ok = true;
  return pl;
}

template <int SL>
int32_t launch_slices(const qagnn_shape& s, const SlicePlan& pl, const int32_t* base, const qagnn_prep_layout& L,
                      const float* qkm, const float* ke, const float* me, float* score, float* aggr, float* alpha_out,
                      cudaStream_t st) {
  auto I = [&](size_t off) { return (const int32_t*)((const char*)base + off); };
  static size_t attr[kMaxDevices] = {0};
  const int dev = current_device();
  if (pl.smem > attr[dev]) {
    QAGNN_CHECK_CUDA(cudaFuncSetAttribute(mp_slice_scores_kernel<SL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    QAGNN_CHECK_CUDA(cudaFuncSetAttribute(mp_slice_aggregate_kernel<SL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    attr[dev] = pl.smem;
  }
  const int C = s.R * s.T * s.T + s.T;
  const unsigned grid = (unsigned)(pl.slices * pl.parts);
  if (pl.per_head > 1) QAGNN_CHECK_CUDA(cudaMemsetAsync(score, 0, (size_t)(s.N + s.E) * s.H * sizeof(float), st));
  mp_slice_scores_kernel<SL><<<grid, kSliceThreads, pl.smem, st>>>(s.N, s.D, s.H, C, pl.parts, I(L.rowptr_src), I(L.csr_src_tgt),
                                                                   I(L.csr_src_combo), qkm, ke, score, pl.per_head);
  QAGNN_CHECK_LAUNCH();
  mp_slice_softmax_kernel<<<(unsigned)((s.N * 32 + 255) / 256), 256, 0, st>>>(s.N, s.H, I(L.rowptr_src), I(L.perm_src), score,
                                                                              alpha_out);
  QAGNN_CHECK_LAUNCH();
  mp_slice_aggregate_kernel<SL><<<grid, kSliceThreads, pl.smem, st>>>(s.N, s.D, s.H, C, pl.parts, I(L.rowptr_tgt), I(L.csr_tgt_src),
                                                                      I(L.csr_tgt_combo), I(L.csr_tgt_apos), qkm, me, score, aggr);
  QAGNN_CHECK_LAUNCH();
  return QAGNN_OK;
}

}  // namespace

bool slice_supported(const qagnn_shape& s) { return s.D % 4 == 0 && make_slice_plan(s).ok; }

// `score` [E', H] is scratch on entry and holds a' (out-degree-scaled softmax, by-source order) on return.
int32_t launch_message_passing_slice(const qagnn_shape& s, const int32_t* prep_base, const qagnn_prep_layout& pl,
                                     const float* qkm, const float* ke, const float* me, float* score, float* aggr,
                                     float* alpha_out, cudaStream_t st) {
  const SlicePlan plan = make_slice_plan(s);
  if (!plan.ok) return QAGNN_ERR_UNSUPPORTED;
  switch (plan.SL) {
    case 16: return launch_slices<16>(s, plan, prep_base, pl, qkm, ke, me, score, aggr, alpha_out, st);
    case 32: return launch_slices<32>(s, plan, prep_base, pl, qkm, ke, me, score, aggr, alpha_out, st);
    default: return launch_slices<64>(s, plan, prep_base, pl, qkm, ke, me, score, aggr, alpha_out, st);
  }
}

}  // namespace qagnn
