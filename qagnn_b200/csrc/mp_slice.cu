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
constexpr int kUnroll = 2;  // edges in flight per lane group (x 2 chunks per lane at SL = 64)

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int SL>
__device__ __forceinline__ void load_table_slice(float* tab, const float* __restrict__ table, int C, int D, int col0) {
  constexpr int CH = SL / 4;
  for (int i = threadIdx.x; i < C * CH; i += blockDim.x) {
    const int c = i / CH, k = i - c * CH;
    reinterpret_cast<float4*>(tab)[i] = ldg4(table + (size_t)c * D + col0 + 4 * k);
  }
}

// lanes per edge: 8 lanes x 2 float4 for SL = 64 (half the shuffles, index loads and address arithmetic per gathered byte of
// a 16-lane x 1-chunk mapping: the kernels are instruction bound, profiles/r2_cfg5_ncu.md), else one float4 per lane
template <int SL> struct SliceMap {
  static constexpr int CPL = SL >= 64 ? 2 : 1;   // float4 chunks per lane
  static constexpr int LPE = SL / (4 * CPL);     // lanes per edge
  static constexpr int EPW = 32 / LPE;           // edges a warp handles side by side
};

// A warp walks TWO consecutive nodes per step: their CSR ranges are adjacent, so the pair is one run of ~2*degree edges
// (at degree 11 and 8 edge slots per iteration a single node wastes 31 % of the slots, a pair 8 %), and the index words
// of the next iteration are requested before the current rows are consumed (one L2 round trip hidden per iteration).
template <int SL>
__global__ void __launch_bounds__(kSliceThreads, 1) mp_slice_scores_kernel(int64_t N, int D, int H, int C, int parts,
                                                                           const int32_t* __restrict__ rowptr_src,
                                                                           const int32_t* __restrict__ csr_src_tgt,
                                                                           const int32_t* __restrict__ csr_src_combo,
                                                                           const float* __restrict__ qkm,
                                                                           const float* __restrict__ ke, float* __restrict__ score,
                                                                           int slices_per_head) {
  extern __shared__ __align__(16) float tab[];  // [C][SL]
  constexpr int CPL = SliceMap<SL>::CPL, LPE = SliceMap<SL>::LPE, EPW = SliceMap<SL>::EPW, RCH = SL / 4;
  const int slice = blockIdx.x % (D / SL), part = blockIdx.x / (D / SL);
  const int col0 = slice * SL, h = col0 / (D / H);
  load_table_slice<SL>(tab, ke, C, D, col0);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane / LPE, lc = lane % LPE;
  const uint32_t ld = 3u * (uint32_t)D;
  const float4* tab4 = reinterpret_cast<const float4*>(tab);
  const float* kbase = qkm + D + col0 + 4 * lc;  // Kx slice, this lane's first chunk (chunk k is LPE float4 further)
  const int64_t u0 = N * part / parts, u1 = N * (part + 1) / parts;
  for (int64_t u = u0 + 2 * warp; u < u1; u += 2 * nwarps) {
    const bool two = u + 1 < u1;
    float4 qa[CPL], qb[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      qa[k] = ldg4(qkm + u * ld + col0 + 4 * (lc + LPE * k));
      qb[k] = two ? ldg4(qkm + (u + 1) * ld + col0 + 4 * (lc + LPE * k)) : qa[k];
    }
    const int beg = rowptr_src[u], mid = rowptr_src[u + 1], end = two ? rowptr_src[u + 2] : mid;
    // index words of the first iteration
    uint32_t tn[kUnroll];
    int cn[kUnroll];
#pragma unroll
    for (int r = 0; r < kUnroll; ++r) {
      const int p = beg + sub + r * EPW;
      tn[r] = p < end ? (uint32_t)csr_src_tgt[p] : 0u;
      cn[r] = p < end ? csr_src_combo[p] : 0;
    }
    for (int pb = beg; pb < end; pb += EPW * kUnroll) {  // warp-uniform trip count: the shuffles below need all lanes
      const int p0 = pb + sub;
      float4 kx[kUnroll][CPL];
      int cb[kUnroll];
#pragma unroll
      for (int r = 0; r < kUnroll; ++r) {
        const bool ok = p0 + r * EPW < end;
        cb[r] = cn[r];
        const float* row = kbase + (size_t)(tn[r] * ld);  // N * 3D < 2^31 (checked by the plan)
#pragma unroll
        for (int k = 0; k < CPL; ++k) kx[r][k] = ok ? ldg4(row + 4 * LPE * k) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int r = 0; r < kUnroll; ++r) {  // next iteration's index words, in flight while this one's rows arrive
        const int p = p0 + EPW * kUnroll + r * EPW;
        tn[r] = p < end ? (uint32_t)csr_src_tgt[p] : 0u;
        cn[r] = p < end ? csr_src_combo[p] : 0;
      }
#pragma unroll
      for (int r = 0; r < kUnroll; ++r) {
        const int p = p0 + r * EPW;
        const bool first = p < mid;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const float4 kt = tab4[cb[r] * RCH + lc + LPE * k];
          const float4 q = first ? qa[k] : qb[k];
          s += (q.x * (kx[r][k].x + kt.x) + q.y * (kx[r][k].y + kt.y)) + (q.z * (kx[r][k].z + kt.z) + q.w * (kx[r][k].w + kt.w));
        }
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
  constexpr int CPL = SliceMap<SL>::CPL, LPE = SliceMap<SL>::LPE, EPW = SliceMap<SL>::EPW, RCH = SL / 4;
  const int slice = blockIdx.x % (D / SL), part = blockIdx.x / (D / SL);
  const int col0 = slice * SL, h = col0 / (D / H);
  load_table_slice<SL>(tab, me, C, D, col0);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane / LPE, lc = lane % LPE;
  const uint32_t ld = 3u * (uint32_t)D;
  const float4* tab4 = reinterpret_cast<const float4*>(tab);
  const float* mbase = qkm + 2 * D + col0 + 4 * lc;
  const float* abase = alpha + h;
  const int64_t v0 = N * part / parts, v1 = N * (part + 1) / parts;
  for (int64_t v = v0 + 2 * warp; v < v1; v += 2 * nwarps) {  // two consecutive targets per step (see the scores kernel)
    const bool two = v + 1 < v1;
    const int beg = rowptr_tgt[v], mid = rowptr_tgt[v + 1], end = two ? rowptr_tgt[v + 2] : mid;
    // lane group `sub` sums edges beg+sub, beg+sub+EPW, ... of the pair's run in that order, each edge into its own target's
    // accumulator; the groups are combined in a fixed tree below, so the result is run-to-run identical
    float4 acca[CPL], accb[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) acca[k] = accb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t sn[kUnroll], an[kUnroll];
    int cn[kUnroll];
#pragma unroll
    for (int r = 0; r < kUnroll; ++r) {
      const int p = beg + sub + r * EPW;
      sn[r] = p < end ? (uint32_t)csr_tgt_src[p] : 0u;
      cn[r] = p < end ? csr_tgt_combo[p] : 0;
      an[r] = p < end ? (uint32_t)csr_tgt_apos[p] : 0u;
    }
    for (int pb = beg; pb < end; pb += EPW * kUnroll) {
      const int p0 = pb + sub;
      float4 mx[kUnroll][CPL];
      int cb[kUnroll];
      float w[kUnroll];
#pragma unroll
      for (int r = 0; r < kUnroll; ++r) {
        const bool ok = p0 + r * EPW < end;
        cb[r] = cn[r];
        w[r] = ok ? abase[(size_t)(an[r] * (uint32_t)H)] : 0.f;
        const float* row = mbase + (size_t)(sn[r] * ld);
#pragma unroll
        for (int k = 0; k < CPL; ++k) mx[r][k] = ok ? ldg4(row + 4 * LPE * k) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int r = 0; r < kUnroll; ++r) {
        const int p = p0 + EPW * kUnroll + r * EPW;
        sn[r] = p < end ? (uint32_t)csr_tgt_src[p] : 0u;
        cn[r] = p < end ? csr_tgt_combo[p] : 0;
        an[r] = p < end ? (uint32_t)csr_tgt_apos[p] : 0u;
      }
#pragma unroll
      for (int r = 0; r < kUnroll; ++r) {
        const bool first = p0 + r * EPW < mid;
        const float wa = first ? w[r] : 0.f, wb = first ? 0.f : w[r];
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const float4 mt = tab4[cb[r] * RCH + lc + LPE * k];
          const float mxv = mx[r][k].x + mt.x, myv = mx[r][k].y + mt.y, mzv = mx[r][k].z + mt.z, mwv = mx[r][k].w + mt.w;
          acca[k].x += mxv * wa; acca[k].y += myv * wa; acca[k].z += mzv * wa; acca[k].w += mwv * wa;
          accb[k].x += mxv * wb; accb[k].y += myv * wb; accb[k].z += mzv * wb; accb[k].w += mwv * wb;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
#pragma unroll
      for (int o = 16; o >= LPE; o >>= 1) {
        acca[k].x += __shfl_xor_sync(0xffffffffu, acca[k].x, o);
        acca[k].y += __shfl_xor_sync(0xffffffffu, acca[k].y, o);
        acca[k].z += __shfl_xor_sync(0xffffffffu, acca[k].z, o);
        acca[k].w += __shfl_xor_sync(0xffffffffu, acca[k].w, o);
        accb[k].x += __shfl_xor_sync(0xffffffffu, accb[k].x, o);
        accb[k].y += __shfl_xor_sync(0xffffffffu, accb[k].y, o);
        accb[k].z += __shfl_xor_sync(0xffffffffu, accb[k].z, o);
        accb[k].w += __shfl_xor_sync(0xffffffffu, accb[k].w, o);
      }
      if (sub == 0) {
        *reinterpret_cast<float4*>(aggr + v * D + col0 + 4 * (lc + LPE * k)) = acca[k];
        if (two) *reinterpret_cast<float4*>(aggr + (v + 1) * D + col0 + 4 * (lc + LPE * k)) = accb[k];
      }
    }
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
  if (d % SL != 0 || d / SL > 2) return pl;
  if ((long long)s.N * 3 * s.D >= ((long long)1 << 31) || (long long)(s.N + s.E) * s.H >= ((long long)1 << 31)) return pl;  // 32-bit element offsets  // at most two addends per logit: the atomic accumulation stays deterministic
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
