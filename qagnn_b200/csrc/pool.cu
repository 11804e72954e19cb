// The step after the hot path: QAGNN's masked multi-head attention pooling over the node
// representations (utils/layers.py:324-371 MultiheadAttPoolLayer + :276-299 MatrixVectorScaledDotProductAttention,
// called at modeling/modeling_qagnn.py:180), eval mode.
//
//   qs = w_qs(sent) [B, nh*dk];  ks = w_ks(X) [B, n, nh, dk];  vs = w_vs(X) [B, n, nh, dv]
//   attn[h,b,:] = softmax_i( mask ? -inf : qs[b,h]·ks[b,i,h] / sqrt(dk) );  pooled[b, h*dv:(h+1)*dv] = Σ_i attn·vs[b,i,h]
//
// The reference materialises ks and vs with two [B*n, D] x [D, D] GEMMs.  Both projections are linear, so they fold
// into the query / out of the sum:   qs_h·(W_k,h x_i + b_k,h) = (W_k,hᵀ qs_h)·x_i + qs_h·b_k,h   and
// Σ_i a_i (W_v,h x_i + b_v,h) = W_v,h (Σ_i a_i x_i) + b_v,h   (Σ_i a_i = 1).  One CTA per graph then reads the
// graph's [n, D] node tile ONCE: logits for all heads, masked softmax, attention-weighted node sum; the two tiny
// per-graph matrix-vector products run in the same CTA.  No N-sized GEMM, no ks/vs round trip through HBM.
#include "common.cuh"

namespace qagnn {
namespace {

constexpr int kPoolThreads = 256;  // latency-bound kernel: <= 64 registers so that 4 CTAs (graphs) share an SM

// dynamic smem: qk[nh][D] | logit[nh][n] | xs[nh][D] | cst[nh, padded to 8] | part[warps][nh][D]
__global__ void __launch_bounds__(kPoolThreads, 4) attention_pool_kernel(int n, int D, int nh, const float* __restrict__ X,
                                                                        const float* __restrict__ qs, const unsigned char* __restrict__ mask,
                                                                        const float* __restrict__ wk, const float* __restrict__ bk,
                                                                        const float* __restrict__ wv, const float* __restrict__ bv,
                                                                        float* __restrict__ pooled, float* __restrict__ attn_out, int B,
                                                                        const int64_t* __restrict__ node_type = nullptr,
                                                                        const int64_t* __restrict__ adj_lengths = nullptr,
                                                                        const float* __restrict__ sent = nullptr, int S = 0,
                                                                        int pooled_ld = 0) {
  // mask == nullptr: "decoder tail" mode (modeling_qagnn.py:172-187) — the pool mask is derived here from adj_lengths and
  // node_type, and the row of cat(graph_vecs, sent_vecs, Z) is written in place: pooled -> columns [0, D) of a
  // [B, pooled_ld = 2D+S] buffer, sent_vecs -> [D, D+S), Z = X[b, 0, :] -> [D+S, 2D+S)
  extern __shared__ float sm[];
  float* qk = sm;                  // [nh][D]   W_k,hᵀ qs_h
  float* logit = qk + nh * D;      // [nh][n]
  float* xs = logit + nh * n;      // [nh][D]   Σ_i attn_i x_i
  float* cst = xs + nh * D;        // [nh]      qs_h·b_k,h
  const int b = blockIdx.x, tid = threadIdx.x, dk = D / nh;
  const float inv_temp = 1.0f / sqrtf((float)dk);
  const float* Xb = X + (size_t)b * n * D;
  const bool tail = mask == nullptr;
  if (!tail) pooled_ld = D;
  int len = n;
  bool all_masked = false;
  if (tail) {
    len = (int)min((int64_t)n, max((int64_t)0, adj_lengths[b]));
    // mask = (i >= len) | (type == 3); a fully masked row keeps node 0 (:174-176)
    int any_open = 0;
    for (int i = tid; i < len; i += kPoolThreads) any_open |= node_type[(size_t)b * n + i] != 3;
    all_masked = __syncthreads_or(any_open) == 0;
  }
  // fold the key projection into the query: qk[h][j] = Σ_{r in head h} qs[b, r] * wk[r, j]
  for (int idx = tid; idx < nh * D; idx += kPoolThreads) {
    const int h = idx / D, j = idx % D;
    float acc = 0.f;
#pragma unroll 10
    for (int r = 0; r < dk; ++r) acc = fmaf(qs[(size_t)b * D + h * dk + r], wk[(size_t)(h * dk + r) * D + j], acc);
    qk[idx] = acc;
  }
  for (int h = tid >> 5; h < nh; h += kPoolThreads / 32) {  // one warp per head: qs_h . b_k,h
    float acc = 0.f;
    for (int r = tid & 31; r < dk; r += 32) acc = fmaf(qs[(size_t)b * D + h * dk + r], bk[h * dk + r], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) cst[h] = acc;
  }
  __syncthreads();
  // logits: warps split the nodes, lanes the columns; 4 nodes per step so that up to 16 independent 128-byte row-segment
  // loads are in flight per warp (the tile comes from HBM/L2 exactly here)
  const int warp = tid >> 5, lane = tid & 31, nwarps = kPoolThreads / 32;
  for (int i0 = warp * 4; i0 < n; i0 += nwarps * 4) {
    for (int h0 = 0; h0 < nh; h0 += 4) {  // up to 4 heads per pass (register budget)
      float accl[4][4];  // [node][head]
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) accl[r][hh] = 0.f;
      for (int j0 = 0; j0 < D; j0 += 32 * 4) {
        float x[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int j = j0 + 32 * c + lane;
            x[r][c] = (i0 + r < n && j < D) ? Xb[(size_t)(i0 + r) * D + j] : 0.f;
          }
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          if (h0 + hh < nh) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int j = j0 + 32 * c + lane;
              const float qv = j < D ? qk[(h0 + hh) * D + j] : 0.f;
#pragma unroll
              for (int r = 0; r < 4; ++r) accl[r][hh] = fmaf(qv, x[r][c], accl[r][hh]);
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = i0 + r;
        if (i < n) {
          const bool masked = tail ? ((i >= len || node_type[(size_t)b * n + i] == 3) && !(all_masked && i == 0))
                                   : mask[(size_t)b * n + i] != 0;
#pragma unroll
          for (int hh = 0; hh < 4; ++hh) {
            if (h0 + hh < nh) {
              float acc = accl[r][hh];
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
              if (lane == 0) logit[(h0 + hh) * n + i] = masked ? -INFINITY : (acc + cst[h0 + hh]) * inv_temp;
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // masked softmax per head (one warp per head)
  for (int h = warp; h < nh; h += nwarps) {
    float m = -INFINITY;
    for (int i = lane; i < n; i += 32) m = fmaxf(m, logit[h * n + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int i = lane; i < n; i += 32) {
      const float e = expf(logit[h * n + i] - m);
      logit[h * n + i] = e;
      s += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    for (int i = lane; i < n; i += 32) {
      const float a = logit[h * n + i] / s;
      logit[h * n + i] = a;
      attn_out[((size_t)h * B + b) * n + i] = a;  // head-major [nh*B, n] like the reference
    }
  }
  __syncthreads();
  // attention-weighted node sum: xs[h][j] = Σ_i attn[h][i] * X[i][j].  Warps split the nodes, lanes the columns (coalesced
  // 128-byte row segments, every load independent of the accumulators: the kernel is latency bound, so what matters is
  // how many loads are in flight); per-warp partial sums meet in shared memory.
  float* part = cst + ((nh + 7) / 8) * 8;  // [nwarps][nh][D]
  for (int j0 = 0; j0 < D; j0 += 32 * 4) {
    float acc[4][4];  // [column chunk][head] — up to 4 heads per pass
    for (int h0 = 0; h0 < nh; h0 += 4) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) acc[c][hh] = 0.f;
#pragma unroll 4
      for (int i = warp; i < n; i += nwarps) {
        float x[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int j = j0 + 32 * c + lane;
          x[c] = j < D ? Xb[(size_t)i * D + j] : 0.f;
        }
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          if (h0 + hh < nh) {
            const float a = logit[(h0 + hh) * n + i];
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c][hh] = fmaf(a, x[c], acc[c][hh]);
          }
        }
      }
#pragma unroll
      for (int hh = 0; hh < 4; ++hh)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int j = j0 + 32 * c + lane;
          if (h0 + hh < nh && j < D) part[((size_t)warp * nh + h0 + hh) * D + j] = acc[c][hh];
        }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < nh * D; idx += kPoolThreads) {
    float acc = 0.f;
    for (int w = 0; w < nwarps; ++w) acc += part[(size_t)w * nh * D + idx];
    xs[idx] = acc;
  }
  __syncthreads();
  // value projection of the pooled vector: pooled[b, h*dv + r] = wv[h*dv + r, :]·xs[h] + bv   (one warp per output row:
  // coalesced reads of the weight row, shuffle reduction)
  for (int r0 = warp * 4; r0 < D; r0 += nwarps * 4) {  // 4 rows per step: their weight loads are all independent
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = lane; j < D; j += 32) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int idx = min(r0 + r, D - 1);
        acc[r] = fmaf(wv[(size_t)idx * D + j], xs[(idx / dk) * D + j], acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float a = acc[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0 && r0 + r < D) pooled[(size_t)b * pooled_ld + r0 + r] = a + bv[r0 + r];
    }
  }
  if (tail) {
    float* row = pooled + (size_t)b * pooled_ld;
    for (int j = tid; j < S; j += kPoolThreads) row[D + j] = sent[(size_t)b * S + j];
    for (int j = tid; j < D; j += kPoolThreads) row[D + S + j] = Xb[j];
  }
}

}  // namespace
}  // namespace qagnn

using namespace qagnn;

namespace qagnn {
namespace {

// The step before the hot path: QAGNN.forward's input assembly (modeling/modeling_qagnn.py:153-167), eval mode.
//   H[b,0,:]   = ctx[b,:]                      (= GELU(svec2nvec(sent_vecs)), computed by the caller: a [B, sent_dim] GEMM)
//   H[b,i,:]   = table[concept_ids[b,i] - 1]    i >= 1; `table` = the concept embedding AFTER cpt_transform + GELU, folded
//                                               once per weight set (the table is frozen in eval), so this is a pure gather
//   s          = -(score - score[:,0]) * (i < adj_len);  scores_out = s / (sum_i |s| / adj_len + 1e-5)
// One CTA per graph: the rows are copied as float4, the score norm is one block reduction.
__global__ void __launch_bounds__(256) decoder_head_kernel(int n, int D, const int64_t* __restrict__ concept_ids, int64_t n_concept,
                                                            const float* __restrict__ table, const float* __restrict__ ctx,
                                                            const float* __restrict__ node_scores,
                                                            const int64_t* __restrict__ adj_lengths, float* __restrict__ H_out,
                                                            float* __restrict__ scores_out) {
  __shared__ float red[32];
  __shared__ float s_total;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t* cid = concept_ids + (size_t)b * n;
  const int D4 = D / 4;
  float* Hb = H_out + (size_t)b * n * D;
  for (int i = tid; i < n * D4; i += 256) {
    const int row = i / D4, c = i - row * D4;
    const float* srcp;
    if (row == 0) {
      srcp = ctx + (size_t)b * D;
    } else {
      int64_t id = cid[row] - 1;
      id = id < 0 ? 0 : (id >= n_concept ? n_concept - 1 : id);
      srcp = table + (size_t)id * D;
    }
    reinterpret_cast<float4*>(Hb + (size_t)row * D)[c] = __ldg(reinterpret_cast<const float4*>(srcp) + c);
  }
  const float* sc = node_scores + (size_t)b * n;
  const float s0 = -sc[0];
  const int64_t len = adj_lengths[b];
  float part = 0.f;
  for (int i = tid; i < n; i += 256) part += (i < len) ? fabsf(-sc[i] - s0) : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    s_total = t;
  }
  __syncthreads();
  const float denom = s_total / (float)len + 1e-05f;
  for (int i = tid; i < n; i += 256) scores_out[(size_t)b * n + i] = ((i < len) ? (-sc[i] - s0) : 0.f) / denom;
}

}  // namespace
}  // namespace qagnn

extern "C" int32_t qagnn_decoder_head(int32_t B, int32_t n, int32_t D, const int64_t* concept_ids, int64_t n_concept,
                                      const float* table, const float* ctx, const float* node_scores,
                                      const int64_t* adj_lengths, float* H_out, float* scores_out, void* stream) {
  if (B <= 0 || n <= 0 || D <= 0 || D % 4 != 0 || n_concept <= 0) return QAGNN_ERR_INVALID_ARGUMENT;
  if (!concept_ids || !table || !ctx || !node_scores || !adj_lengths || !H_out || !scores_out) return QAGNN_ERR_INVALID_ARGUMENT;
  qagnn::decoder_head_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(n, D, concept_ids, n_concept, table, ctx, node_scores,
                                                                   adj_lengths, H_out, scores_out);
  QAGNN_CHECK_LAUNCH();
  return QAGNN_OK;
}

extern "C" int32_t qagnn_decoder_tail(int32_t B, int32_t n, int32_t D, int32_t n_head, int32_t S, const float* X, const float* qs,
                                      const int64_t* node_type, const int64_t* adj_lengths, const float* sent_vecs,
                                      const float* wk, const float* bk, const float* wv, const float* bv, float* concat,
                                      float* attn, void* stream) {
  if (B <= 0 || n <= 0 || D <= 0 || n_head <= 0 || D % n_head != 0 || S < 0) return QAGNN_ERR_INVALID_ARGUMENT;
  if (n_head > 8) return QAGNN_ERR_UNSUPPORTED;
  if (!X || !qs || !node_type || !adj_lengths || !sent_vecs || !wk || !bk || !wv || !bv || !concat || !attn)
    return QAGNN_ERR_INVALID_ARGUMENT;
  const size_t smem = ((size_t)2 * n_head * D + (size_t)n_head * n + n_head + 16 + (size_t)(kPoolThreads / 32) * n_head * D) * sizeof(float);
  if (smem > 200 * 1024) return QAGNN_ERR_UNSUPPORTED;
  static size_t attr[kMaxDevices] = {0};
  const int dev = current_device();
  if (smem > 48 * 1024 && smem > attr[dev]) {
    QAGNN_CHECK_CUDA(cudaFuncSetAttribute(attention_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[dev] = smem;
  }
  attention_pool_kernel<<<B, kPoolThreads, smem, (cudaStream_t)stream>>>(n, D, n_head, X, qs, nullptr, wk, bk, wv, bv, concat, attn, B,
                                                                         node_type, adj_lengths, sent_vecs, S, 2 * D + S);
  QAGNN_CHECK_LAUNCH();
  return QAGNN_OK;
}

extern "C" int32_t qagnn_attention_pool(int32_t B, int32_t n, int32_t D, int32_t n_head, const float* X, const float* qs,
                                        const uint8_t* mask, const float* wk, const float* bk, const float* wv, const float* bv,
                                        float* pooled, float* attn, void* stream) {
  if (B <= 0 || n <= 0 || D <= 0 || n_head <= 0 || D % n_head != 0) return QAGNN_ERR_INVALID_ARGUMENT;
  if (n_head > 8) return QAGNN_ERR_UNSUPPORTED;
  if (!X || !qs || !mask || !wk || !bk || !wv || !bv || !pooled || !attn) return QAGNN_ERR_INVALID_ARGUMENT;
  const size_t smem = ((size_t)2 * n_head * D + (size_t)n_head * n + n_head + 16 + (size_t)(kPoolThreads / 32) * n_head * D) * sizeof(float);
  if (smem > 200 * 1024) return QAGNN_ERR_UNSUPPORTED;
  static size_t attr[kMaxDevices] = {0};
  const int dev = current_device();
  if (smem > 48 * 1024 && smem > attr[dev]) {
    QAGNN_CHECK_CUDA(cudaFuncSetAttribute(attention_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[dev] = smem;
  }
  attention_pool_kernel<<<B, kPoolThreads, smem, (cudaStream_t)stream>>>(n, D, n_head, X, qs, mask, wk, bk, wv, bv, pooled, attn, B);
  QAGNN_CHECK_LAUNCH();
  return QAGNN_OK;
}
