// Weight folding: turns the reference's parameters (state_dict layout) into what the node-level
// formulation needs (SURVEY.md §8a "exact per-layer math"):
//   tab[c]  = edge_encoder(onehot(c)), c in [0, R*T*T + T)      modeling_qagnn.py:30,419-433  (layer-invariant)
//   Ke, Me  = tab @ W_k[:,2D:]^T + b_k,  tab @ W_m[:,2D:]^T + b_m  (edge part of linear_key/linear_msg :464-465)
//   Wp      = [W_q/sqrt(d) ; W_k[:,:2D] ; W_m[:,:2D]]            (node part, one [3D,2D] projection :464-466,469)
//   W1',b1' = BatchNorm1d(eval) folded into mlp.0                  (:408)
//   type_tab= GELU(emb_node_type(onehot(t)))                       (:65-66)
//   Vcat    = [Vh | Vx], vbias = b_h + b_x                          (:92)
// Runs on the device through this library's own kernels; the caller owns the blob.
#include <math.h>

#include "common.cuh"

namespace qagnn {

FoldLayout make_fold_layout(const qagnn_shape& s) {
  FoldLayout L;
  L.D = s.D; L.H = s.H; L.T = s.T; L.R = s.R; L.k = s.k;
  L.C = s.R * s.T * s.T + s.T;  // R*T*T real-edge combos + T self-loop combos
  const size_t D = s.D, C = L.C, Dh = s.D / 2;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += align_up(n * 4) / 4; return r; };
  L.tab = take(C * D);
  L.hidden = take(C * D);
  L.type_tab = take((size_t)s.T * Dh);
  L.basis = take(Dh);
  L.ws = take(Dh * Dh);
  L.bs = take(Dh);
  L.vcat = take(D * 2 * D);
  L.vbias = take(D);
  L.vcat_hi = take(D * 2 * D / 2);
  L.vcat_lo = take(D * 2 * D / 2);
  const size_t KSh = (size_t)round_up8((int)Dh), KS = (size_t)round_up8((int)(D + Dh));
  L.ws_hi = take(Dh * KSh / 2 + 8);
  L.ws_lo = take(Dh * KSh / 2 + 8);
  L.layer0 = o;
  size_t lo = 0;
  auto ltake = [&](size_t n) { size_t r = lo; lo += align_up(n * 4) / 4; return r; };
  L.wp = ltake(3 * D * 2 * D);
  L.bp = ltake(3 * D);
  L.ke = ltake(C * D);
  L.me = ltake(C * D);
  L.w1 = ltake(D * D);
  L.b1 = ltake(D);
  L.w2 = ltake(D * D);
  L.b2 = ltake(D);
  const size_t DP = (size_t)head_dim_padded(s.D / s.H);
  L.keh = ltake((size_t)s.H * C * DP);
  L.meh = ltake((size_t)s.H * C * DP);
  L.wph = ltake(3 * (size_t)s.H * DP * 2 * D);
  L.bph = ltake(3 * (size_t)s.H * DP);
  L.wph_hi = ltake(3 * (size_t)s.H * DP * 2 * D / 2);
  L.wph_lo = ltake(3 * (size_t)s.H * DP * 2 * D / 2);
  L.wp_hi = ltake(3 * D * 2 * D / 2);
  L.wp_lo = ltake(3 * D * 2 * D / 2);
  L.w1_hi = ltake(D * D / 2);
  L.w1_lo = ltake(D * D / 2);
  L.w2_hi = ltake(D * D / 2);
  L.w2_lo = ltake(D * D / 2);
  L.wps = ltake(3 * (size_t)s.H * DP * KS);
  L.wps_hi = ltake(3 * (size_t)s.H * DP * KS / 2 + 8);
  L.wps_lo = ltake(3 * (size_t)s.H * DP * KS / 2 + 8);
  L.tbias = ltake((size_t)s.T * 3 * s.H * DP);
  L.layer_stride = lo;
  L.total = o + lo * (size_t)(s.k > 0 ? s.k : 0);
  return L;
}

namespace {

__global__ void fold_edge_hidden_kernel(int C, int D, int T, int R, const float* __restrict__ w0,
                                        const float* __restrict__ b0, const float* __restrict__ g,
                                        const float* __restrict__ beta, const float* __restrict__ mean,
                                        const float* __restrict__ var, float* __restrict__ hidden) {
  const int F = R + 1 + 2 * T;  // width of the reference's one-hot edge feature
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)C * D; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i / D), f = (int)(i % D);
    int r = c / (T * T), ts = (c / T) % T, tt = c % T;
    if (c >= R * T * T) { r = R; ts = tt = c - R * T * T; }  // self loop of a type-ts node (:420-429)
    const float* w = w0 + (size_t)f * F;
    float pre = ((w[r] + w[R + 1 + ts]) + w[R + 1 + T + tt]) + b0[f];
    float y = (pre - mean[f]) / sqrtf(var[f] + 1e-5f) * g[f] + beta[f];
    hidden[i] = fmaxf(y, 0.f);
  }
}

__global__ void fold_layer_kernel(int D, int H, const float* __restrict__ qw, const float* __restrict__ qb,
                                  const float* __restrict__ kw, const float* __restrict__ mw,
                                  const float* __restrict__ w1, const float* __restrict__ b1,
                                  const float* __restrict__ g, const float* __restrict__ beta,
                                  const float* __restrict__ mean, const float* __restrict__ var,
                                  const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ wp,
                                  float* __restrict__ bp, float* __restrict__ w1o, float* __restrict__ b1o,
                                  float* __restrict__ w2o, float* __restrict__ b2o) {
  const float inv = 1.0f / sqrtf((float)(D / H));
  const int64_t nwp = (int64_t)3 * D * 2 * D;
  const int64_t total = nwp + 3 * D + (int64_t)D * D + D + (int64_t)D * D + D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t j = i;
    if (j < nwp) {
      const int row = (int)(j / (2 * D)), col = (int)(j % (2 * D));
      float v;
      if (row < D) v = qw[(size_t)row * 2 * D + col] * inv;
      else if (row < 2 * D) v = kw[(size_t)(row - D) * 3 * D + col];
      else v = mw[(size_t)(row - 2 * D) * 3 * D + col];
      wp[j] = v;
      continue;
    }
    j -= nwp;
    if (j < 3 * D) { bp[j] = j < D ? qb[j] * inv : 0.f; continue; }
    j -= 3 * D;
    if (j < (int64_t)D * D) {
      const int row = (int)(j / D);
      w1o[j] = w1[j] * (g[row] / sqrtf(var[row] + 1e-5f));
      continue;
    }
    j -= (int64_t)D * D;
    if (j < D) { b1o[j] = (b1[j] - mean[j]) * (g[j] / sqrtf(var[j] + 1e-5f)) + beta[j]; continue; }
    j -= D;
    if (j < (int64_t)D * D) { w2o[j] = w2[j]; continue; }
    j -= (int64_t)D * D;
    b2o[j] = b2[j];
  }
}

__global__ void fold_mp_kernel(int D, int T, const float* __restrict__ tw, const float* __restrict__ tb,
                               const float* __restrict__ sw, const float* __restrict__ sb,
                               const float* __restrict__ vhw, const float* __restrict__ vhb,
                               const float* __restrict__ vxw, const float* __restrict__ vxb,
                               const float* __restrict__ basis, float* __restrict__ type_tab,
                               float* __restrict__ basis_o, float* __restrict__ ws, float* __restrict__ bs,
                               float* __restrict__ vcat, float* __restrict__ vbias) {
  const int Dh = D / 2;
  const int64_t n0 = (int64_t)T * Dh, n1 = Dh, n2 = (int64_t)Dh * Dh, n3 = Dh, n4 = (int64_t)D * 2 * D, n5 = D;
  const int64_t total = n0 + n1 + n2 + n3 + n4 + n5;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t j = i;
    if (j < n0) { const int t = (int)(j / Dh), f = (int)(j % Dh); type_tab[j] = gelu_tanh(tw[(size_t)f * T + t] + tb[f]); continue; }
    j -= n0;
    if (j < n1) { basis_o[j] = basis[j]; continue; }
    j -= n1;
    if (j < n2) { ws[j] = sw[j]; continue; }
    j -= n2;
    if (j < n3) { bs[j] = sb[j]; continue; }
    j -= n3;
    if (j < n4) {
      const int row = (int)(j / (2 * D)), col = (int)(j % (2 * D));
      vcat[j] = col < D ? vhw[(size_t)row * D + col] : vxw[(size_t)row * D + (col - D)];
      continue;
    }
    j -= n4;
    vbias[j] = vhb[j] + vxb[j];
  }
}

// projection weight [3D, 2D] (+bias [3D]) -> rows regrouped as [3][H][DP] with zero rows in the pads
__global__ void fold_pad_projection_kernel(int D, int H, int DP, const float* __restrict__ wp, const float* __restrict__ bp,
                                           float* __restrict__ wph, float* __restrict__ bph) {
  const int d = D / H, K = 2 * D;
  const int64_t rows = (int64_t)3 * H * DP;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows * (K + 1); i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (K + 1);
    const int c = (int)(i % (K + 1));
    const int j = (int)(r % DP), slab = (int)(r / DP);  // slab = which*H + h
    const int src_row = (slab / H) * D + (slab % H) * d + j;
    const bool ok = j < d;
    if (c < K) wph[r * K + c] = ok ? wp[(size_t)src_row * K + c] : 0.f;
    else bph[r] = ok ? bp[src_row] : 0.f;
  }
}

// fast projection operands: wps[r, :] = [wph[r, 0:D] | wph[r, D+Dh:2D] | 0-pad]  and
// tbias[t, r] = bph[r] + sum_j wph[r, D + j] * type_tab[t, j]     (the type-embedding half of node_feature_extra, :65-66,:86)
__global__ void fold_type_bias_kernel(int rows, int D, int T, int KS, const float* __restrict__ wph, const float* __restrict__ bph,
                                      const float* __restrict__ type_tab, float* __restrict__ wps, float* __restrict__ tbias) {
  const int Dh = D / 2, K = 2 * D;
  const int64_t n_w = (int64_t)rows * KS, n_b = (int64_t)T * rows;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_w + n_b; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n_w) {
      const int64_t r = i / KS;
      const int c = (int)(i % KS);
      wps[i] = c < D ? wph[r * K + c] : (c < D + Dh ? wph[r * K + D + Dh + (c - D)] : 0.f);
    } else {
      const int64_t j = i - n_w;
      const int t = (int)(j / rows);
      const int64_t r = j % rows;
      float acc = 0.f;
      for (int q = 0; q < Dh; ++q) acc = fmaf(wph[r * K + D + q], type_tab[(size_t)t * Dh + q], acc);
      tbias[j] = bph[r] + acc;
    }
  }
}

// [C, D] -> head-major zero-padded [H, C, DP]
__global__ void fold_head_major_kernel(int C, int D, int H, int DP, const float* __restrict__ ke,
                                       const float* __restrict__ me, float* __restrict__ keh,
                                       float* __restrict__ meh) {
  const int d = D / H;
  const int64_t total = (int64_t)H * C * DP;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % DP);
    const int c = (int)((i / DP) % C);
    const int h = (int)(i / ((int64_t)DP * C));
    const bool ok = j < d;
    keh[i] = ok ? ke[(size_t)c * D + h * d + j] : 0.f;
    meh[i] = ok ? me[(size_t)c * D + h * d + j] : 0.f;
  }
}

}  // namespace
}  // namespace qagnn

using namespace qagnn;

static int32_t check_shape(const qagnn_shape* s) {
  if (!s) return QAGNN_ERR_INVALID_ARGUMENT;
  if (s->N <= 0 || s->E < 0 || s->D <= 0 || s->H <= 0 || s->T <= 0 || s->R <= 0 || s->k < 0)
    return QAGNN_ERR_INVALID_ARGUMENT;
  if (s->D % s->H != 0 || s->D % 2 != 0) return QAGNN_ERR_INVALID_ARGUMENT;  // modeling_qagnn.py:391,399
  if (s->N + s->E >= (int64_t)1 << 31) return QAGNN_ERR_INVALID_ARGUMENT;
  return QAGNN_OK;
}

extern "C" size_t qagnn_fold_bytes(const qagnn_shape* shape) {
  if (check_shape(shape) != QAGNN_OK) return 0;
  return align_up(make_fold_layout(*shape).total * sizeof(float));
}

extern "C" int32_t qagnn_fold_weights(const qagnn_shape* shape, const qagnn_edge_encoder_params* ee,
                                      const qagnn_layer_params* layers, const qagnn_mp_params* mp, void* folded,
                                      size_t folded_bytes, void* stream) {
  QAGNN_RETURN_IF(check_shape(shape));
  if (!ee || !folded || (shape->k > 0 && !layers)) return QAGNN_ERR_INVALID_ARGUMENT;
  const FoldLayout L = make_fold_layout(*shape);
  if (folded_bytes < L.total * sizeof(float)) return QAGNN_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* f = (float*)folded;
  const int D = shape->D, C = L.C;

  const int64_t nh = (int64_t)C * D;
  fold_edge_hidden_kernel<<<(unsigned)((nh + 255) / 256), 256, 0, st>>>(C, D, shape->T, shape->R, ee->lin0_w, ee->lin0_b,
                                                                        ee->bn_w, ee->bn_b, ee->bn_mean, ee->bn_var,
                                                                        f + L.hidden);
  QAGNN_CHECK_LAUNCH();
  QAGNN_RETURN_IF(sgemm_tn(f + L.hidden, D, D, nullptr, 0, 0, ee->lin3_w, D, ee->lin3_b, f + L.tab, D, C, D, ACT_NONE, st));

  for (int l = 0; l < shape->k; ++l) {
    const qagnn_layer_params& p = layers[l];
    float* lb = f + L.layer0 + (size_t)l * L.layer_stride;
    const int64_t total = (int64_t)3 * D * 2 * D + 3 * D + 2 * ((int64_t)D * D + D);
    fold_layer_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        D, shape->H, p.query_w, p.query_b, p.key_w, p.msg_w, p.mlp0_w, p.mlp0_b, p.bn_w, p.bn_b, p.bn_mean, p.bn_var,
        p.mlp3_w, p.mlp3_b, lb + L.wp, lb + L.bp, lb + L.w1, lb + L.b1, lb + L.w2, lb + L.b2);
    QAGNN_CHECK_LAUNCH();
    // edge part of linear_key / linear_msg: columns [2D,3D) of the [D,3D] weights
    QAGNN_RETURN_IF(sgemm_tn(f + L.tab, D, D, nullptr, 0, 0, p.key_w + 2 * D, 3 * D, p.key_b, lb + L.ke, D, C, D, ACT_NONE, st));
    QAGNN_RETURN_IF(sgemm_tn(f + L.tab, D, D, nullptr, 0, 0, p.msg_w + 2 * D, 3 * D, p.msg_b, lb + L.me, D, C, D, ACT_NONE, st));
    const int DP = head_dim_padded(D / shape->H);
    const int64_t nhm = (int64_t)shape->H * C * DP;
    fold_head_major_kernel<<<(unsigned)((nhm + 255) / 256), 256, 0, st>>>(C, D, shape->H, DP, lb + L.ke, lb + L.me,
                                                                          lb + L.keh, lb + L.meh);
    QAGNN_CHECK_LAUNCH();
    if (D % 2 == 0) {  // split-bf16 planes of the dense weights for the tensor-core GEMMs
      QAGNN_RETURN_IF(split_bf16(lb + L.wp, 2 * D, 3 * D, 2 * D, lb + L.wp_hi, lb + L.wp_lo, 2 * D, st));
      const int64_t rows = (int64_t)3 * shape->H * DP;
      fold_pad_projection_kernel<<<(unsigned)((rows * (2 * D + 1) + 255) / 256), 256, 0, st>>>(D, shape->H, DP, lb + L.wp, lb + L.bp,
                                                                                          lb + L.wph, lb + L.bph);
      QAGNN_CHECK_LAUNCH();
      QAGNN_RETURN_IF(split_bf16(lb + L.wph, 2 * D, rows, 2 * D, lb + L.wph_hi, lb + L.wph_lo, 2 * D, st));
      QAGNN_RETURN_IF(split_bf16(lb + L.w1, D, D, D, lb + L.w1_hi, lb + L.w1_lo, D, st));
      QAGNN_RETURN_IF(split_bf16(lb + L.w2, D, D, D, lb + L.w2_hi, lb + L.w2_lo, D, st));
    }
  }
  if (mp) {
    const int Dh = D / 2;
    const int64_t total = (int64_t)shape->T * Dh + Dh + (int64_t)Dh * Dh + Dh + (int64_t)D * 2 * D + D;
    fold_mp_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        D, shape->T, mp->emb_node_type_w, mp->emb_node_type_b, mp->emb_score_w, mp->emb_score_b, mp->vh_w, mp->vh_b,
        mp->vx_w, mp->vx_b, mp->score_basis, f + L.type_tab, f + L.basis, f + L.ws, f + L.bs, f + L.vcat, f + L.vbias);
    QAGNN_CHECK_LAUNCH();
    if (D % 2 == 0) QAGNN_RETURN_IF(split_bf16(f + L.vcat, 2 * D, D, 2 * D, f + L.vcat_hi, f + L.vcat_lo, 2 * D, st));
    if (D % 8 == 0 && Dh % 2 == 0) {  // operands of the fast projection / tensor-core emb_score (qagnn_mp_forward)
      const int KSh = round_up8(Dh), KS = round_up8(D + Dh), DP = head_dim_padded(D / shape->H);
      QAGNN_RETURN_IF(split_bf16(f + L.ws, Dh, Dh, Dh, f + L.ws_hi, f + L.ws_lo, KSh, st));
      const int rows = 3 * shape->H * DP;
      for (int l = 0; l < shape->k; ++l) {
        float* lb = f + L.layer0 + (size_t)l * L.layer_stride;
        const int64_t tot = (int64_t)rows * KS + (int64_t)shape->T * rows;
        fold_type_bias_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(rows, D, shape->T, KS, lb + L.wph, lb + L.bph,
                                                                           f + L.type_tab, lb + L.wps, lb + L.tbias);
        QAGNN_CHECK_LAUNCH();
        QAGNN_RETURN_IF(split_bf16(lb + L.wps, KS, rows, KS, lb + L.wps_hi, lb + L.wps_lo, KS, st));
      }
    }
  }
  return QAGNN_OK;
}
