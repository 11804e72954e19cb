// FP32 (FFMA) "TN" GEMM with fused bias/activation epilogue:
//     C[M,N] = act( [A1 | A2] @ W^T + bias )
// A1 [M,K1], A2 [M,K2] and W [N,K1+K2] are all K-contiguous, which is exactly how nn.Linear stores
// its weight, so the reference's parameters are used in place.  The two-source A operand implements
// the reference's torch.cat([x, node_feature_extra], dim=1) (modeling_qagnn.py:440) and the
// Vh(H)+Vx(X) sum (:92) without materialising the concatenation.
//
// This is the exact-fp32 dense path (bit-comparable to a CPU fp32 GEMM up to summation order); the
// tensor-core split-precision path (gemm_tc.cu) replaces it for the large per-layer projections.
#include "common.cuh"

namespace qagnn {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, NT = 256;
constexpr int LDS = BM + 4;

template <bool VEC>
__device__ __forceinline__ void load_tile(const float* __restrict__ A1, int lda1, int K1, const float* __restrict__ A2,
                                          int lda2, int K, int64_t rows, int64_t row0, int k0, float (&reg)[2][4]) {
  // 128 rows x 16 k: thread t loads float4 #(t%4) of rows t/4 and t/4+64
  const int kq = (threadIdx.x & 3) * 4;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int64_t r = row0 + (threadIdx.x >> 2) + 64 * i;
    const int k = k0 + kq;
    if (VEC) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows && k < K) {
        const float* p = (k < K1) ? (A1 + r * lda1 + k) : (A2 + r * lda2 + (k - K1));
        v = *reinterpret_cast<const float4*>(p);
      }
      reg[i][0] = v.x; reg[i][1] = v.y; reg[i][2] = v.z; reg[i][3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = k + j;
        float v = 0.f;
        if (r < rows && kk < K) v = (kk < K1) ? A1[r * lda1 + kk] : A2[r * lda2 + (kk - K1)];
        reg[i][j] = v;
      }
    }
  }
}

__device__ __forceinline__ void store_tile(float (*S)[LDS], const float (&reg)[2][4]) {
  const int kq = (threadIdx.x & 3) * 4;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = (threadIdx.x >> 2) + 64 * i;
#pragma unroll
    for (int j = 0; j < 4; ++j) S[kq + j][r] = reg[i][j];
  }
}

template <bool VEC, int ACT>
__global__ void __launch_bounds__(NT) sgemm_tn_kernel(const float* __restrict__ A1, int lda1, int K1,
                                                      const float* __restrict__ A2, int lda2, int K2,
                                                      const float* __restrict__ W, int ldw,
                                                      const float* __restrict__ bias, float* __restrict__ C, int ldc,
                                                      int64_t M, int N, int vecC, HeadMajorOut hm) {
  __shared__ __align__(16) float As[2][BK][LDS];
  __shared__ __align__(16) float Ws[2][BK][LDS];
  const int K = K1 + K2;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[2][4], rw[2][4];
  load_tile<VEC>(A1, lda1, K1, A2, lda2, K, M, m0, 0, ra);
  load_tile<VEC>(W, ldw, K, nullptr, 0, K, N, n0, 0, rw);
  store_tile(As[0], ra);
  store_tile(Ws[0], rw);
  __syncthreads();

  const int nk = (K + BK - 1) / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      load_tile<VEC>(A1, lda1, K1, A2, lda2, K, M, m0, (kt + 1) * BK, ra);
      load_tile<VEC>(W, ldw, K, nullptr, 0, K, N, n0, (kt + 1) * BK, rw);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Ws[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Ws[cur][k][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile(As[cur ^ 1], ra);
      store_tile(Ws[cur ^ 1], rw);
    }
    __syncthreads();
  }

  // epilogue: rows {ty*4+i, 64+ty*4+i}, cols {tx*4+j, 64+tx*4+j}
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int c = n0 + jh * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = acc[i][jh * 4 + j];
        if (bias != nullptr && c + j < N) x += bias[c + j];
        if (ACT == ACT_RELU) x = fmaxf(x, 0.f);
        if (ACT == ACT_GELU) x = gelu_tanh(x);
        v[j] = x;
      }
      if (hm.enabled) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cc = c + j;
          if (cc < N) {
            const int which = cc / hm.D, jd = cc % hm.D;
            C[((size_t)(which * hm.H + jd / hm.d) * M + r) * hm.DP + jd % hm.d] = v[j];
          }
        }
      } else if (vecC && c + 3 < N) {
        *reinterpret_cast<float4*>(C + r * ldc + c) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < N) C[r * ldc + c + j] = v[j];
      }
    }
  }
}

template <bool VEC>
int32_t launch(const float* A1, int lda1, int K1, const float* A2, int lda2, int K2, const float* W, int ldw,
               const float* bias, float* C, int ldc, int64_t M, int N, Act act, int vecC, cudaStream_t st,
               HeadMajorOut hm) {
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
  switch (act) {
    case ACT_NONE:
      sgemm_tn_kernel<VEC, ACT_NONE><<<grid, NT, 0, st>>>(A1, lda1, K1, A2, lda2, K2, W, ldw, bias, C, ldc, M, N, vecC, hm);
      break;
    case ACT_RELU:
      sgemm_tn_kernel<VEC, ACT_RELU><<<grid, NT, 0, st>>>(A1, lda1, K1, A2, lda2, K2, W, ldw, bias, C, ldc, M, N, vecC, hm);
      break;
    case ACT_GELU:
      sgemm_tn_kernel<VEC, ACT_GELU><<<grid, NT, 0, st>>>(A1, lda1, K1, A2, lda2, K2, W, ldw, bias, C, ldc, M, N, vecC, hm);
      break;
  }
  QAGNN_CHECK_LAUNCH();
  return QAGNN_OK;
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

int32_t sgemm_tn(const float* A1, int lda1, int K1, const float* A2, int lda2, int K2, const float* W, int ldw,
                 const float* bias, float* C, int ldc, int64_t M, int N, Act act, cudaStream_t st, HeadMajorOut hm) {
  if (M <= 0 || N <= 0) return QAGNN_OK;
  if (!A1 || !W || !C || K1 <= 0 || (K2 > 0 && !A2)) return QAGNN_ERR_INVALID_ARGUMENT;
  if (K2 <= 0) { A2 = nullptr; lda2 = 0; K2 = 0; }
  const bool vec = aligned16(A1) && aligned16(W) && (lda1 % 4 == 0) && (ldw % 4 == 0) && (K1 % 4 == 0) &&
                   (K2 == 0 || (aligned16(A2) && lda2 % 4 == 0 && K2 % 4 == 0));
  const int vecC = aligned16(C) && (ldc % 4 == 0);
  return vec ? launch<true>(A1, lda1, K1, A2, lda2, K2, W, ldw, bias, C, ldc, M, N, act, vecC, st, hm)
             : launch<false>(A1, lda1, K1, A2, lda2, K2, W, ldw, bias, C, ldc, M, N, act, vecC, st, hm);
}

}  // namespace qagnn
