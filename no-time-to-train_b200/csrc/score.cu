// Pooling contraction, L2 normalisation, prototype preparation, cosine similarity and top-1 label.
//
// Reference: compute_sim_global_avg (matching_baseline_utils.py:869-904) and the top-k section
// (Sam2MatchingBaseline_noAMG.py:602-612).
#include "common.cuh"

namespace nttt {

// ---------------------------------------------------------------------------------------------------
// fp32 tiled GEMM on the CUDA cores (first correct path; the tcgen05 kernel in gemm_tc.cu replaces it
// for the hot contractions).  C[M,N] = A[M,K] * op(B), op(B) = B[K,N] (kBT=false) or B[N,K]^T (kBT=true).
// ---------------------------------------------------------------------------------------------------
template <bool kBT>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C,
             int ldc, int M, int N, int K) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float sA[BK][BM + 4];
  __shared__ float sB[BK][BN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int i = threadIdx.x; i < BM * BK; i += 256) {
      const int m = i / BK, k = i % BK;
      sA[k][m] = (m0 + m < M && k0 + k < K) ? A[(size_t)(m0 + m) * lda + k0 + k] : 0.0f;
    }
    if (kBT) {
      for (int i = threadIdx.x; i < BN * BK; i += 256) {
        const int n = i / BK, k = i % BK;
        sB[k][n] = (n0 + n < N && k0 + k < K) ? B[(size_t)(n0 + n) * ldb + k0 + k] : 0.0f;
      }
    } else {
      for (int i = threadIdx.x; i < BN * BK; i += 256) {
        const int k = i / BN, n = i % BN;
        sB[k][n] = (n0 + n < N && k0 + k < K) ? B[(size_t)(k0 + k) * ldb + n0 + n] : 0.0f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) C[(size_t)m * ldc + n] = acc[i][j];
    }
}

int launch_sgemm(bool b_transposed, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N,
                 int K, cudaStream_t s) {
  if (M <= 0 || N <= 0) return NTTT_OK;
  dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
  if (b_transposed) sgemm_kernel<true><<<grid, 256, 0, s>>>(A, lda, B, ldb, C, ldc, M, N, K);
  else sgemm_kernel<false><<<grid, 256, 0, s>>>(A, lda, B, ldb, C, ldc, M, N, K);
  NTTT_LAUNCH_CHECK();
  return NTTT_OK;
}

// ---------------------------------------------------------------------------------------------------
// rows: x = sums / max(area,1)  (area==0 -> 1, matching_baseline_utils.py:887-888); x /= max(||x||, 1e-12)
// one warp per row.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
normalize_rows_kernel(const float* __restrict__ sums, const int32_t* __restrict__ area, int n, int c,
                      float* __restrict__ out) {
  const int row = blockIdx.x * 8 + warp_id();
  if (row >= n) return;
  const int lane = lane_id();
  float denom = 1.0f;
  if (area) { const int a = area[row]; denom = a == 0 ? 1.0f : (float)a; }
  const float* src = sums + (size_t)row * c;
  float ss = 0.0f;
  for (int i = lane; i < c; i += 32) { const float v = __fdiv_rn(src[i], denom); ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float nrm = fmaxf(sqrtf(ss), 1e-12f);
  float* dst = out + (size_t)row * c;
  for (int i = lane; i < c; i += 32) dst[i] = __fdiv_rn(__fdiv_rn(src[i], denom), nrm);
}

int launch_normalize_rows(const float* sums, const int32_t* area, int n, int c, float* out, cudaStream_t s) {
  if (n <= 0) return NTTT_OK;
  normalize_rows_kernel<<<ceil_div(n, 8), 256, 0, s>>>(sums, area, n, c, out);
  NTTT_LAUNCH_CHECK();
  return NTTT_OK;
}

// prototypes: mean over ALL `shots` slots (unfilled zeros included), then L2-normalise
// (matching_baseline_utils.py:893-894).  One warp per class.
__global__ void __launch_bounds__(256)
proto_prepare_kernel(const float* __restrict__ ins_avg, int n_cls, int shots, int c, float* __restrict__ proto) {
  const int cls = blockIdx.x * 8 + warp_id();
  if (cls >= n_cls) return;
  const int lane = lane_id();
  const float* src = ins_avg + (size_t)cls * shots * c;
  float* dst = proto + (size_t)cls * c;
  float ss = 0.0f;
  for (int i = lane; i < c; i += 32) {
    float acc = 0.0f;
    for (int l = 0; l < shots; ++l) acc += src[(size_t)l * c + i];
    const float m = __fdiv_rn(acc, (float)shots);
    dst[i] = m;
    ss = fmaf(m, m, ss);
  }
  ss = warp_sum(ss);
  const float nrm = fmaxf(sqrtf(ss), 1e-12f);
  __syncwarp();
  for (int i = lane; i < c; i += 32) dst[i] = __fdiv_rn(dst[i], nrm);
}

int launch_proto_prepare(const float* ins_avg, int n_cls, int shots, int c, float* proto, cudaStream_t s) {
  if (n_cls <= 0) return NTTT_OK;
  proto_prepare_kernel<<<ceil_div(n_cls, 8), 256, 0, s>>>(ins_avg, n_cls, shots, c, proto);
  NTTT_LAUNCH_CHECK();
  return NTTT_OK;
}

// top-1 over classes (lowest index on ties == torch.topk/argmax on CPU); when n_cls == 1 the reference's
// `k == n_cls` branch applies: score *= (score > 0.6*score)  (Sam2MatchingBaseline_noAMG.py:606-609).
__global__ void __launch_bounds__(256)
top1_kernel(const float* __restrict__ sim, int ld, int n, int n_cls, float* __restrict__ top_score,
            int32_t* __restrict__ top_label) {
  const int row = blockIdx.x * 8 + warp_id();
  if (row >= n) return;
  const int lane = lane_id();
  const float* src = sim + (size_t)row * ld;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  for (int i = lane; i < n_cls; i += 32) {
    const float v = src[i];
    if (v > best || (v == best && i < arg)) { best = v; arg = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(kFull, best, o);
    const int oa = __shfl_xor_sync(kFull, arg, o);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
  }
  if (lane == 0) {
    if (n_cls == 1) best = best * (float)(best > __fmul_rn(best, 0.6f));
    top_score[row] = best;
    top_label[row] = arg == 0x7fffffff ? 0 : arg;
  }
}

int launch_top1(const float* sim, int ld, int n, int n_cls, float* top_score, int32_t* top_label, cudaStream_t s) {
  if (n <= 0) return NTTT_OK;
  top1_kernel<<<ceil_div(n, 8), 256, 0, s>>>(sim, ld, n, n_cls, top_score, top_label);
  NTTT_LAUNCH_CHECK();
  return NTTT_OK;
}

}  // namespace nttt
