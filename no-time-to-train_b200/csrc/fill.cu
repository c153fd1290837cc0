// Memory-bank fill and post-process on mask-weighted sums instead of raw [n_cls, L, E, C] features.
//
// Reference: forward_fill_memory (Sam2MatchingBaseline_noAMG.py:465-485) stores raw patch features and the
// nearest-resized soft mask per (class, slot); MemoryBank.postprocess (matching_baseline_utils.py:574-599)
// later reduces them to feats_ins_avg / feats_avg.  The reductions are linear in the raw features, so each
// reference shot is reduced on arrival: sum[c,l,:] = sum_e mask[e]*feat[e,:], wsum[c,l] = sum_e mask[e].
#include "common.cuh"

namespace nttt {

// aten nearest: src = min(int(floorf(dst * scale)), in-1), scale = in/out in fp32 (identity / >>1 shortcuts
// agree with the formula for the sizes they cover)
__device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size) {
  if (in_size == out_size) return dst;
  const float scale = __fdiv_rn((float)in_size, (float)out_size);
  return min((int)floorf(__fmul_rn((float)dst, scale)), in_size - 1);
}

__global__ void __launch_bounds__(256)
fill_pool_kernel(const float* __restrict__ feat, const float* __restrict__ soft_mask, int mh, int mw, int eh, int ew,
                 int c, float* __restrict__ sum_slot, float* __restrict__ wsum_slot, float* __restrict__ mask_out) {
  extern __shared__ float s_mask[];  // eh*ew, then 8*32 partials
  const int e_total = eh * ew;
  float* s_part = s_mask + e_total;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  for (int e = tid; e < e_total; e += 256) {
    const int ey = e / ew, ex = e - ey * ew;
    const float m = soft_mask[(size_t)nearest_src(ey, mh, eh) * mw + nearest_src(ex, mw, ew)];
    s_mask[e] = m;
    if (blockIdx.x == 0 && mask_out) mask_out[e] = m;
  }
  __syncthreads();
  const int col = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.0f;
  if (col < c)
    for (int e = threadIdx.y; e < e_total; e += 8) acc = fmaf(s_mask[e], feat[(size_t)e * c + col], acc);
  s_part[threadIdx.y * 32 + threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < c) {
    float t = 0.0f;
    for (int q = 0; q < 8; ++q) t += s_part[q * 32 + threadIdx.x];
    sum_slot[col] += t;
  }
  if (blockIdx.x == 0 && threadIdx.y == 1) {
    float w = 0.0f;
    for (int e = threadIdx.x; e < e_total; e += 32) w += s_mask[e];
    w = warp_sum(w);
    if (threadIdx.x == 0) wsum_slot[0] += w;
  }
}

int launch_fill_pool(const float* feat, const float* soft_mask, int mh, int mw, int eh, int ew, int c, float* sum_slot,
                     float* wsum_slot, float* mask_out, cudaStream_t s) {
  const size_t smem = sizeof(float) * ((size_t)eh * ew + 256);
  if (smem > 200 * 1024) return NTTT_EUNSUPPORTED;
  if (smem > 48 * 1024)
    NTTT_CUDA(cudaFuncSetAttribute(fill_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fill_pool_kernel<<<ceil_div(c, 32), dim3(32, 8), smem, s>>>(feat, soft_mask, mh, mw, eh, ew, c, sum_slot, wsum_slot,
                                                             mask_out);
  NTTT_LAUNCH_CHECK();
  return NTTT_OK;
}

// feats_ins_avg[c,l,:] = sum[c,l,:] / (wsum[c,l] or 1);  feats_avg[c,:] = sum_l sum[c,l,:] / (sum_l wsum[c,l] or 1)
__global__ void __launch_bounds__(256)
fill_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ wsum, int n_cls, int shots, int c,
                     float* __restrict__ ins_avg, float* __restrict__ avg) {
  const int cls = blockIdx.x;
  float wall = 0.0f;
  for (int l = 0; l < shots; ++l) wall += wsum[cls * shots + l];
  if (wall == 0.0f) wall = 1.0f;
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    float tot = 0.0f;
    for (int l = 0; l < shots; ++l) {
      const float v = sum[((size_t)cls * shots + l) * c + i];
      float w = wsum[cls * shots + l];
      if (w == 0.0f) w = 1.0f;
      ins_avg[((size_t)cls * shots + l) * c + i] = __fdiv_rn(v, w);
      tot += v;
    }
    avg[(size_t)cls * c + i] = __fdiv_rn(tot, wall);
  }
}

int launch_fill_finalize(const float* sum, const float* wsum, int n_cls, int shots, int c, float* ins_avg, float* avg,
                         cudaStream_t s) {
  if (n_cls <= 0) return NTTT_OK;
  fill_finalize_kernel<<<n_cls, 256, 0, s>>>(sum, wsum, n_cls, shots, c, ins_avg, avg);
  NTTT_LAUNCH_CHECK();
  return NTTT_OK;
}

}  // namespace nttt
