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

// One CTA per (128-column strip, shot): 8 warps split the E patches, a lane owns 4 consecutive columns (one 128-bit
// load per patch row: a warp reads 512 contiguous bytes).  blockIdx.y = shot, so a batch of reference shots is ONE
// launch of b * ceil(c/128) CTAs.  The reduction order is fixed (patches e = w, w+8, ... per warp with fma, then the
// 8 warp partials in order), so a shot pools to the same bits whether it arrives alone or in a batch.
//   accumulate != 0: sums[b,:] += ..., wsums[b] += ..., mask_out[b,:] += ... (the single-shot entry writes straight
//                    into a bank slot, as `feats[c, slot] += f; masks[c, slot] += m` at :482-484)
//   accumulate == 0: plain stores into staging rows
__global__ void __launch_bounds__(256)
fill_pool_kernel(const float* __restrict__ feat, const float* __restrict__ soft_mask, int mh, int mw, int eh, int ew,
                 int c, float* __restrict__ sums, float* __restrict__ wsums, float* __restrict__ mask_out,
                 int accumulate) {
  extern __shared__ __align__(16) float s_mask[];  // eh*ew (padded to 4 floats), then 8 * 128 partials
  const int e_total = eh * ew;
  float* s_part = s_mask + ((e_total + 3) & ~3);  // read back as float4
  const int shot = blockIdx.y;
  feat += (size_t)shot * e_total * c;
  soft_mask += (size_t)shot * mh * mw;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < e_total; e += 256) {
    const int ey = e / ew, ex = e - ey * ew;
    const float m = soft_mask[(size_t)nearest_src(ey, mh, eh) * mw + nearest_src(ex, mw, ew)];
    s_mask[e] = m;
    if (blockIdx.x == 0 && mask_out) {
      float* mo = mask_out + (size_t)shot * e_total + e;
      *mo = accumulate ? *mo + m : m;
    }
  }
  __syncthreads();
  const int col = blockIdx.x * 128 + lane * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col + 3 < c && (c & 3) == 0) {
    const float4* src = reinterpret_cast<const float4*>(feat + col);
    const size_t stride4 = (size_t)c >> 2;
#pragma unroll 4
    for (int e = warp; e < e_total; e += 8) {
      const float m = s_mask[e];
      const float4 v = __ldg(src + (size_t)e * stride4);
      acc.x = fmaf(m, v.x, acc.x); acc.y = fmaf(m, v.y, acc.y);
      acc.z = fmaf(m, v.z, acc.z); acc.w = fmaf(m, v.w, acc.w);
    }
  } else {
    for (int e = warp; e < e_total; e += 8) {
      const float m = s_mask[e];
      const float* row = feat + (size_t)e * c;
      if (col + 0 < c) acc.x = fmaf(m, row[col + 0], acc.x);
      if (col + 1 < c) acc.y = fmaf(m, row[col + 1], acc.y);
      if (col + 2 < c) acc.z = fmaf(m, row[col + 2], acc.z);
      if (col + 3 < c) acc.w = fmaf(m, row[col + 3], acc.w);
    }
  }
  reinterpret_cast<float4*>(s_part)[warp * 32 + lane] = acc;
  __syncthreads();
  if (tid < 128) {
    const int oc = blockIdx.x * 128 + tid;
    if (oc < c) {
      float t = 0.0f;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += s_part[q * 128 + tid];
      float* dst = sums + (size_t)shot * c + oc;
      *dst = accumulate ? *dst + t : t;
    }
  }
  if (blockIdx.x == 0 && warp == 7) {
    float w = 0.0f;
    for (int e = lane; e < e_total; e += 32) w += s_mask[e];
    w = warp_sum(w);
    if (lane == 0) wsums[shot] = accumulate ? wsums[shot] + w : w;
  }
}

// b shots: feat [b, eh*ew, c], soft_mask [b, mh, mw] -> sums [b, c], wsums [b], mask_out [b, eh*ew] (nullable)
int launch_fill_pool(const float* feat, const float* soft_mask, int b, int mh, int mw, int eh, int ew, int c, float* sums,
                     float* wsums, float* mask_out, int accumulate, cudaStream_t s) {
  if (b <= 0) return NTTT_OK;
  if (b > 65535) return NTTT_EUNSUPPORTED;
  const size_t smem = sizeof(float) * ((((size_t)eh * ew + 3) & ~(size_t)3) + 8 * 128);
  if (smem > 200 * 1024) return NTTT_EUNSUPPORTED;
  if (smem > 48 * 1024)
    NTTT_CUDA(set_dyn_smem(fill_pool_kernel, (int)smem));
  fill_pool_kernel<<<dim3(ceil_div(c, 128), b), 256, smem, s>>>(feat, soft_mask, mh, mw, eh, ew, c, sums, wsums, mask_out,
                                                                accumulate);
  NTTT_LAUNCH_CHECK();
  return NTTT_OK;
}

// Staged rows -> bank slots (the slot loop of forward_fill_memory, Sam2MatchingBaseline_noAMG.py:478-485, for n shots
// at once): dst = slot[i] (= class * L + position; negative: skip)
//   feats_sum[dst,:] += sums[i,:];  mask_sum[dst] += wsums[i];  masks[dst,:] += masks_lowres[i,:]
// Every destination slot has one writer per launch (slots are unique), so there are no atomics and no order
// dependence.  grid (n), block 256.
__global__ void __launch_bounds__(256)
fill_scatter_kernel(const float* __restrict__ sums, const float* __restrict__ wsums,
                    const float* __restrict__ masks_lowres, const int32_t* __restrict__ slot, int c, int e,
                    float* __restrict__ feats_sum, float* __restrict__ mask_sum, float* __restrict__ masks) {
  const int i = blockIdx.x;
  const int dst = slot[i];
  if (dst < 0) return;
  const float* src = sums + (size_t)i * c;
  float* out = feats_sum + (size_t)dst * c;
  for (int k = threadIdx.x; k < c; k += 256) out[k] += src[k];
  if (masks && masks_lowres) {
    const float* ms = masks_lowres + (size_t)i * e;
    float* mo = masks + (size_t)dst * e;
    for (int k = threadIdx.x; k < e; k += 256) mo[k] += ms[k];
  }
  if (threadIdx.x == 0) mask_sum[dst] += wsums[i];
}

int launch_fill_scatter(const float* sums, const float* wsums, const float* masks_lowres, const int32_t* slot, int n,
                        int c, int e, float* feats_sum, float* mask_sum, float* masks, cudaStream_t s) {
  if (n <= 0) return NTTT_OK;
  fill_scatter_kernel<<<n, 256, 0, s>>>(sums, wsums, masks_lowres, slot, c, e, feats_sum, mask_sum, masks);
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
