// Pooling contraction, L2 normalisation, prototype preparation, cosine similarity and top-1 label.
//
// Reference: compute_sim_global_avg (matching_baseline_utils.py:869-904) and the top-k section
// (Sam2MatchingBaseline_noAMG.py:602-612).
#include "common.cuh"

namespace nttt {

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
