// Full-resolution mask pass: antialiased bilinear resize of the selected low-res logits to the original
// image size, strict > 0, bit-pack, area and box — without materialising the fp32 [K,H,W] tensor.
//
// Reference: Sam2MatchingBaseline_noAMG.py:657-665 (F.interpolate(..., antialias=True) > 0,
// batched_mask_to_box).  Arithmetic: horizontal pass on each contributing input row, then vertical
// pass, each `acc = s0*w0; acc = fma(s_j, w_j, acc)` in fp32 — the association of aten's
// upsample_gen2d_aa kernel (ATen/native/cuda/UpSample.cuh interpolate_aa_single_dim).
//
// Two exact shortcuts keep the kernel off the instruction roofline:
//   1. rect bound: an output sample whose footprint lies outside the low-res box sees only
//      non-positive logits and non-negative weights, so it is not > 0.  Only the rect is computed/stored.
//   2. uniform footprints: if every low-res bit under the footprint of a 32-pixel output word is 0 the
//      word is 0; if every bit is 1 (and the mask's positives are finite, flags bit0) the word is all
//      ones.  Only words whose footprint crosses the mask boundary evaluate the interpolation.
#include "common.cuh"

namespace nttt {

constexpr int kUpThreads = 256;
constexpr int kUpSplit = 4;  // CTAs per mask (rows interleaved)

struct UpTables {
  const int32_t* xmin; const int32_t* xsize; const float* wx; int tx;
  const int32_t* ymin; const int32_t* ysize; const float* wy; int ty;
  const int32_t* x_tlo; const int32_t* x_tlen;  // input col -> output range
  const int32_t* y_tlo; const int32_t* y_tlen;  // input row -> output range
  const int32_t* y_grp_of; const int32_t* y_grp_start;  // output rows grouped by identical input span
};

// acc = s0*w0; acc = fma(s_j, w_j, acc)
__device__ __forceinline__ float aa_dot(const float* __restrict__ src, int stride, const float* __restrict__ w, int n) {
  float acc = __fmul_rn(__ldg(src), __ldg(w));
  for (int j = 1; j < n; ++j) acc = __fmaf_rn(__ldg(src + (size_t)j * stride), __ldg(w + j), acc);
  return acc;
}

// scratch per mask (zero-initialised by the launcher): {area, maxx+1, maxy+1, BIG-minx, BIG-miny, done}
constexpr int kScratchInts = 8;
constexpr int kBig = 1 << 30;
constexpr int kTapsReg = 4;  // footprints of up to 4 input rows keep their horizontal-pass values in registers

// Work decomposition: the unit of work is one STRIP = one 32-pixel word column of one selected mask, walked top
// to bottom by one warp with lane = pixel.  Everything that depends only on the column (input span, horizontal
// weights, footprint bit masks) is loaded once per strip and lives in registers; per row group the warp tests the
// footprint bits, and only if they are mixed loads the 2-3 taps x 2-4 input rows, runs the horizontal pass once
// and the vertical pass per output row, each output word being one ballot.  A single-CTA pre-kernel resolves
// every mask's rect once and builds the exclusive prefix of the strip counts; the main kernel is a fixed grid
// whose warps take strips item = warp_id, warp_id + n_warps, ... and locate (mask, word) by binary search.
struct UpMeta {
  int src;      // index into logits / bits_lr
  int r0, r1;   // output rows [r0, r1)
  int w0, w1;   // output words [w0, w1)
  int g0, g1;   // row groups [g0, g1)
  int safe;     // flags bit0
};

__global__ void __launch_bounds__(1024)
upsample_meta_kernel(const uint32_t* __restrict__ bits_lr, const int32_t* __restrict__ box_lr,
                     const int32_t* __restrict__ flags_lr, int ih, int iw, const int32_t* __restrict__ sel,
                     const int32_t* __restrict__ n_sel, int max_sel, UpTables t, UpMeta* __restrict__ meta,
                     int32_t* __restrict__ rect, int32_t* __restrict__ scratch, int32_t* __restrict__ prefix,
                     int32_t* __restrict__ area_full, int32_t* __restrict__ box_full) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int lane = lane_id(), warp = warp_id();
  const int nsel = min(*n_sel, max_sel);
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < max_sel; base += 1024) {
    const int k = base + threadIdx.x;
    int cnt = 0;
    if (k < max_sel) {
#pragma unroll
      for (int q = 0; q < kScratchInts; ++q) scratch[(size_t)k * kScratchInts + q] = 0;
    }
    if (k < nsel) {
      UpMeta m;
      m.src = sel[k];
      const int4 b = reinterpret_cast<const int4*>(box_lr)[m.src];
      // empty low-res mask <=> box all zero AND bit (0,0) clear
      const bool lr_empty = (b.x | b.y | b.z | b.w) == 0 && (bits_lr[(size_t)m.src * ih * (iw >> 5)] & 1u) == 0;
      m.r0 = m.r1 = m.w0 = m.w1 = m.g0 = m.g1 = 0;
      if (!lr_empty) {
        m.r0 = t.y_tlo[b.y];
        m.r1 = t.y_tlo[b.w] + t.y_tlen[b.w];
        const int c0 = t.x_tlo[b.x];
        const int c1 = t.x_tlo[b.z] + t.x_tlen[b.z];
        m.w0 = c0 >> 5;
        m.w1 = (c1 + 31) >> 5;
        m.g0 = t.y_grp_of[m.r0];
        m.g1 = t.y_grp_of[m.r1 - 1] + 1;
      }
      m.safe = flags_lr[m.src] & 1;
      meta[k] = m;
      cnt = (m.r1 > m.r0) ? (m.w1 - m.w0) : 0;  // strips of this mask
      reinterpret_cast<int4*>(rect)[k] = make_int4(m.r0, m.r1, m.w0, m.w1);
      if (cnt == 0) {  // nothing to compute: publish the empty statistics here
        area_full[k] = 0;
        reinterpret_cast<int4*>(box_full)[k] = make_int4(0, 0, 0, 0);
      }
    }
    // block-wide exclusive scan of cnt
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, wi, o);
        if (lane >= o) wi += v;
      }
      s_warp[lane] = wi - w;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    const int excl = incl - cnt + s_warp[warp] + s_carry;
    if (k < max_sel) prefix[k] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + cnt;
    __syncthreads();
  }
  if (threadIdx.x == 0) prefix[max_sel] = s_carry;
}

__global__ void __launch_bounds__(kUpThreads, 4)
upsample_pack_kernel(const float* __restrict__ logits, const uint32_t* __restrict__ bits_lr,
                     const UpMeta* __restrict__ meta, const int32_t* __restrict__ prefix, int ih, int iw, int max_sel,
                     int oh, int ow, UpTables t, uint32_t* __restrict__ bits_full, int32_t* __restrict__ area_full,
                     int32_t* __restrict__ box_full, int32_t* __restrict__ scratch) {
  const int lr_wpr = iw >> 5;
  const int ow_words = (ow + 31) >> 5;
  const int lane = lane_id();
  constexpr int kWarps = kUpThreads / 32;
  const int total = prefix[max_sel];
  const int n_warps = gridDim.x * kWarps;
  for (int item = blockIdx.x * kWarps + warp_id(); item < total; item += n_warps) {
    // (mask, word) of this strip: last k with prefix[k] <= item
    int klo = 0, khi = max_sel;
    while (khi - klo > 1) {
      const int mid = (klo + khi) >> 1;
      if (__ldg(prefix + mid) <= item) klo = mid; else khi = mid;
    }
    const int k = klo;
    const UpMeta mt = meta[k];
    const int wi = mt.w0 + (item - __ldg(prefix + k));
    const int r0 = mt.r0, r1 = mt.r1;
    const bool safe = mt.safe != 0;
    const float* src = logits + (size_t)mt.src * ih * iw;
    const uint32_t* lr = bits_lr + (size_t)mt.src * ih * lr_wpr;
    uint32_t* dst = bits_full + (size_t)k * oh * ow_words + wi;

    // ---- column constants (registers for the whole strip) ----
    const int x = (wi << 5) + lane;
    const bool inb = x < ow;
    const int xc = inb ? x : ow - 1;
    const int cx = t.xmin[xc], cs = t.xsize[xc];
    const float* wxp = t.wx + (size_t)xc * t.tx;
    const bool fast_x = t.tx <= 3;
    float wx0 = 0.f, wx1 = 0.f, wx2 = 0.f;
    if (fast_x) {
      wx0 = __ldg(wxp);
      if (cs > 1) wx1 = __ldg(wxp + 1);
      if (cs > 2) wx2 = __ldg(wxp + 2);
    }
    const int x0 = wi << 5, x1 = min(x0 + 31, ow - 1);
    const uint32_t valid = (x1 - x0 == 31) ? 0xffffffffu : ((1u << (x1 - x0 + 1)) - 1u);
    const int c0 = t.xmin[x0];
    const int c1 = t.xmin[x1] + t.xsize[x1];  // exclusive: low-res columns [c0, c1) feed this word
    const int cw0 = c0 >> 5, cw1 = (c1 - 1) >> 5;
    const float* pcol = src + cx;

    int area = 0, miny = kBig, maxy = -1;
    uint32_t colbits = 0;
    for (int g = mt.g0; g < mt.g1; ++g) {
      // rows [ya, yb) of this group share the input rows [ry0, ry0 + rys)
      const int ya = max(__ldg(t.y_grp_start + g), r0), yb = min(__ldg(t.y_grp_start + g + 1), r1);
      const int nrows = yb - ya;
      const int ry0 = __ldg(t.ymin + ya), rys = __ldg(t.ysize + ya);
      // footprint test: lane r looks at input row ry0 + r (rys <= 32 always holds for supported scales)
      bool z = true, o = true;
      for (int rr = lane; rr < rys; rr += 32) {
        const uint32_t* row = lr + (size_t)(ry0 + rr) * lr_wpr;
        for (int cw = cw0; cw <= cw1; ++cw) {
          const int lo = max(c0 - (cw << 5), 0), hi = min(c1 - (cw << 5), 32);
          const uint32_t m = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
          const uint32_t v = __ldg(row + cw) & m;
          z = z && (v == 0);
          o = o && (v == m);
        }
      }
      const bool all0 = __all_sync(kFull, z), all1 = __all_sync(kFull, o);
      uint32_t words[kGrpMax];
#pragma unroll
      for (int j = 0; j < kGrpMax; ++j) words[j] = 0;
      if (all0) {
        // stays zero
      } else if (all1 && safe) {
#pragma unroll
        for (int j = 0; j < kGrpMax; ++j) words[j] = valid;
      } else if (rys <= kTapsReg && fast_x) {
        // horizontal pass once per group (taps beyond cs are never read), vertical pass per output row
        float T[kTapsReg];
#pragma unroll
        for (int r = 0; r < kTapsReg; ++r) {
          float acc = 0.0f;
          if (r < rys && inb) {
            const float* p = pcol + (size_t)(ry0 + r) * iw;
            acc = __fmul_rn(__ldg(p), wx0);
            if (cs > 1) acc = __fmaf_rn(__ldg(p + 1), wx1, acc);
            if (cs > 2) acc = __fmaf_rn(__ldg(p + 2), wx2, acc);
          }
          T[r] = acc;
        }
#pragma unroll
        for (int j = 0; j < kGrpMax; ++j) {
          if (j < nrows) {
            const float* wy = t.wy + (size_t)(ya + j) * t.ty;
            float acc = __fmul_rn(T[0], __ldg(wy));
#pragma unroll
            for (int r = 1; r < kTapsReg; ++r)
              if (r < rys) acc = __fmaf_rn(T[r], __ldg(wy + r), acc);
            words[j] = __ballot_sync(kFull, inb && acc > 0.0f);
          }
        }
      } else {
        // long footprints (down-scaling): evaluate each output row directly
#pragma unroll
        for (int j = 0; j < kGrpMax; ++j) {
          if (j < nrows) {
            const float* wy = t.wy + (size_t)(ya + j) * t.ty;
            float acc = 0.0f;
            if (inb) {
              const float* p = pcol + (size_t)ry0 * iw;
              acc = __fmul_rn(aa_dot(p, 1, wxp, cs), __ldg(wy));
              for (int r = 1; r < rys; ++r) acc = __fmaf_rn(aa_dot(p + (size_t)r * iw, 1, wxp, cs), __ldg(wy + r), acc);
            }
            words[j] = __ballot_sync(kFull, inb && acc > 0.0f);
          }
        }
      }
      // the words are warp-uniform: lane j stores row j
#pragma unroll
      for (int j = 0; j < kGrpMax; ++j) {
        if (j < nrows) {
          if (lane == j) dst[(size_t)(ya + j) * ow_words] = words[j];
          if (words[j]) {
            area += __popc(words[j]);
            colbits |= words[j];
            miny = min(miny, ya + j);
            maxy = max(maxy, ya + j);
          }
        }
      }
    }
    // per-mask statistics (all values are warp-uniform): atomics into the scratch; the last strip publishes
    if (lane == 0) {
      int32_t* sc = scratch + (size_t)k * kScratchInts;
      if (area > 0) {
        atomicAdd(&sc[0], area);
        atomicMax(&sc[1], (wi << 5) + 31 - __clz(colbits) + 1);
        atomicMax(&sc[2], maxy + 1);
        atomicMax(&sc[3], kBig - ((wi << 5) + __ffs(colbits) - 1));
        atomicMax(&sc[4], kBig - miny);
      }
      __threadfence();
      const int prev = atomicAdd(&sc[5], 1);
      if (prev == (mt.w1 - mt.w0) - 1) {
        __threadfence();
        const int a = atomicAdd(&sc[0], 0);
        const int mx1 = atomicMax(&sc[1], 0), my1 = atomicMax(&sc[2], 0);
        const int bx = atomicMax(&sc[3], 0), by = atomicMax(&sc[4], 0);
        area_full[k] = a;
        int4 o4 = make_int4(0, 0, 0, 0);
        if (a > 0) o4 = make_int4(kBig - bx, kBig - by, mx1 - 1, my1 - 1);
        reinterpret_cast<int4*>(box_full)[k] = o4;
      }
    }
  }
}

int launch_upsample_pack(const AxisTable& tx, const AxisTable& ty, const float* logits, const uint32_t* bits_lr,
                         const int32_t* box_lr, const int32_t* flags_lr, int ih, int iw, const int32_t* sel,
                         const int32_t* n_sel, int max_sel, int oh, int ow, uint32_t* bits_full, int32_t* rect,
                         int32_t* area_full, int32_t* box_full, int32_t* scratch, cudaStream_t s) {
  if (max_sel <= 0) return NTTT_OK;
  if (iw % 32 != 0) return NTTT_EUNSUPPORTED;
  UpTables t{tx.xmin, tx.xsize, tx.w, tx.taps, ty.xmin, ty.xsize, ty.w, ty.taps, tx.t_lo, tx.t_len, ty.t_lo, ty.t_len,
             ty.grp_of, ty.grp_start};
  UpMeta* meta = reinterpret_cast<UpMeta*>(scratch + (size_t)kScratchInts * max_sel);
  int32_t* prefix = reinterpret_cast<int32_t*>(meta + max_sel);
  upsample_meta_kernel<<<1, 1024, 0, s>>>(bits_lr, box_lr, flags_lr, ih, iw, sel, n_sel, max_sel, t, meta, rect, scratch,
                                          prefix, area_full, box_full);
  NTTT_LAUNCH_CHECK();
  upsample_pack_kernel<<<148 * 4, kUpThreads, 0, s>>>(logits, bits_lr, meta, prefix, ih, iw, max_sel, oh, ow, t,
                                                      bits_full, area_full, box_full, scratch);
  NTTT_LAUNCH_CHECK();
  return NTTT_OK;
}
size_t upsample_scratch_bytes(int max_sel) {
  return (sizeof(int32_t) * kScratchInts + sizeof(UpMeta)) * (size_t)max_sel + sizeof(int32_t) * ((size_t)max_sel + 1);
}

// ---------------------------------------------------------------------------------------------------
// unpack: packed words -> torch.bool bytes.  `index` (nullable) maps output slot j -> packed mask k.
// HBM-bound: oh*ow bytes written per mask.  One thread expands one 32-pixel word into two 128-bit stores
// (nibble -> 4 bytes is one multiply and one mask); rows outside the mask's rect are zero-filled unread.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t nibble_to_bytes(uint32_t nb) { return (nb * 0x00204081u) & 0x01010101u; }

constexpr int kUnpackRows = 32;   // output rows per CTA
constexpr int kUnpackUnroll = 4;  // words per thread in flight

__global__ void __launch_bounds__(256)
unpack_masks_kernel(const uint32_t* __restrict__ bits_full, const int32_t* __restrict__ rect,
                    const int32_t* __restrict__ index, const int32_t* __restrict__ count, int max_count, int oh, int ow,
                    uint8_t* __restrict__ out) {
  const int j = blockIdx.y;
  if (j >= min(*count, max_count)) return;
  const int k = index ? index[j] : j;
  const int ow_words = (ow + 31) >> 5;
  const int4 rc = reinterpret_cast<const int4*>(rect)[k];
  const uint32_t* src = bits_full + (size_t)k * oh * ow_words;
  uint8_t* dst = out + (size_t)j * oh * ow;
  const int y0 = blockIdx.x * kUnpackRows;
  const int rows = min(kUnpackRows, oh - y0);
  const bool vec = (ow & 15) == 0;  // every 16-pixel half word starts 16-byte aligned
  const int total = rows * ow_words;
  for (int base = 0; base < total; base += 256 * kUnpackUnroll) {
    uint32_t word[kUnpackUnroll];
    int yy[kUnpackUnroll], ww[kUnpackUnroll];
#pragma unroll
    for (int u = 0; u < kUnpackUnroll; ++u) {  // all loads first
      const int it = base + u * 256 + threadIdx.x;
      const int ry = it / ow_words;
      yy[u] = y0 + ry;
      ww[u] = it - ry * ow_words;
      word[u] = 0;
      if (it < total && yy[u] >= rc.x && yy[u] < rc.y && ww[u] >= rc.z && ww[u] < rc.w)
        word[u] = __ldg(src + (size_t)yy[u] * ow_words + ww[u]);
    }
#pragma unroll
    for (int u = 0; u < kUnpackUnroll; ++u) {
      const int it = base + u * 256 + threadIdx.x;
      if (it >= total) continue;
      const int x = ww[u] << 5;
      const uint32_t w = word[u];
      uint8_t* p = dst + (size_t)yy[u] * ow + x;
      if (vec) {
        *reinterpret_cast<uint4*>(p) = make_uint4(nibble_to_bytes(w & 15u), nibble_to_bytes((w >> 4) & 15u),
                                                   nibble_to_bytes((w >> 8) & 15u), nibble_to_bytes((w >> 12) & 15u));
        if (x + 16 < ow)
          *reinterpret_cast<uint4*>(p + 16) = make_uint4(nibble_to_bytes((w >> 16) & 15u), nibble_to_bytes((w >> 20) & 15u),
                                                        nibble_to_bytes((w >> 24) & 15u), nibble_to_bytes(w >> 28));
      } else {
        for (int q = 0; q < 32 && x + q < ow; ++q) p[q] = (w >> q) & 1u;
      }
    }
  }
}

int launch_unpack(const uint32_t* bits_full, const int32_t* rect, const int32_t* index, const int32_t* count,
                  int max_count, int oh, int ow, uint8_t* out, cudaStream_t s) {
  if (max_count <= 0) return NTTT_OK;
  dim3 grid(ceil_div(oh, kUnpackRows), max_count);
  unpack_masks_kernel<<<grid, 256, 0, s>>>(bits_full, rect, index, count, max_count, oh, ow, out);
  NTTT_LAUNCH_CHECK();
  return NTTT_OK;
}

}  // namespace nttt
