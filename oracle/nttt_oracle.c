/*
 * TEST INFRASTRUCTURE — plain-C restatement of the arithmetic inside the reference's matching stage.
 *
 * This is the oracle, not the product.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load it.  It restates, scalar and single-threaded, what the reference computes
 * through library calls, so that the CUDA path can be checked bit-for-bit on integer/byte outputs:
 *
 *   orc_threshold_stats      lr_masks > 0, areas, boxes, stability counts
 *                            (Sam2MatchingBaseline_noAMG.py:548-549, sam2/utils/amg.py:158-178, 305-348)
 *   orc_aa_*                 aten _upsample_bilinear2d_aa (call sites Sam2MatchingBaseline_noAMG.py:551-558
 *                            and :657-663; arithmetic restated from the published algorithm in torch's
 *                            shipped header ATen/native/cuda/UpSample.cuh — Pillow-style antialiased
 *                            separable resize; torch pinned 2.4.1 by pyproject.toml:56, 2.11.0 in this image)
 *   orc_pool                 masks @ feat (matching_baseline_utils.py:884-890), accumulated in double
 *   orc_box_nms              torchvision batched_nms, coordinate-trick form
 *                            (call site Sam2MatchingBaseline_noAMG.py:624-629; torchvision pinned 0.19.1 by
 *                            pyproject.toml:57, 0.26.0 in this image)
 *   orc_semantic_ios         compute_semantic_ios (matching_baseline_utils.py:831-867)
 *   orc_rle_*                COCO run-length encoding of the result masks: pycocotools mask_utils.encode
 *                            (call site dataset/coco_ref_dataset.py:601-604; pycocotools pinned 2.0.8 by
 *                            pyproject.toml:35, NOT installed in this image and not under /root/reference, so its
 *                            published maskApi.c algorithm — rleEncode, rleToString, rleFrString — is restated).
 *                            The counts half is pinned by the reference's own mask_to_rle_pytorch
 *                            (sam2/utils/amg.py:111-140, golden vectors tests/golden/rle_*.npz); the string half has
 *                            no reference output to pin against here: PARITY UNPINNED for rleToString (checked
 *                            structurally: alphabet, decode(encode(x)) == x through the restated rleFrString).
 *
 * Parity pinning: tests/test_oracle_golden.py compares every function here with vectors produced by
 * executing the real reference (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared oracle/nttt_oracle.c -o oracle/libnttt_oracle.so -lm
 * (-ffp-contract=off matters: the only fused multiply-adds are the explicit fmaf() calls.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ */
/* low-res threshold + per-mask statistics                                                           */
/* ------------------------------------------------------------------------------------------------ */

/* box layout: x1,y1,x2,y2 with INCLUSIVE max index; empty mask -> 0,0,0,0 (amg.py:339-343). */
static void box_of_mask(const uint8_t* m, int h, int w, int64_t* box) {
  int top = h, bottom = -1, left = w, right = -1;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      if (m[(size_t)y * w + x]) {
        if (y < top) top = y;
        if (y > bottom) bottom = y;
        if (x < left) left = x;
        if (x > right) right = x;
      }
  if (right < left || bottom < top) {
    box[0] = box[1] = box[2] = box[3] = 0;
  } else {
    box[0] = left; box[1] = top; box[2] = right; box[3] = bottom;
  }
}

void orc_threshold_stats(const float* logits, int n, int h, int w, float thr, float off,
                         uint8_t* mask_out, int32_t* area, int64_t* box, int32_t* stab_hi,
                         int32_t* stab_lo) {
  const size_t p = (size_t)h * w;
  const float hi_t = thr + off, lo_t = thr - off;
  for (int i = 0; i < n; ++i) {
    const float* l = logits + (size_t)i * p;
    uint8_t* m = mask_out + (size_t)i * p;
    int32_t a = 0, hi = 0, lo = 0;
    for (size_t j = 0; j < p; ++j) {
      const uint8_t b = l[j] > 0.0f; /* strict, NaN -> 0 */
      m[j] = b;
      a += b;
      hi += l[j] > hi_t;
      lo += l[j] > lo_t;
    }
    area[i] = a;
    stab_hi[i] = hi;
    stab_lo[i] = lo;
    box_of_mask(m, h, w, box + (size_t)i * 4);
  }
}

void orc_mask_boxes(const uint8_t* masks, int n, int h, int w, int64_t* box, int32_t* area) {
  const size_t p = (size_t)h * w;
  for (int i = 0; i < n; ++i) {
    box_of_mask(masks + (size_t)i * p, h, w, box + (size_t)i * 4);
    int32_t a = 0;
    for (size_t j = 0; j < p; ++j) a += masks[(size_t)i * p + j];
    area[i] = a;
  }
}

/* ------------------------------------------------------------------------------------------------ */
/* antialiased bilinear resize (separable, Pillow-style spans)                                       */
/* ------------------------------------------------------------------------------------------------ */

static float aa_scale(int in_size, int out_size) { return (float)in_size / (float)out_size; }

static float aa_support(float scale) { return scale >= 1.0f ? scale : 1.0f; }

/* upper bound on taps per output sample */
int orc_aa_max_taps(int in_size, int out_size) {
  const float support = aa_support(aa_scale(in_size, out_size));
  return (int)ceilf(support) * 2 + 1;
}

static float tri(float x) {
  if (x < 0.0f) x = -x;
  if (x < 1.0f) return 1.0f - x;
  return 0.0f;
}

/* spans and normalised weights of every output coordinate of one axis.
 * xmin[out], xsize[out], wts[out*max_taps] (unused taps zero). */
void orc_aa_weights(int in_size, int out_size, int max_taps, int32_t* xmin, int32_t* xsize, float* wts) {
  const float scale = aa_scale(in_size, out_size);
  const float support = aa_support(scale);
  /* the published kernel divides a double literal by the float scale, then narrows */
  const float invscale = scale >= 1.0f ? (float)(1.0 / (double)scale) : 1.0f;
  for (int i = 0; i < out_size; ++i) {
    const float center = scale * ((float)i + 0.5f);
    int lo = (int)(center - support + 0.5f);
    if (lo < 0) lo = 0;
    int hi = (int)(center + support + 0.5f);
    if (hi > in_size) hi = in_size;
    const int size = hi - lo;
    const float lo_m_center = (float)lo - center;
    float* w = wts + (size_t)i * max_taps;
    float total = 0.0f;
    int j = 0;
    for (; j < size; ++j) {
      w[j] = tri(((float)j + lo_m_center + 0.5f) * invscale);
      total += w[j];
    }
    if (total != 0.0f)
      for (j = 0; j < size; ++j) w[j] /= total;
    for (j = size; j < max_taps; ++j) w[j] = 0.0f;
    xmin[i] = lo;
    xsize[i] = size;
  }
}

static inline float aa_dot(const float* src, size_t stride, const float* w, int size) {
  float acc = src[0] * w[0];
  for (int j = 1; j < size; ++j) acc = fmaf(src[j * stride], w[j], acc);
  return acc;
}

/* src [n, ih, iw] -> dst [n, oh, ow] float; horizontal pass on the contributing rows, then vertical. */
void orc_aa_resize(const float* src, int n, int ih, int iw, int oh, int ow, float* dst) {
  const int tx = orc_aa_max_taps(iw, ow), ty = orc_aa_max_taps(ih, oh);
  int32_t* xmin = malloc(sizeof(int32_t) * ow), * xsz = malloc(sizeof(int32_t) * ow);
  int32_t* ymin = malloc(sizeof(int32_t) * oh), * ysz = malloc(sizeof(int32_t) * oh);
  float* wx = malloc(sizeof(float) * (size_t)ow * tx);
  float* wy = malloc(sizeof(float) * (size_t)oh * ty);
  float* tmp = malloc(sizeof(float) * (size_t)ih * ow);
  orc_aa_weights(iw, ow, tx, xmin, xsz, wx);
  orc_aa_weights(ih, oh, ty, ymin, ysz, wy);
  for (int i = 0; i < n; ++i) {
    const float* s = src + (size_t)i * ih * iw;
    float* d = dst + (size_t)i * oh * ow;
    for (int y = 0; y < ih; ++y)
      for (int x = 0; x < ow; ++x)
        tmp[(size_t)y * ow + x] = aa_dot(s + (size_t)y * iw + xmin[x], 1, wx + (size_t)x * tx, xsz[x]);
    for (int y = 0; y < oh; ++y)
      for (int x = 0; x < ow; ++x)
        d[(size_t)y * ow + x] = aa_dot(tmp + (size_t)ymin[y] * ow + x, ow, wy + (size_t)y * ty, ysz[y]);
  }
  free(xmin); free(xsz); free(ymin); free(ysz); free(wx); free(wy); free(tmp);
}

/* resize + strict > 0 (Sam2MatchingBaseline_noAMG.py:657-663); dst uint8 0/1 */
void orc_aa_resize_threshold(const float* src, int n, int ih, int iw, int oh, int ow, uint8_t* dst) {
  float* buf = malloc(sizeof(float) * (size_t)oh * ow);
  for (int i = 0; i < n; ++i) {
    orc_aa_resize(src + (size_t)i * ih * iw, 1, ih, iw, oh, ow, buf);
    uint8_t* d = dst + (size_t)i * oh * ow;
    for (size_t j = 0; j < (size_t)oh * ow; ++j) d[j] = buf[j] > 0.0f;
  }
  free(buf);
}

/* ------------------------------------------------------------------------------------------------ */
/* mask pooling: pooled[n, c] = sum_p mask[n, p] * feat[c, p]   (feat given channel-major [C, P])     */
/* ------------------------------------------------------------------------------------------------ */
void orc_pool(const uint8_t* masks, const float* feat_cp, int n, int p, int c, float* pooled) {
  for (int i = 0; i < n; ++i) {
    const uint8_t* m = masks + (size_t)i * p;
    for (int ch = 0; ch < c; ++ch) {
      const float* f = feat_cp + (size_t)ch * p;
      double acc = 0.0;
      for (int j = 0; j < p; ++j)
        if (m[j]) acc += (double)f[j];
      pooled[(size_t)i * c + ch] = (float)acc;
    }
  }
}

/* ------------------------------------------------------------------------------------------------ */
/* class-aware box NMS, coordinate-trick form                                                        */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { float s; int32_t i; } ScoreIdx;

static int cmp_desc_stable(const void* a, const void* b) {
  const ScoreIdx* x = a; const ScoreIdx* y = b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i > y->i) - (x->i < y->i);
}

/* boxes [n,4] float XYXY, scores [n], labels [n]; returns count, keep[] = kept indices by score desc */
int orc_box_nms(const float* boxes, const float* scores, const int64_t* labels, int n, float thr,
                int64_t* keep) {
  if (n == 0) return 0;
  float maxc = boxes[0];
  for (int i = 1; i < n * 4; ++i) if (boxes[i] > maxc) maxc = boxes[i];
  float* b = malloc(sizeof(float) * 4 * (size_t)n);
  float* area = malloc(sizeof(float) * (size_t)n);
  for (int i = 0; i < n; ++i) {
    const float off = (float)labels[i] * (maxc + 1.0f);
    for (int k = 0; k < 4; ++k) b[i * 4 + k] = boxes[i * 4 + k] + off;
    area[i] = (b[i * 4 + 2] - b[i * 4 + 0]) * (b[i * 4 + 3] - b[i * 4 + 1]);
  }
  ScoreIdx* order = malloc(sizeof(ScoreIdx) * (size_t)n);
  for (int i = 0; i < n; ++i) { order[i].s = scores[i]; order[i].i = i; }
  qsort(order, n, sizeof(ScoreIdx), cmp_desc_stable);
  uint8_t* dead = calloc(n, 1);
  int cnt = 0;
  for (int a = 0; a < n; ++a) {
    const int i = order[a].i;
    if (dead[i]) continue;
    keep[cnt++] = i;
    for (int c = a + 1; c < n; ++c) {
      const int j = order[c].i;
      if (dead[j]) continue;
      const float xx1 = fmaxf(b[i * 4 + 0], b[j * 4 + 0]), yy1 = fmaxf(b[i * 4 + 1], b[j * 4 + 1]);
      const float xx2 = fminf(b[i * 4 + 2], b[j * 4 + 2]), yy2 = fminf(b[i * 4 + 3], b[j * 4 + 3]);
      const float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
      const float inter = w * h;
      const float ovr = inter / (area[i] + area[j] - inter);
      if (ovr > thr) dead[j] = 1;
    }
  }
  free(b); free(area); free(order); free(dead);
  return cnt;
}

/* ------------------------------------------------------------------------------------------------ */
/* intersection-over-self decay term                                                                 */
/* ------------------------------------------------------------------------------------------------ */

/* masks [k, hw] uint8 0/1, labels [k], obj_sim [k,k] (already clamped >= 0); ios [k]; inter_out [k,k]
 * (optional, may be NULL) receives the integer intersection counts of same-label pairs (0 elsewhere). */
void orc_semantic_ios(const uint8_t* masks, int k, size_t hw, const int64_t* labels, const float* obj_sim,
                      float* ios, int32_t* inter_out) {
  int32_t* area = malloc(sizeof(int32_t) * (size_t)k);
  for (int i = 0; i < k; ++i) {
    int32_t a = 0;
    for (size_t p = 0; p < hw; ++p) a += masks[(size_t)i * hw + p];
    area[i] = a;
  }
  if (inter_out) memset(inter_out, 0, sizeof(int32_t) * (size_t)k * k);
  for (int i = 0; i < k; ++i) {
    /* the diagonal entry (inter zeroed) takes part in the row max: 0*s/area*s, NaN when area == 0 */
    float best = ((0.0f * obj_sim[(size_t)i * k + i]) / (float)area[i]) * obj_sim[(size_t)i * k + i];
    for (int j = 0; j < k; ++j) {
      if (j == i || labels[j] != labels[i]) continue;
      int32_t inter = 0;
      const uint8_t* a = masks + (size_t)i * hw; const uint8_t* b = masks + (size_t)j * hw;
      for (size_t p = 0; p < hw; ++p) inter += a[p] & b[p];
      if (inter_out) inter_out[(size_t)i * k + j] = inter;
      const float s = obj_sim[(size_t)i * k + j];
      const float v = (((float)inter * s) / (float)area[i]) * s;
      /* torch.max propagates NaN */
      if (isnan(v) || isnan(best)) best = NAN; else if (v > best) best = v;
    }
    ios[i] = best;
  }
  free(area);
}


/* ---------------------------------------------------------------------------------------------------
 * COCO RLE (pycocotools maskApi.c, restated).  mask is ROW-major [h, w] u8 as the reference holds it; the
 * encoder walks it in column-major order (np.asfortranarray at coco_ref_dataset.py:602).
 * --------------------------------------------------------------------------------------------------- */
/* rleEncode: alternating run lengths starting with zeros; returns the number of counts (may exceed cap: then only
 * the first cap counts are stored) */
long orc_rle_counts(const uint8_t* mask, int h, int w, uint32_t* cnts, long cap) {
  long m = 0;
  uint32_t run = 0;
  uint8_t prev = 0;
  for (int x = 0; x < w; ++x)
    for (int y = 0; y < h; ++y) {
      const uint8_t v = mask[(size_t)y * w + x] != 0;
      if (v != prev) {
        if (m < cap) cnts[m] = run;
        ++m;
        run = 0;
        prev = v;
      }
      ++run;
    }
  if (m < cap) cnts[m] = run;
  return m + 1;
}

/* rleToString: x = cnts[i] - (i > 2 ? cnts[i-2] : 0), 5 bits per char, low group first, continuation bit 0x20,
 * sign carried by bit 0x10 of the last group, chars 48..111; returns the string length (no terminator written) */
long orc_rle_to_string(const uint32_t* cnts, long m, char* s) {
  long p = 0;
  for (long i = 0; i < m; ++i) {
    long long x = (long long)cnts[i];
    if (i > 2) x -= (long long)cnts[i - 2];
    int more = 1;
    while (more) {
      char c = (char)(x & 0x1f);
      x >>= 5;
      more = (c & 0x10) ? x != -1 : x != 0;
      if (more) c |= 0x20;
      c += 48;
      s[p++] = c;
    }
  }
  return p;
}

/* rleFrString: the inverse; returns the number of counts */
long orc_rle_from_string(const char* s, long len, uint32_t* cnts, long cap) {
  long m = 0, p = 0;
  while (p < len) {
    long long x = 0;
    int k = 0, more = 1;
    while (more) {
      const char c = s[p] - 48;
      x |= (long long)(c & 0x1f) << (5 * k);
      more = c & 0x20;
      ++p;
      ++k;
      if (!more && (c & 0x10)) x |= -1LL << (5 * k);
    }
    if (m > 2) x += (long long)cnts[m - 2];
    if (m < cap) cnts[m] = (uint32_t)x;
    ++m;
  }
  return m;
}
