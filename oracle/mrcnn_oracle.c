/*
 * mrcnn_oracle.c -- CPU restatement of the ROI-stage hot path of miguelalejo/maskrcnn_tf2.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (maskrcnn_tf2_b200/, include/) may
 * include, link, import or call anything in this directory.  Allowed users: tests/,
 * __graft_entry__.smoke(), and bench.py's cpu_baseline / --impl reference legs, and there only as
 * the checker or the timed CPU baseline.
 *
 * PARITY: the reference ships no golden vectors, known-answer tests or fixtures for this path (its
 * tests/ only pin library versions), and its arithmetic lives in the un-vendored PyPI dependency
 * tensorflow==2.2.0/2.3.4/2.4.3/2.5.1 (requirements/requirements_tf2.*.txt:2), not installable
 * here.  This file restates (a) the reference's own Python control flow (src/layers/mrcnn_layers.py,
 * src/common/utils.py -- cited per function as L:/U:) and (b) the published algorithms of the TF
 * CPU kernels it calls (TopKV2, NonMaxSuppressionV3, CropAndResize, CropAndResizeGradImage).
 *   (a) is PINNED: tests/golden/make_reference_layers_golden.py executes the reference's own
 *       ProposalLayer / PyramidROIAlign / DetectionLayer / DetectionTargetLayer code, unmodified, on
 *       a numpy stand-in for tf.*, and this file reproduces the committed outputs bit for bit
 *       (tests/test_reference_layers.py); so are the numpy twins (anchors, build_rpn_targets, ...:
 *       tests/test_reference_pins.py, tests/test_rpn_targets.py).
 *   (b) stays PARITY UNPINNED by TensorFlow itself: pinned instead by hand-derived known answers,
 *       property tests, a second independent numpy restatement (the stand-in above) and independent
 *       implementations available offline (torchvision.ops.nms, torch.topk/sort, grid_sample +
 *       autograd) -- tests/test_oracle_*.py.
 *
 * Numerics: every float operation below is IEEE binary32 in the written order (compile with
 * -ffp-contract=off, no -ffast-math).  exp/log are NOT libm: TF uses Eigen's Cephes-derived packet
 * pexp/plog whose exact polynomial differs between the pinned TF versions, so this file fixes one
 * Cephes-style sequence built from fmaf (orc_expf/orc_logf, <= 2 ulp from the true value).  The
 * CUDA kernels implement the same sequence, which makes decoded boxes -- and therefore NMS keep
 * indices of the fused layers -- bit-identical between oracle and GPU.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* deterministic transcendental functions                                                      */
/* ------------------------------------------------------------------------------------------ */

static inline float orc_pow2i(int e) { /* 2^e for e in [-126,127] */
    union { uint32_t u; float f; } v;
    v.u = (uint32_t)(e + 127) << 23;
    return v.f;
}

ORC_API float orc_expf(float x) {
    if (x != x) return x;
    if (x > 88.72283935546875f) return INFINITY;
    if (x < -103.972076416015625f) return 0.0f;
    float m = floorf(fmaf(x, 1.44269504088896341f, 0.5f));
    float r = fmaf(m, -0.693359375f, x);
    r = fmaf(m, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float y = fmaf(p, r2, r);
    y = y + 1.0f;
    int mi = (int)m;
    int m1 = mi >> 1; /* arithmetic shift: floor(mi/2) */
    int m2 = mi - m1;
    return (y * orc_pow2i(m1)) * orc_pow2i(m2);
}

ORC_API float orc_logf(float x) {
    if (x != x) return x;
    if (x < 0.0f) return NAN;
    if (x == 0.0f) return -INFINITY;
    if (x == INFINITY) return x;
    int e = 0;
    if (x < 1.17549435e-38f) { x = x * 8388608.0f; e = -23; }
    union { uint32_t u; float f; } v;
    v.f = x;
    e += (int)((v.u >> 23) & 0xff) - 126;
    v.u = (v.u & 0x007fffffu) | 0x3f000000u; /* mantissa in [0.5,1) */
    float m = v.f;
    if (m < 0.707106781186547524f) { e -= 1; m = m + m - 1.0f; } else { m = m - 1.0f; }
    float z = m * m;
    float p = 7.0376836292e-2f;
    p = fmaf(p, m, -1.1514610310e-1f);
    p = fmaf(p, m, 1.1676998740e-1f);
    p = fmaf(p, m, -1.2420140846e-1f);
    p = fmaf(p, m, 1.4249322787e-1f);
    p = fmaf(p, m, -1.6668057665e-1f);
    p = fmaf(p, m, 2.0000714765e-1f);
    p = fmaf(p, m, -2.4999993993e-1f);
    p = fmaf(p, m, 3.3333331174e-1f);
    float y = (m * z) * p;
    float fe = (float)e;
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    r = fmaf(fe, 0.693359375f, r);
    return r;
}

/* x86 cvttss2si semantics of Eigen's float->int32 static_cast (TF CPU tf.cast): out-of-range
 * and NaN give INT32_MIN. */
static inline int32_t orc_cast_i32(float v) {
    if (!(v >= -2147483648.0f && v < 2147483648.0f)) return INT32_MIN;
    return (int32_t)v;
}

static inline float orc_minf(float a, float b) { return (b < a) ? b : a; }
static inline float orc_maxf(float a, float b) { return (a < b) ? b : a; }

/* ------------------------------------------------------------------------------------------ */
/* TF TopKV2 (tensorflow/core/kernels/topk_op.cc): k largest, descending, ties -> lower index.  */
/* reference call sites L:246, L:489, L:657                                                    */
/* ------------------------------------------------------------------------------------------ */

#define ORC_BETTER_F(s, a, b) ((s)[a] > (s)[b] || ((s)[a] == (s)[b] && (a) < (b)))

static void sift_down_f(const float* s, int32_t* h, int n, int i) { /* root = worst element */
    for (;;) {
        int l = 2 * i + 1, r = l + 1, w = i;
        if (l < n && ORC_BETTER_F(s, h[w], h[l])) w = l;
        if (r < n && ORC_BETTER_F(s, h[w], h[r])) w = r;
        if (w == i) return;
        int32_t t = h[i]; h[i] = h[w]; h[w] = t;
        i = w;
    }
}

ORC_API void orc_topk(const float* scores, int n, int k, int32_t* idx_out) {
    if (k > n) k = n;
    if (k <= 0) return;
    int32_t* h = (int32_t*)malloc(sizeof(int32_t) * (size_t)k);
    int hs = 0;
    for (int32_t i = 0; i < n; ++i) {
        if (hs < k) {
            h[hs++] = i;
            if (hs == k) for (int j = k / 2 - 1; j >= 0; --j) sift_down_f(scores, h, k, j);
        } else if (ORC_BETTER_F(scores, i, h[0])) {
            h[0] = i;
            sift_down_f(scores, h, k, 0);
        }
    }
    /* heap root is the worst: pop into the tail */
    for (int m = k; m > 0; --m) {
        idx_out[m - 1] = h[0];
        h[0] = h[m - 1];
        sift_down_f(scores, h, m - 1, 0);
    }
    free(h);
}

/* int32 variant used by PyramidROIAlign's sort key (L:656-657) */
#define ORC_BETTER_I(s, a, b) ((s)[a] > (s)[b] || ((s)[a] == (s)[b] && (a) < (b)))
static void sift_down_i(const int32_t* s, int32_t* h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, w = i;
        if (l < n && ORC_BETTER_I(s, h[w], h[l])) w = l;
        if (r < n && ORC_BETTER_I(s, h[w], h[r])) w = r;
        if (w == i) return;
        int32_t t = h[i]; h[i] = h[w]; h[w] = t;
        i = w;
    }
}
ORC_API void orc_topk_i32(const int32_t* keys, int n, int k, int32_t* idx_out) {
    if (k > n) k = n;
    if (k <= 0) return;
    int32_t* h = (int32_t*)malloc(sizeof(int32_t) * (size_t)k);
    int hs = 0;
    for (int32_t i = 0; i < n; ++i) {
        if (hs < k) {
            h[hs++] = i;
            if (hs == k) for (int j = k / 2 - 1; j >= 0; --j) sift_down_i(keys, h, k, j);
        } else if (ORC_BETTER_I(keys, i, h[0])) {
            h[0] = i;
            sift_down_i(keys, h, k, 0);
        }
    }
    for (int m = k; m > 0; --m) {
        idx_out[m - 1] = h[0];
        h[0] = h[m - 1];
        sift_down_i(keys, h, m - 1, 0);
    }
    free(h);
}

/* ------------------------------------------------------------------------------------------ */
/* U:830-851 apply_box_deltas_graph, U:854-869 clip_boxes_graph                                 */
/* ------------------------------------------------------------------------------------------ */

ORC_API void orc_apply_box_deltas(const float* boxes, const float* deltas, int n, float* out) {
    for (int i = 0; i < n; ++i) {
        const float* b = boxes + 4 * (size_t)i;
        const float* d = deltas + 4 * (size_t)i;
        float height = b[2] - b[0];
        float width = b[3] - b[1];
        float center_y = b[0] + 0.5f * height;
        float center_x = b[1] + 0.5f * width;
        center_y = center_y + d[0] * height;
        center_x = center_x + d[1] * width;
        height = height * orc_expf(d[2]);
        width = width * orc_expf(d[3]);
        float y1 = center_y - 0.5f * height;
        float x1 = center_x - 0.5f * width;
        float y2 = y1 + height;
        float x2 = x1 + width;
        float* o = out + 4 * (size_t)i;
        o[0] = y1; o[1] = x1; o[2] = y2; o[3] = x2;
    }
}

ORC_API void orc_clip_boxes(float* boxes, int n, const float* window) {
    const float wy1 = window[0], wx1 = window[1], wy2 = window[2], wx2 = window[3];
    for (int i = 0; i < n; ++i) {
        float* b = boxes + 4 * (size_t)i;
        b[0] = orc_maxf(orc_minf(b[0], wy2), wy1);
        b[1] = orc_maxf(orc_minf(b[1], wx2), wx1);
        b[2] = orc_maxf(orc_minf(b[2], wy2), wy1);
        b[3] = orc_maxf(orc_minf(b[3], wx2), wx1);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* TF NonMaxSuppressionV3 (kernels/image/non_max_suppression_op.cc), as called by               */
/* tf.image.non_max_suppression (score_threshold = -inf).  reference call sites L:225, L:455    */
/* ------------------------------------------------------------------------------------------ */

ORC_API float orc_tf_iou(const float* boxes, int i, int j) {
    const float* a = boxes + 4 * (size_t)i;
    const float* b = boxes + 4 * (size_t)j;
    const float ymin_i = orc_minf(a[0], a[2]), xmin_i = orc_minf(a[1], a[3]);
    const float ymax_i = orc_maxf(a[0], a[2]), xmax_i = orc_maxf(a[1], a[3]);
    const float ymin_j = orc_minf(b[0], b[2]), xmin_j = orc_minf(b[1], b[3]);
    const float ymax_j = orc_maxf(b[0], b[2]), xmax_j = orc_maxf(b[1], b[3]);
    const float area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i);
    const float area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j);
    if (area_i <= 0.0f || area_j <= 0.0f) return 0.0f;
    const float iy0 = orc_maxf(ymin_i, ymin_j), ix0 = orc_maxf(xmin_i, xmin_j);
    const float iy1 = orc_minf(ymax_i, ymax_j), ix1 = orc_minf(xmax_i, xmax_j);
    const float inter = orc_maxf(iy1 - iy0, 0.0f) * orc_maxf(ix1 - ix0, 0.0f);
    return inter / (area_i + area_j - inter);
}

ORC_API int orc_nms(const float* boxes, const float* scores, int m, int max_out, float iou_thr,
                    int32_t* keep) {
    if (m <= 0 || max_out <= 0) return 0;
    /* candidates: score > -inf (NaN never passes) in (score desc, index asc) order */
    float* s = (float*)malloc(sizeof(float) * (size_t)m);
    int32_t* cand = (int32_t*)malloc(sizeof(int32_t) * (size_t)m);
    int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)m);
    int nc = 0;
    for (int i = 0; i < m; ++i)
        if (scores[i] > -INFINITY) { cand[nc] = i; s[nc] = scores[i]; ++nc; }
    orc_topk(s, nc, nc, order); /* positions within cand; ties -> lower position == lower index */
    int nsel = 0;
    for (int c = 0; c < nc && nsel < max_out; ++c) {
        const int idx = cand[order[c]];
        int suppressed = 0;
        for (int j = nsel - 1; j >= 0; --j) {
            if (orc_tf_iou(boxes, idx, keep[j]) > iou_thr) { suppressed = 1; break; }
        }
        if (!suppressed) keep[nsel++] = idx;
    }
    free(s); free(cand); free(order);
    return nsel;
}

/* ------------------------------------------------------------------------------------------ */
/* TF CropAndResize / CropAndResizeGradImage (kernels/image/crop_and_resize_op.cc), bilinear,   */
/* extrapolation_value = 0.  reference call sites L:641, L:948                                  */
/* ------------------------------------------------------------------------------------------ */

typedef struct { int valid; int lo, hi; float lerp; } orc_tap;

static inline orc_tap orc_make_tap(float c1, float c2, int size, int crop, int t, float scale) {
    orc_tap r;
    float in;
    if (crop > 1) in = c1 * (float)(size - 1) + (float)t * scale;
    else in = (float)(0.5 * (double)(c1 + c2) * (double)(size - 1));
    /* TF writes `in < 0 || in > size-1`; the negated form only differs for NaN (TF: undefined indexing) */
    if (!(in >= 0.0f && in <= (float)(size - 1))) { r.valid = 0; r.lo = r.hi = 0; r.lerp = 0.0f; return r; }
    r.valid = 1;
    r.lo = (int)floorf(in);
    r.hi = (int)ceilf(in);
    r.lerp = in - (float)r.lo;
    return r;
}

static void orc_crop_one(const float* img /*[H,W,C] of the selected batch entry*/, int H, int W, int C,
                         const float* box, int ph, int pw, float* out /*[ph,pw,C]*/) {
    const float y1 = box[0], x1 = box[1], y2 = box[2], x2 = box[3];
    const float hs = (ph > 1) ? (y2 - y1) * (float)(H - 1) / (float)(ph - 1) : 0.0f;
    const float ws = (pw > 1) ? (x2 - x1) * (float)(W - 1) / (float)(pw - 1) : 0.0f;
    for (int y = 0; y < ph; ++y) {
        orc_tap ty = orc_make_tap(y1, y2, H, ph, y, hs);
        float* orow = out + (size_t)y * pw * C;
        if (!ty.valid) { memset(orow, 0, sizeof(float) * (size_t)pw * C); continue; }
        for (int x = 0; x < pw; ++x) {
            orc_tap tx = orc_make_tap(x1, x2, W, pw, x, ws);
            float* o = orow + (size_t)x * C;
            if (!tx.valid) { memset(o, 0, sizeof(float) * (size_t)C); continue; }
            const float* tl = img + ((size_t)ty.lo * W + tx.lo) * C;
            const float* tr = img + ((size_t)ty.lo * W + tx.hi) * C;
            const float* bl = img + ((size_t)ty.hi * W + tx.lo) * C;
            const float* br = img + ((size_t)ty.hi * W + tx.hi) * C;
            for (int d = 0; d < C; ++d) {
                const float top = tl[d] + (tr[d] - tl[d]) * tx.lerp;
                const float bot = bl[d] + (br[d] - bl[d]) * tx.lerp;
                o[d] = top + (bot - top) * ty.lerp;
            }
        }
    }
}

ORC_API void orc_crop_and_resize(const float* image, int B, int H, int W, int C, const float* boxes,
                                 const int32_t* box_ind, int nb, int ph, int pw, float* out) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int b = 0; b < nb; ++b) {
        const int bi = box_ind[b];
        if (bi < 0 || bi >= B) continue;
        orc_crop_one(image + (size_t)bi * H * W * C, H, W, C, boxes + 4 * (size_t)b, ph, pw,
                     out + (size_t)b * ph * pw * C);
    }
}

static void orc_crop_grad_one(const float* g /*[ph,pw,C]*/, int H, int W, int C, const float* box, int ph,
                              int pw, float* gimg /*[H,W,C]*/) {
    const float y1 = box[0], x1 = box[1], y2 = box[2], x2 = box[3];
    const float hs = (ph > 1) ? (y2 - y1) * (float)(H - 1) / (float)(ph - 1) : 0.0f;
    const float ws = (pw > 1) ? (x2 - x1) * (float)(W - 1) / (float)(pw - 1) : 0.0f;
    for (int y = 0; y < ph; ++y) {
        orc_tap ty = orc_make_tap(y1, y2, H, ph, y, hs);
        if (!ty.valid) continue;
        for (int x = 0; x < pw; ++x) {
            orc_tap tx = orc_make_tap(x1, x2, W, pw, x, ws);
            if (!tx.valid) continue;
            const float* gv = g + ((size_t)y * pw + x) * C;
            float* tl = gimg + ((size_t)ty.lo * W + tx.lo) * C;
            float* tr = gimg + ((size_t)ty.lo * W + tx.hi) * C;
            float* bl = gimg + ((size_t)ty.hi * W + tx.lo) * C;
            float* br = gimg + ((size_t)ty.hi * W + tx.hi) * C;
            for (int d = 0; d < C; ++d) {
                const float dtop = (1.0f - ty.lerp) * gv[d];
                tl[d] += (1.0f - tx.lerp) * dtop;
                tr[d] += tx.lerp * dtop;
                const float dbot = ty.lerp * gv[d];
                bl[d] += (1.0f - tx.lerp) * dbot;
                br[d] += tx.lerp * dbot;
            }
        }
    }
}

/* grads_image must be zero-initialised by the caller (TF's kernel zero-fills it itself). */
ORC_API void orc_crop_and_resize_grad_image(const float* grads, int B, int H, int W, int C,
                                            const float* boxes, const int32_t* box_ind, int nb, int ph,
                                            int pw, float* grads_image) {
    for (int b = 0; b < nb; ++b) {
        const int bi = box_ind[b];
        if (bi < 0 || bi >= B) continue;
        orc_crop_grad_one(grads + (size_t)b * ph * pw * C, H, W, C, boxes + 4 * (size_t)b, ph, pw,
                          grads_image + (size_t)bi * H * W * C);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* ProposalLayer.call  L:233-269 (+ nms L:224-231)                                              */
/* ------------------------------------------------------------------------------------------ */

ORC_API void orc_proposal_layer(const float* rpn_probs /*[B,A,2]*/, const float* rpn_bbox /*[B,A,4]*/,
                                const float* anchors /*[B,A,4]*/, int B, int A, int pre_nms_limit,
                                int proposal_count, const float* std_dev /*[4]*/, float nms_thr,
                                float* proposals /*[B,P,4]*/, int32_t* topk_idx /*[B,K] or NULL*/,
                                int32_t* keep_idx /*[B,P] -1 padded, or NULL*/,
                                int32_t* keep_count /*[B] or NULL*/, float* pre_nms_boxes /*[B,K,4] or NULL*/) {
    const int K = pre_nms_limit < A ? pre_nms_limit : A; /* L:245 */
    const int P = proposal_count;
    const float window[4] = {0.0f, 0.0f, 1.0f, 1.0f}; /* L:259 */
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        float* fg = (float*)malloc(sizeof(float) * (size_t)A);
        int32_t* ix = (int32_t*)malloc(sizeof(int32_t) * (size_t)K);
        float* sc = (float*)malloc(sizeof(float) * (size_t)K);
        float* dl = (float*)malloc(sizeof(float) * 4 * (size_t)K);
        float* an = (float*)malloc(sizeof(float) * 4 * (size_t)K);
        float* bx = (float*)malloc(sizeof(float) * 4 * (size_t)K);
        int32_t* keep = (int32_t*)malloc(sizeof(int32_t) * (size_t)(P > 0 ? P : 1));
        for (int a = 0; a < A; ++a) fg[a] = rpn_probs[((size_t)b * A + a) * 2 + 1]; /* L:235 */
        orc_topk(fg, A, K, ix);                                                       /* L:246 */
        for (int k = 0; k < K; ++k) {                                                 /* L:247-250 */
            const size_t src = (size_t)b * A + ix[k];
            sc[k] = fg[ix[k]];
            for (int c = 0; c < 4; ++c) {
                dl[4 * k + c] = rpn_bbox[src * 4 + c] * std_dev[c]; /* L:238 (scale first) */
                an[4 * k + c] = anchors[src * 4 + c];
            }
        }
        orc_apply_box_deltas(an, dl, K, bx); /* L:254 */
        orc_clip_boxes(bx, K, window);       /* L:260 */
        const int n = orc_nms(bx, sc, K, P, nms_thr, keep); /* L:225 */
        float* out = proposals + (size_t)b * P * 4;
        for (int p = 0; p < P; ++p) {                        /* L:227-230 gather + zero pad */
            for (int c = 0; c < 4; ++c) out[4 * p + c] = (p < n) ? bx[4 * keep[p] + c] : 0.0f;
            if (keep_idx) keep_idx[(size_t)b * P + p] = (p < n) ? keep[p] : -1;
        }
        if (keep_count) keep_count[b] = n;
        if (topk_idx) memcpy(topk_idx + (size_t)b * K, ix, sizeof(int32_t) * (size_t)K);
        if (pre_nms_boxes) memcpy(pre_nms_boxes + (size_t)b * K * 4, bx, sizeof(float) * 4 * (size_t)K);
        free(fg); free(ix); free(sc); free(dl); free(an); free(bx); free(keep);
    }
}

/* Gradient of ProposalLayer.call w.r.t. rpn_bbox (TF autodiff; the reference has no stop_gradient on the
 * proposals, SURVEY Q7: M:155-157,168 -> L:227 gather -> U:854-869 clip -> U:830-851 decode -> L:238 scale).
 * tf.minimum passes the gradient to x where x <= y, tf.maximum where x >= y; padded rows (L:229-230) and the
 * anchors receive none.  grad_rpn_bbox [B,A,4] is zero-filled here. */
ORC_API void orc_proposal_layer_grad(const float* grad_proposals /*[B,P,4]*/, const float* rpn_bbox /*[B,A,4]*/,
                                     const float* anchors /*[B,A,4]*/, const int32_t* topk_idx /*[B,K]*/,
                                     const int32_t* keep_idx /*[B,P]*/, int B, int A, int K, int P,
                                     const float* std_dev, float* grad_rpn_bbox /*[B,A,4]*/) {
    memset(grad_rpn_bbox, 0, sizeof(float) * 4 * (size_t)B * A);
    for (int b = 0; b < B; ++b) {
        for (int p = 0; p < P; ++p) {
            const int k = keep_idx[(size_t)b * P + p];
            if (k < 0) continue;
            const int a = topk_idx[(size_t)b * K + k];
            const float* an = anchors + ((size_t)b * A + a) * 4;
            const float* raw = rpn_bbox + ((size_t)b * A + a) * 4;
            const float* g = grad_proposals + ((size_t)b * P + p) * 4;
            float d[4];
            for (int c = 0; c < 4; ++c) d[c] = raw[c] * std_dev[c];
            /* forward (U:836-850) */
            const float height = an[2] - an[0], width = an[3] - an[1];
            float cy = an[0] + 0.5f * height, cx = an[1] + 0.5f * width;
            cy = cy + d[0] * height;
            cx = cx + d[1] * width;
            const float eh = orc_expf(d[2]), ew = orc_expf(d[3]);
            const float h2 = height * eh, w2 = width * ew;
            const float y1 = cy - 0.5f * h2, x1 = cx - 0.5f * w2;
            const float v[4] = {y1, x1, y1 + h2, x1 + w2};
            /* clip backward, window [0,0,1,1]: min(v,1) passes where v <= 1, max(.,0) where min(v,1) >= 0 */
            float gv[4];
            for (int c = 0; c < 4; ++c) {
                const float m = orc_minf(v[c], 1.0f);
                gv[c] = (v[c] <= 1.0f && m >= 0.0f) ? g[c] : 0.0f;
            }
            /* decode backward: y2 = y1 + h2, y1 = cy - 0.5 h2 */
            const float gy1 = gv[0] + gv[2], gx1 = gv[1] + gv[3];
            const float gh2 = gv[2] + (-0.5f) * gy1, gw2 = gv[3] + (-0.5f) * gx1;
            const float gd0 = gy1 * height, gd1 = gx1 * width;
            const float gd2 = (gh2 * height) * eh, gd3 = (gw2 * width) * ew;
            float* o = grad_rpn_bbox + ((size_t)b * A + a) * 4;
            o[0] = gd0 * std_dev[0];
            o[1] = gd1 * std_dev[1];
            o[2] = gd2 * std_dev[2];
            o[3] = gd3 * std_dev[3];
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* PyramidROIAlign.call  L:583-664 -- literal control flow, including the first-appearance map   */
/* table (L:613-619,641), concat/truncate (L:644,648) and the batch*100000+box re-sort           */
/* (L:656-659)                                                                                  */
/* ------------------------------------------------------------------------------------------ */

ORC_API int32_t orc_roi_level(const float* box, float img_h, float img_w, float denominator) {
    const float h = box[2] - box[0];
    const float w = box[3] - box[1];
    const float image_area = img_h * img_w;                                  /* L:604 */
    const float x = sqrtf(h * w) / (denominator / sqrtf(image_area));        /* L:605 */
    const float lv = orc_logf(x) / 0.693147182464599609375f;                 /* U:825-827, fp32 log(2.0) */
    const int32_t r = orc_cast_i32(rintf(lv));                               /* tf.round = half-to-even */
    int32_t level = 4 + r;                                                   /* r >= INT32_MIN: no overflow */
    if (level < 2) level = 2;
    if (level > 5) level = 5;
    return level;
}

/* map_mode 0 = reference (first-appearance rank, quirk Q2); 1 = canonical level-2 */
ORC_API void orc_pyramid_roi_align(const float* boxes /*[B,N,4]*/, int B, int N, float img_h, float img_w,
                                   const float* const* fmaps /*4 x [B,H,W,C]*/, const int* Hs, const int* Ws,
                                   int C, int ph, int pw, float denominator, int map_mode,
                                   float* out /*[B,N,ph,pw,C]*/, int32_t* roi_level_out /*[B,N] or NULL*/,
                                   int32_t* roi_map_out /*[B,N] or NULL*/) {
    const int BN = B * N;
    const size_t row = (size_t)ph * pw * C;
    int32_t* level = (int32_t*)malloc(sizeof(int32_t) * (size_t)BN);
    for (int i = 0; i < BN; ++i) level[i] = orc_roi_level(boxes + 4 * (size_t)i, img_h, img_w, denominator);
    if (roi_level_out) memcpy(roi_level_out, level, sizeof(int32_t) * (size_t)BN);

    /* tf.unique: first-occurrence order; pad with 2 and keep 4 (L:613-615) */
    int32_t padded[8];
    int nu = 0;
    if (map_mode == 0) {
        for (int i = 0; i < BN; ++i) {
            int seen = 0;
            for (int u = 0; u < nu; ++u) if (padded[u] == level[i]) { seen = 1; break; }
            if (!seen) padded[nu++] = level[i];
        }
        for (int u = nu; u < nu + 4; ++u) padded[u] = 2;
    } else {
        padded[0] = 2; padded[1] = 3; padded[2] = 4; padded[3] = 5;
    }

    /* per-level where/gather/crop, concatenated, truncated to B*N rows (L:617-648) */
    float* pooled = (float*)malloc(sizeof(float) * row * (size_t)BN);
    int32_t* b2l_batch = (int32_t*)malloc(sizeof(int32_t) * (size_t)BN);
    int32_t* b2l_box = (int32_t*)malloc(sizeof(int32_t) * (size_t)BN);
    int32_t* b2l_map = (int32_t*)malloc(sizeof(int32_t) * (size_t)BN);
    int rows = 0;
    for (int i = 0; i < 4 && rows < BN; ++i) {
        for (int f = 0; f < BN && rows < BN; ++f) { /* tf.where: row-major (batch, box) order */
            if (level[f] != padded[i]) continue;
            b2l_batch[rows] = f / N;
            b2l_box[rows] = f % N;
            b2l_map[rows] = i;
            ++rows;
        }
    }
    /* rows == BN always: every ROI's level is in the padded list before the duplicates start */
#pragma omp parallel for schedule(dynamic, 8)
    for (int r = 0; r < rows; ++r) {
        const int i = b2l_map[r];
        const int bi = b2l_batch[r];
        orc_crop_one(fmaps[i] + (size_t)bi * Hs[i] * Ws[i] * C, Hs[i], Ws[i], C,
                     boxes + 4 * ((size_t)bi * N + b2l_box[r]), ph, pw, pooled + row * (size_t)r);
    }
    /* sorting_tensor = batch*100000 + box; top_k(all) reversed; gather (L:656-659) */
    int32_t* key = (int32_t*)malloc(sizeof(int32_t) * (size_t)BN);
    int32_t* ord = (int32_t*)malloc(sizeof(int32_t) * (size_t)BN);
    for (int r = 0; r < rows; ++r) key[r] = b2l_batch[r] * 100000 + b2l_box[r];
    orc_topk_i32(key, rows, rows, ord);
#pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; ++r) {
        const int src = ord[rows - 1 - r]; /* [::-1] */
        memcpy(out + row * (size_t)r, pooled + row * (size_t)src, sizeof(float) * row);
        if (roi_map_out) roi_map_out[r] = b2l_map[src];
    }
    free(level); free(pooled); free(b2l_batch); free(b2l_box); free(b2l_map); free(key); free(ord);
}

/* Gradient of the flow above w.r.t. the four feature maps (TF autodiff: gather -> slice -> concat ->
 * CropAndResizeGradImage per map; boxes are stop_gradient'ed L:628-629).  grad_fmaps are zero-filled
 * here.  Accumulation order: per map, ROIs in (batch, box) order, as the per-level tf.where lists them. */
ORC_API void orc_pyramid_roi_align_grad(const float* grad_out /*[B,N,ph,pw,C]*/, const float* boxes, int B, int N,
                                        float img_h, float img_w, const int* Hs, const int* Ws, int C, int ph,
                                        int pw, float denominator, int map_mode, float* const* grad_fmaps) {
    const int BN = B * N;
    const size_t row = (size_t)ph * pw * C;
    int32_t* level = (int32_t*)malloc(sizeof(int32_t) * (size_t)BN);
    for (int i = 0; i < BN; ++i) level[i] = orc_roi_level(boxes + 4 * (size_t)i, img_h, img_w, denominator);
    int32_t padded[8];
    int nu = 0;
    if (map_mode == 0) {
        for (int i = 0; i < BN; ++i) {
            int seen = 0;
            for (int u = 0; u < nu; ++u) if (padded[u] == level[i]) { seen = 1; break; }
            if (!seen) padded[nu++] = level[i];
        }
        for (int u = nu; u < nu + 4; ++u) padded[u] = 2;
    } else {
        padded[0] = 2; padded[1] = 3; padded[2] = 4; padded[3] = 5;
    }
    for (int i = 0; i < 4; ++i) memset(grad_fmaps[i], 0, sizeof(float) * (size_t)B * Hs[i] * Ws[i] * C);
    int rows = 0;
    for (int i = 0; i < 4 && rows < BN; ++i) {
        for (int f = 0; f < BN && rows < BN; ++f) {
            if (level[f] != padded[i]) continue;
            ++rows;
            const int bi = f / N;
            orc_crop_grad_one(grad_out + row * (size_t)f, Hs[i], Ws[i], C, boxes + 4 * (size_t)f, ph, pw,
                              grad_fmaps[i] + (size_t)bi * Hs[i] * Ws[i] * C);
        }
    }
    free(level);
}

/* ------------------------------------------------------------------------------------------ */
/* DetectionLayer.call L:503-524 + refine_detections L:369-501 -- literal, incl. the            */
/* class-agnostic NMS (Q3) and the O(n^2) broadcast intersections (Q4)                          */
/* ------------------------------------------------------------------------------------------ */

static int orc_refine_detections(const float* rois /*[N,4]*/, const float* probs /*[N,NC]*/,
                                 const float* deltas /*[N,NC,4]*/, const float* window /*[4]*/, int N, int NC,
                                 const float* std_dev, float min_conf, int use_min_conf, int max_inst,
                                 float nms_thr, float* det /*[max_inst,6]*/) {
    int32_t* class_ids = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
    float* class_scores = (float*)malloc(sizeof(float) * (size_t)N);
    float* dspec = (float*)malloc(sizeof(float) * 4 * (size_t)N);
    float* refined = (float*)malloc(sizeof(float) * 4 * (size_t)N);
    for (int i = 0; i < N; ++i) { /* L:385-393: argmax = first maximum */
        const float* p = probs + (size_t)i * NC;
        int best = 0;
        for (int c = 1; c < NC; ++c) if (p[c] > p[best]) best = c;
        class_ids[i] = best;
        class_scores[i] = p[best];
        for (int c = 0; c < 4; ++c) dspec[4 * i + c] = deltas[((size_t)i * NC + best) * 4 + c] * std_dev[c]; /* L:396 */
    }
    orc_apply_box_deltas(rois, dspec, N, refined); /* L:396 */
    orc_clip_boxes(refined, N, window);            /* L:398 */

    int32_t* keep = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
    int nkeep = 0;
    for (int i = 0; i < N; ++i) if (class_ids[i] > 0) keep[nkeep++] = i; /* L:402 */
    if (use_min_conf) {                                                   /* L:404-414 */
        int32_t* conf_keep = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
        int nconf = 0;
        for (int i = 0; i < N; ++i) if (class_scores[i] >= min_conf) conf_keep[nconf++] = i;
        int32_t* nk = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
        int nn = 0;
        for (int c = 0; c < nconf; ++c) { /* boolean_mask(conf_keep, column sums): order of conf_keep */
            int sum = 0;
            for (int k = 0; k < nkeep; ++k) sum += (conf_keep[c] == keep[k]);
            if (sum) nk[nn++] = conf_keep[c];
        }
        memcpy(keep, nk, sizeof(int32_t) * (size_t)nn);
        nkeep = nn;
        free(conf_keep); free(nk);
    }
    /* L:418-420 */
    float* pre_scores = (float*)malloc(sizeof(float) * (size_t)(nkeep + 1));
    float* pre_rois = (float*)malloc(sizeof(float) * 4 * (size_t)(nkeep + 1));
    for (int k = 0; k < nkeep; ++k) {
        pre_scores[k] = class_scores[keep[k]];
        memcpy(pre_rois + 4 * k, refined + 4 * keep[k], sizeof(float) * 4);
    }
    /* L:440-468: _nms_keep_func's bool_mask is all-ones -> ONE class-agnostic NMS over every kept ROI */
    int32_t* class_keep = (int32_t*)malloc(sizeof(int32_t) * (size_t)(max_inst > 0 ? max_inst : 1));
    const int nsel = orc_nms(pre_rois, pre_scores, nkeep, max_inst, nms_thr, class_keep);
    int32_t* nms_keep = (int32_t*)malloc(sizeof(int32_t) * (size_t)(max_inst > 0 ? max_inst : 1));
    for (int s = 0; s < nsel; ++s) nms_keep[s] = keep[class_keep[s]]; /* L:460; -1 pad added L:462-463, removed L:472 */
    /* L:475-478: intersection, order of nms_keep */
    int32_t* fin = (int32_t*)malloc(sizeof(int32_t) * (size_t)(max_inst > 0 ? max_inst : 1));
    int nfin = 0;
    for (int s = 0; s < nsel; ++s) {
        int sum = 0;
        for (int k = 0; k < nkeep; ++k) sum += (nms_keep[s] == keep[k]);
        if (sum) fin[nfin++] = nms_keep[s];
    }
    /* L:486-490: top_k by score over the survivors */
    float* fs = (float*)malloc(sizeof(float) * (size_t)(nfin + 1));
    int32_t* top = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nfin + 1));
    for (int s = 0; s < nfin; ++s) fs[s] = class_scores[fin[s]];
    const int num_keep = nfin < max_inst ? nfin : max_inst;
    orc_topk(fs, nfin, num_keep, top);
    memset(det, 0, sizeof(float) * 6 * (size_t)max_inst); /* L:498-500 */
    for (int s = 0; s < num_keep; ++s) {                   /* L:494-496 */
        const int i = fin[top[s]];
        memcpy(det + 6 * s, refined + 4 * i, sizeof(float) * 4);
        det[6 * s + 4] = (float)class_ids[i];
        det[6 * s + 5] = class_scores[i];
    }
    free(class_ids); free(class_scores); free(dspec); free(refined); free(keep); free(pre_scores);
    free(pre_rois); free(class_keep); free(nms_keep); free(fin); free(fs); free(top);
    return num_keep;
}

ORC_API void orc_detection_layer(const float* rois /*[B,N,4]*/, const float* probs /*[B,N,NC]*/,
                                 const float* deltas /*[B,N,NC,4]*/, const float* image_meta /*[B,meta_len]*/,
                                 int B, int N, int NC, int meta_len, const float* std_dev, float min_conf,
                                 int use_min_conf, int max_inst, float nms_thr, float* detections /*[B,max_inst,6]*/,
                                 int32_t* det_count /*[B] or NULL*/) {
    /* L:513-515: window = (meta.window - [0,0,1,1]) / ([h,w,h,w] - 1), h,w from image 0 (Q6) */
    const float ih = image_meta[4], iw = image_meta[5];
    const float sh = ih - 1.0f, sw = iw - 1.0f;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const float* wm = image_meta + (size_t)b * meta_len + 7;
        float window[4];
        window[0] = (wm[0] - 0.0f) / sh;
        window[1] = (wm[1] - 0.0f) / sw;
        window[2] = (wm[2] - 1.0f) / sh;
        window[3] = (wm[3] - 1.0f) / sw;
        const int n = orc_refine_detections(rois + (size_t)b * N * 4, probs + (size_t)b * N * NC,
                                            deltas + (size_t)b * N * NC * 4, window, N, NC, std_dev, min_conf,
                                            use_min_conf, max_inst, nms_thr, detections + (size_t)b * max_inst * 6);
        if (det_count) det_count[b] = n;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* DetectionTargetLayer.call L:313-325 + detection_targets_graph L:844-967                      */
/* tf.random.shuffle (L:905,910) is unseeded in the reference (Q8); here the permutation is     */
/* injected: candidates are ordered by (rand_key asc, index asc) with one uint32 key per         */
/* (untrimmed) proposal row.  Inputs must be NaN-free with positive-area GT boxes.               */
/* ------------------------------------------------------------------------------------------ */

static float orc_plain_iou(const float* a, const float* b) { /* overlaps_graph L:982-1007 */
    const float y1 = orc_maxf(a[0], b[0]);
    const float x1 = orc_maxf(a[1], b[1]);
    const float y2 = orc_minf(a[2], b[2]);
    const float x2 = orc_minf(a[3], b[3]);
    const float inter = orc_maxf(x2 - x1, 0.0f) * orc_maxf(y2 - y1, 0.0f);
    const float a1 = (a[2] - a[0]) * (a[3] - a[1]);
    const float a2 = (b[2] - b[0]) * (b[3] - b[1]);
    const float uni = a1 + a2 - inter;
    return inter / uni;
}

static void orc_sort_by_key(int32_t* idx, int n, const uint32_t* keys /*indexed by orig row*/,
                            const int32_t* orig_row) {
    /* insertion sort on (key, idx): n <= a few thousand in tests; stable and obviously correct */
    for (int i = 1; i < n; ++i) {
        int32_t v = idx[i];
        uint32_t kv = keys[orig_row[v]];
        int j = i - 1;
        while (j >= 0 && (keys[orig_row[idx[j]]] > kv || (keys[orig_row[idx[j]]] == kv && idx[j] > v))) {
            idx[j + 1] = idx[j];
            --j;
        }
        idx[j + 1] = v;
    }
}

static void orc_detection_targets_one(const float* proposals_in /*[P,4]*/, const int32_t* gt_class_ids_in /*[G]*/,
                                      const float* gt_boxes_in /*[G,4]*/, const uint8_t* gt_masks /*[MH,MW,G]*/,
                                      const uint32_t* rand_keys /*[P]*/, int P, int G, int MH, int MW, int T,
                                      double roi_positive_ratio, const float* bbox_std_dev, int mask_h, int mask_w,
                                      int use_mini_masks, float* rois /*[T,4]*/, int32_t* class_ids /*[T]*/,
                                      float* deltas /*[T,4]*/, float* masks /*[T,mask_h,mask_w]*/,
                                      int32_t* counts /*[2] pos,neg or NULL*/) {
    /* L:871-874 trim zero padding */
    float* prop = (float*)malloc(sizeof(float) * 4 * (size_t)(P + 1));
    int32_t* prop_row = (int32_t*)malloc(sizeof(int32_t) * (size_t)(P + 1));
    int np = 0;
    for (int i = 0; i < P; ++i) {
        const float* p = proposals_in + 4 * (size_t)i;
        if ((fabsf(p[0]) + fabsf(p[1]) + fabsf(p[2]) + fabsf(p[3])) != 0.0f) {
            memcpy(prop + 4 * np, p, sizeof(float) * 4);
            prop_row[np++] = i;
        }
    }
    float* gtb = (float*)malloc(sizeof(float) * 4 * (size_t)(G + 1));
    int32_t* gtc = (int32_t*)malloc(sizeof(int32_t) * (size_t)(G + 1));
    int32_t* gtm = (int32_t*)malloc(sizeof(int32_t) * (size_t)(G + 1)); /* channel of gt_masks */
    float* crowd = (float*)malloc(sizeof(float) * 4 * (size_t)(G + 1));
    int ng = 0, ncrowd = 0;
    for (int g = 0; g < G; ++g) {
        const float* q = gt_boxes_in + 4 * (size_t)g;
        if ((fabsf(q[0]) + fabsf(q[1]) + fabsf(q[2]) + fabsf(q[3])) == 0.0f) continue;
        /* L:879-884 crowds have negative class ids */
        if (gt_class_ids_in[g] < 0) { memcpy(crowd + 4 * ncrowd, q, sizeof(float) * 4); ++ncrowd; }
        else if (gt_class_ids_in[g] > 0) {
            memcpy(gtb + 4 * ng, q, sizeof(float) * 4);
            gtc[ng] = gt_class_ids_in[g];
            gtm[ng] = g;
            ++ng;
        }
    }
    /* L:887-900 */
    float* iou_max = (float*)malloc(sizeof(float) * (size_t)(np + 1));
    int32_t* iou_arg = (int32_t*)malloc(sizeof(int32_t) * (size_t)(np + 1));
    int32_t* pos = (int32_t*)malloc(sizeof(int32_t) * (size_t)(np + 1));
    int32_t* neg = (int32_t*)malloc(sizeof(int32_t) * (size_t)(np + 1));
    int npos = 0, nneg = 0;
    for (int i = 0; i < np; ++i) {
        float best = -FLT_MAX; /* Eigen MaxReducer init = lowest() */
        int arg = 0;
        for (int g = 0; g < ng; ++g) {
            const float v = orc_plain_iou(prop + 4 * i, gtb + 4 * g);
            if (v > best) { best = v; arg = g; } /* argmax L:918 = first maximum */
        }
        float cmax = -FLT_MAX;
        for (int c = 0; c < ncrowd; ++c) {
            const float v = orc_plain_iou(prop + 4 * i, crowd + 4 * c);
            if (v > cmax) cmax = v;
        }
        iou_max[i] = best;
        iou_arg[i] = arg;
        const int no_crowd = (cmax < 0.001f);
        if (best >= 0.5f) pos[npos++] = i;
        if (best < 0.5f && no_crowd) neg[nneg++] = i;
    }
    /* L:904-910 subsample */
    int positive_count = (int)((double)T * roi_positive_ratio);
    orc_sort_by_key(pos, npos, rand_keys, prop_row);
    if (npos > positive_count) npos = positive_count;
    positive_count = npos;
    const float r = (float)(1.0 / roi_positive_ratio);
    int negative_count = orc_cast_i32(r * (float)positive_count) - positive_count;
    orc_sort_by_key(neg, nneg, rand_keys, prop_row);
    if (negative_count < 0) negative_count = 0; /* [:negative] would drop from the end; never negative for ratio<=1 */
    if (nneg > negative_count) nneg = negative_count;

    memset(rois, 0, sizeof(float) * 4 * (size_t)T);
    memset(class_ids, 0, sizeof(int32_t) * (size_t)T);
    memset(deltas, 0, sizeof(float) * 4 * (size_t)T);
    memset(masks, 0, sizeof(float) * (size_t)T * mask_h * mask_w);
    /* TF would fail on shape mismatch if pos+neg > T; the reference's ratios keep pos+neg <= T */
    for (int s = 0; s < npos && s < T; ++s) {
        const float* box = prop + 4 * pos[s];
        const int g = iou_arg[pos[s]];
        const float* gt = gtb + 4 * g;
        memcpy(rois + 4 * s, box, sizeof(float) * 4);
        class_ids[s] = gtc[g];
        /* U:775-798 box_refinement_graph, then /= bbox_std_dev (L:925) */
        const float height = box[2] - box[0];
        const float width = box[3] - box[1];
        const float center_y = box[0] + 0.5f * height;
        const float center_x = box[1] + 0.5f * width;
        const float gt_height = gt[2] - gt[0];
        const float gt_width = gt[3] - gt[1];
        const float gt_center_y = gt[0] + 0.5f * gt_height;
        const float gt_center_x = gt[1] + 0.5f * gt_width;
        deltas[4 * s + 0] = ((gt_center_y - center_y) / height) / bbox_std_dev[0];
        deltas[4 * s + 1] = ((gt_center_x - center_x) / width) / bbox_std_dev[1];
        deltas[4 * s + 2] = orc_logf(gt_height / (height + 1e-3f)) / bbox_std_dev[2];
        deltas[4 * s + 3] = orc_logf(gt_width / (width + 1e-3f)) / bbox_std_dev[3];
        /* L:929-954 mask target */
        float mb[4] = {box[0], box[1], box[2], box[3]};
        if (use_mini_masks) { /* L:935-946 */
            const float gh = gt[2] - gt[0];
            const float gw = gt[3] - gt[1];
            mb[0] = (box[0] - gt[0]) / gh;
            mb[1] = (box[1] - gt[1]) / gw;
            mb[2] = (box[2] - gt[0]) / gh;
            mb[3] = (box[3] - gt[1]) / gw;
        }
        const int ch = gtm[g];
        const float hs = (mask_h > 1) ? (mb[2] - mb[0]) * (float)(MH - 1) / (float)(mask_h - 1) : 0.0f;
        const float ws = (mask_w > 1) ? (mb[3] - mb[1]) * (float)(MW - 1) / (float)(mask_w - 1) : 0.0f;
        float* mo = masks + (size_t)s * mask_h * mask_w;
        for (int y = 0; y < mask_h; ++y) {
            orc_tap ty = orc_make_tap(mb[0], mb[2], MH, mask_h, y, hs);
            if (!ty.valid) continue;
            for (int x = 0; x < mask_w; ++x) {
                orc_tap tx = orc_make_tap(mb[1], mb[3], MW, mask_w, x, ws);
                if (!tx.valid) continue;
                const float tl = (float)(gt_masks[((size_t)ty.lo * MW + tx.lo) * G + ch] != 0);
                const float tr = (float)(gt_masks[((size_t)ty.lo * MW + tx.hi) * G + ch] != 0);
                const float bl = (float)(gt_masks[((size_t)ty.hi * MW + tx.lo) * G + ch] != 0);
                const float br = (float)(gt_masks[((size_t)ty.hi * MW + tx.hi) * G + ch] != 0);
                const float top = tl + (tr - tl) * tx.lerp;
                const float bot = bl + (br - bl) * tx.lerp;
                mo[y * mask_w + x] = rintf(top + (bot - top) * ty.lerp); /* tf.round L:954 */
            }
        }
    }
    for (int s = 0; s < nneg && npos + s < T; ++s) memcpy(rois + 4 * (npos + s), prop + 4 * neg[s], sizeof(float) * 4);
    if (counts) { counts[0] = npos; counts[1] = nneg; }
    free(prop); free(prop_row); free(gtb); free(gtc); free(gtm); free(crowd); free(iou_max); free(iou_arg);
    free(pos); free(neg);
}

ORC_API void orc_detection_target_layer(const float* proposals /*[B,P,4]*/, const int32_t* gt_class_ids /*[B,G]*/,
                                        const float* gt_boxes /*[B,G,4]*/, const uint8_t* gt_masks /*[B,MH,MW,G]*/,
                                        const uint32_t* rand_keys /*[B,P]*/, int B, int P, int G, int MH, int MW, int T,
                                        double roi_positive_ratio, const float* bbox_std_dev, int mask_h, int mask_w,
                                        int use_mini_masks, float* rois, int32_t* class_ids, float* deltas,
                                        float* masks, int32_t* counts /*[B,2] or NULL*/) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        orc_detection_targets_one(proposals + (size_t)b * P * 4, gt_class_ids + (size_t)b * G,
                                  gt_boxes + (size_t)b * G * 4, gt_masks + (size_t)b * MH * MW * G,
                                  rand_keys + (size_t)b * P, P, G, MH, MW, T, roi_positive_ratio, bbox_std_dev,
                                  mask_h, mask_w, use_mini_masks, rois + (size_t)b * T * 4,
                                  class_ids + (size_t)b * T, deltas + (size_t)b * T * 4,
                                  masks + (size_t)b * T * mask_h * mask_w, counts ? counts + 2 * b : NULL);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* utils.build_rpn_targets (U:154-262) + compute_overlaps / compute_iou (U:114-151), per image  */
/* of a padded batch.  PINNED: unlike the TF kernels above, this function's reference is plain  */
/* numpy and runs here; tests/golden/reference_rpn_targets_golden.npz holds its own outputs.    */
/* float64 throughout, as numpy computes it (anchors float64 pixel boxes, GT boxes int32).      */
/* Padding rows have class id 0 (the loader passes only real instances, preprocess.py:342-348). */
/* np.random.choice (U:219,227) is replaced by injected keys: of the candidate anchors keep the */
/* ones with the largest key (ties -> lower anchor index).                                      */
/* ------------------------------------------------------------------------------------------ */
static double orc_np_iou(const double* a /*anchor*/, double area_a, const int32_t* g, double area_g) {
    /* U:125-133: maximum/minimum, max(.,0) products, union = area_g + area_a - inter, true division */
    const double y1 = (double)g[0] > a[0] ? (double)g[0] : a[0];
    const double y2 = (double)g[2] < a[2] ? (double)g[2] : a[2];
    const double x1 = (double)g[1] > a[1] ? (double)g[1] : a[1];
    const double x2 = (double)g[3] < a[3] ? (double)g[3] : a[3];
    const double dx = x2 - x1 > 0.0 ? x2 - x1 : 0.0, dy = y2 - y1 > 0.0 ? y2 - y1 : 0.0;
    const double inter = dx * dy;
    const double uni = area_g + area_a - inter;
    return inter / uni;
}

typedef struct { float key; int32_t idx; } orc_keyed;
static int orc_keyed_cmp(const void* pa, const void* pb) { /* key desc, idx asc */
    const orc_keyed* a = (const orc_keyed*)pa; const orc_keyed* b = (const orc_keyed*)pb;
    if (a->key != b->key) return a->key > b->key ? -1 : 1;
    return a->idx < b->idx ? -1 : (a->idx > b->idx ? 1 : 0);
}
/* reset all but the `keep` best-keyed anchors whose match equals `label` to neutral (U:215-228) */
static void orc_subsample(int32_t* match, int A, int32_t label, int keep, const float* keys) {
    int n = 0;
    for (int a = 0; a < A; ++a) n += match[a] == label;
    if (n - keep <= 0) return;
    orc_keyed* c = (orc_keyed*)malloc(sizeof(orc_keyed) * (size_t)n);
    int m = 0;
    for (int a = 0; a < A; ++a) if (match[a] == label) { c[m].key = keys[a]; c[m].idx = a; ++m; }
    qsort(c, (size_t)n, sizeof(orc_keyed), orc_keyed_cmp);
    for (int i = keep < 0 ? 0 : keep; i < n; ++i) match[c[i].idx] = 0;
    free(c);
}

static void orc_build_rpn_targets_one(const double* anchors /*[A,4]*/, const int32_t* cls_in /*[G]*/,
                                      const int32_t* box_in /*[G,4]*/, const float* keys /*[A]*/, int A, int G, int R,
                                      const double* std_dev, double eps, int32_t* match /*[A]*/,
                                      double* rpn_bbox /*[R,4]*/, int32_t* counts /*[2] or NULL*/) {
    memset(match, 0, sizeof(int32_t) * (size_t)A);                       /* U:168 */
    memset(rpn_bbox, 0, sizeof(double) * (size_t)R * 4);                 /* U:170 */
    /* U:175-188: crowds (class < 0) are split off only when there is at least one */
    int any_crowd = 0;
    for (int g = 0; g < G; ++g) any_crowd |= cls_in[g] < 0;
    int* gt = (int*)malloc(sizeof(int) * (size_t)(G > 0 ? G : 1));
    int* crowd = (int*)malloc(sizeof(int) * (size_t)(G > 0 ? G : 1));
    int ng = 0, ncrowd = 0;
    for (int g = 0; g < G; ++g) {
        if (cls_in[g] == 0) continue;                                    /* padding row */
        if (cls_in[g] < 0) crowd[ncrowd++] = g;
        else if (!any_crowd || cls_in[g] > 0) gt[ng++] = g;
    }
    double* area_g = (double*)malloc(sizeof(double) * (size_t)(G > 0 ? G : 1));
    for (int g = 0; g < G; ++g) /* U:143: int32 products in numpy; exact in double for pixel boxes */
        area_g[g] = (double)((box_in[4 * g + 2] - box_in[4 * g]) * (box_in[4 * g + 3] - box_in[4 * g + 1]));
    double* ov = (double*)malloc(sizeof(double) * (size_t)A * (size_t)(ng > 0 ? ng : 1));
    double* colmax = (double*)malloc(sizeof(double) * (size_t)(ng > 0 ? ng : 1));
    int32_t* amax_i = (int32_t*)malloc(sizeof(int32_t) * (size_t)A);
    for (int j = 0; j < ng; ++j) colmax[j] = -1.0;
    for (int a = 0; a < A; ++a) {
        const double* an = anchors + 4 * (size_t)a;
        const double area_a = (an[2] - an[0]) * (an[3] - an[1]);         /* U:142 */
        int no_crowd = 1;                                                /* U:184-188 */
        if (ncrowd > 0) {
            double cm = -1.0;
            for (int j = 0; j < ncrowd; ++j) {
                const double v = orc_np_iou(an, area_a, box_in + 4 * crowd[j], area_g[crowd[j]]);
                if (v > cm) cm = v;
            }
            no_crowd = cm < 0.001;
        }
        double best = 0.0; int besti = 0;                                /* U:203-204: argmax = first maximum */
        for (int j = 0; j < ng; ++j) {
            const double v = orc_np_iou(an, area_a, box_in + 4 * gt[j], area_g[gt[j]]);
            ov[(size_t)a * ng + j] = v;
            if (j == 0 || v > best) { best = v; besti = j; }
            if (v > colmax[j]) colmax[j] = v;
        }
        amax_i[a] = besti;
        if (best < 0.3 && no_crowd) match[a] = -1;                       /* U:205 */
    }
    for (int a = 0; a < A; ++a) {                                        /* U:208-209: every anchor tying a column max */
        for (int j = 0; j < ng; ++j) if (ov[(size_t)a * ng + j] == colmax[j]) { match[a] = 1; break; }
    }
    for (int a = 0; a < A; ++a) {                                        /* U:211 */
        if (ng > 0 && ov[(size_t)a * ng + amax_i[a]] >= 0.7) match[a] = 1;
    }
    orc_subsample(match, A, 1, R / 2, keys);                             /* U:215-220 */
    int npos = 0;
    for (int a = 0; a < A; ++a) npos += match[a] == 1;
    orc_subsample(match, A, -1, R - npos, keys);                         /* U:222-228 */
    int ix = 0, nneg = 0;
    for (int a = 0; a < A; ++a) {                                        /* U:232-260, ascending anchor index */
        nneg += match[a] == -1;
        if (match[a] != 1) continue;
        const int32_t* g = box_in + 4 * gt[amax_i[a]];
        const double* an = anchors + 4 * (size_t)a;
        const double gt_h = (double)(g[2] - g[0]), gt_w = (double)(g[3] - g[1]);
        const double gcy = (double)g[0] + 0.5 * gt_h, gcx = (double)g[1] + 0.5 * gt_w;
        const double a_h = an[2] - an[0], a_w = an[3] - an[1];
        const double acy = an[0] + 0.5 * a_h, acx = an[1] + 0.5 * a_w;
        double* o = rpn_bbox + 4 * (size_t)ix;
        o[0] = (gcy - acy) / a_h / std_dev[0];
        o[1] = (gcx - acx) / a_w / std_dev[1];
        o[2] = log(gt_h / (a_h + eps)) / std_dev[2];
        o[3] = log(gt_w / (a_w + eps)) / std_dev[3];
        ++ix;
    }
    if (counts) { counts[0] = npos; counts[1] = nneg; }
    free(gt); free(crowd); free(area_g); free(ov); free(colmax); free(amax_i);
}

ORC_API void orc_build_rpn_targets(const double* anchors /*[A,4]*/, const int32_t* gt_class_ids /*[B,G]*/,
                                   const int32_t* gt_boxes /*[B,G,4]*/, const float* rand_keys /*[B,A]*/, int B,
                                   int A, int G, int R, const double* std_dev, double eps,
                                   int32_t* rpn_match /*[B,A]*/, double* rpn_bbox /*[B,R,4]*/,
                                   int32_t* counts /*[B,2] or NULL*/) {
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < B; ++b)
        orc_build_rpn_targets_one(anchors, gt_class_ids + (size_t)b * G, gt_boxes + (size_t)b * G * 4,
                                  rand_keys + (size_t)b * A, A, G, R, std_dev, eps, rpn_match + (size_t)b * A,
                                  rpn_bbox + (size_t)b * R * 4, counts ? counts + 2 * b : NULL);
}

/* Keras Activation("softmax") over the two RPN logits (mrcnn_layers.py:1081) = tf.nn.softmax = TF SoftmaxEigenImpl
 * (tensorflow/core/kernels/softmax_op_functor.h): shifted = logits - max; e = exp(shifted); p = e * (1 / sum(e)).
 * PARITY UNPINNED like the other TF kernels; exp is the shared deterministic sequence. */
ORC_API void orc_softmax2(const float* logits /*[n,2]*/, int n, float* probs /*[n,2]*/) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        const float a = logits[2 * i], b = logits[2 * i + 1];
        const float m = orc_maxf(a, b);
        const float e0 = orc_expf(a - m), e1 = orc_expf(b - m);
        const float inv = 1.0f / (e0 + e1);
        probs[2 * i] = e0 * inv;
        probs[2 * i + 1] = e1 * inv;
    }
}

ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void)n;
#endif
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
