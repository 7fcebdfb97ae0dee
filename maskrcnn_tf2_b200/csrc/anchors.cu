// anchors.cu -- the model's anchor constant on the device: utils.generate_pyramid_anchors / generate_anchors
// (utils.py:54-111) followed by AnchorsLayer.get_anchors + NormBoxesLayer.call (mrcnn_layers.py:34-39,116-132).
// Row a13 of the scope table: the producer whose bit patterns ProposalLayer and build_rpn_targets consume.
//   float64, numpy's operation order: heights = scale / sqrt(ratio), widths = scale * sqrt(ratio) (IEEE sqrt and
//   division), centres = index * anchor_stride * feature_stride (integers), corners = centre -/+ 0.5 * size;
//   order: level-major, row-major (y, x) over the level's cells, ratio innermost (the reshape at utils.py:89-92);
//   then Keras' fp32 autocast and (a - [0,0,1,1]) / ([h,w,h,w] - 1) in fp32, broadcast to the batch.
// One thread per anchor; pinned bit-exactly by the digests of the reference's own output (tests/test_reference_pins.py).
#include "common.cuh"

namespace mrcnn {
namespace {

constexpr int kMaxLevels = 8, kMaxRatios = 8;

struct AnchorSpec {
    int levels, nratios, anchor_stride;
    int start[kMaxLevels + 1];   // first anchor of each level
    int cells_x[kMaxLevels];     // anchor columns per row of the level
    int stride[kMaxLevels];
    double scale[kMaxLevels];
    double ratio[kMaxRatios];
};

__global__ void __launch_bounds__(256)
anchors_kernel(AnchorSpec sp, int A, int B, float img_h, float img_w, double* __restrict__ anchors_px,
               float4* __restrict__ anchors_norm) {
    const int a = blockIdx.x * 256 + threadIdx.x;
    if (a >= A) return;
    int l = 0;
    while (l + 1 < sp.levels && a >= sp.start[l + 1]) ++l;
    const int local = a - sp.start[l];
    const int r = local % sp.nratios, cell = local / sp.nratios;
    const int yi = cell / sp.cells_x[l], xi = cell - yi * sp.cells_x[l];
    const double sq = sqrt(sp.ratio[r]);                                   // utils.py:73-74
    const double h = __ddiv_rn(sp.scale[l], sq), w = __dmul_rn(sp.scale[l], sq);
    const double cy = (double)(yi * sp.anchor_stride * sp.stride[l]);      // utils.py:77-78 (integers)
    const double cx = (double)(xi * sp.anchor_stride * sp.stride[l]);
    const double hh = __dmul_rn(0.5, h), hw = __dmul_rn(0.5, w);           // utils.py:95-96
    const double y1 = __dsub_rn(cy, hh), x1 = __dsub_rn(cx, hw), y2 = __dadd_rn(cy, hh), x2 = __dadd_rn(cx, hw);
    if (anchors_px) {
        double2* o = reinterpret_cast<double2*>(anchors_px + 4 * (size_t)a);
        o[0] = make_double2(y1, x1);
        o[1] = make_double2(y2, x2);
    }
    if (anchors_norm) {
        const float sh = __fsub_rn(img_h, 1.0f), sw = __fsub_rn(img_w, 1.0f);   // mrcnn_layers.py:37-38
        const float4 v = make_float4(__fdiv_rn((float)y1, sh), __fdiv_rn((float)x1, sw),
                                     __fdiv_rn(__fsub_rn((float)y2, 1.0f), sh), __fdiv_rn(__fsub_rn((float)x2, 1.0f), sw));
        for (int b = 0; b < B; ++b) anchors_norm[(size_t)b * A + a] = v;        // np.broadcast_to (mrcnn_layers.py:132)
    }
}

int anchor_spec(const double* scales, const double* ratios, const int* feat_h, const int* feat_w, const int* strides,
                int levels, int nratios, int anchor_stride, AnchorSpec* sp) {
    if (!scales || !ratios || !feat_h || !feat_w || !strides) return MRCNN_ERR_NULL;
    if (levels < 1 || levels > kMaxLevels || nratios < 1 || nratios > kMaxRatios || anchor_stride < 1)
        return MRCNN_ERR_RANGE;
    sp->levels = levels; sp->nratios = nratios; sp->anchor_stride = anchor_stride;
    long long acc = 0;
    for (int l = 0; l < levels; ++l) {
        if (feat_h[l] < 1 || feat_w[l] < 1 || strides[l] < 1) return MRCNN_ERR_RANGE;
        const int ny = (feat_h[l] + anchor_stride - 1) / anchor_stride;    // len(np.arange(0, shape, anchor_stride))
        const int nx = (feat_w[l] + anchor_stride - 1) / anchor_stride;
        sp->start[l] = (int)acc;
        sp->cells_x[l] = nx;
        sp->stride[l] = strides[l];
        sp->scale[l] = scales[l];
        acc += (long long)ny * nx * nratios;
        if (acc > (1 << 24)) return MRCNN_ERR_RANGE;
    }
    sp->start[levels] = (int)acc;
    for (int r = 0; r < nratios; ++r) sp->ratio[r] = ratios[r];
    return MRCNN_OK;
}

}  // namespace
}  // namespace mrcnn

using namespace mrcnn;

MRCNN_EXPORT int mrcnn_anchors_count(const int* feat_h, const int* feat_w, int levels, int nratios, int anchor_stride,
                                     int* count) {
    if (!feat_h || !feat_w || !count) return MRCNN_ERR_NULL;
    if (levels < 1 || levels > kMaxLevels || nratios < 1 || anchor_stride < 1) return MRCNN_ERR_RANGE;
    long long acc = 0;
    for (int l = 0; l < levels; ++l)
        acc += (long long)((feat_h[l] + anchor_stride - 1) / anchor_stride) *
               ((feat_w[l] + anchor_stride - 1) / anchor_stride) * nratios;
    if (acc > (1 << 24)) return MRCNN_ERR_RANGE;
    *count = (int)acc;
    return MRCNN_OK;
}

MRCNN_EXPORT int mrcnn_anchors_forward(const double* scales, const double* ratios, const int* feat_h, const int* feat_w,
                                       const int* strides, int levels, int nratios, int anchor_stride, int img_h,
                                       int img_w, int B, double* anchors_px, float* anchors_norm, void* stream) {
    AnchorSpec sp;
    const int rc = anchor_spec(scales, ratios, feat_h, feat_w, strides, levels, nratios, anchor_stride, &sp);
    if (rc != MRCNN_OK) return rc;
    if (!anchors_px && !anchors_norm) return MRCNN_ERR_NULL;
    if (img_h < 2 || img_w < 2 || (anchors_norm && B < 1)) return MRCNN_ERR_RANGE;
    if (!aligned16(anchors_px) || !aligned16(anchors_norm)) return MRCNN_ERR_ALIGN;
    const int A = sp.start[levels];
    anchors_kernel<<<(A + 255) / 256, 256, 0, (cudaStream_t)stream>>>(sp, A, B, (float)img_h, (float)img_w, anchors_px,
                                                                      (float4*)anchors_norm);
    return last_error();
}
