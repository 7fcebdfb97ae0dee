// mrcnn_roi_ops.cc -- TensorFlow custom-op shim over the extern "C" launchers of libmrcnn_roi_b200.so.
//
// NOT BUILT IN THIS REPOSITORY'S CI: TensorFlow is not installable in the build image (no network; the
// reference pins tensorflow==2.2-2.5, whose wheels cannot drive sm_100 anyway).  Build it where a CUDA-12
// TensorFlow exists:
//   g++ -std=c++17 -shared -fPIC mrcnn_roi_ops.cc -o libmrcnn_roi_ops.so
//       $(python -c "import tensorflow as tf; print(' '.join(tf.sysconfig.get_compile_flags()))")
//       $(python -c "import tensorflow as tf; print(' '.join(tf.sysconfig.get_link_flags()))")
//       -I../../include -L.. -lmrcnn_roi_b200 -DGOOGLE_CUDA=1        (one command line)
// What IS checked here: the file is compiled against tests/tf_stub/ (a functional stand-in for the TF op API), which
// type-checks every launcher call against include/mrcnn_roi_b200.h, and each OpKernel::Compute below is executed on
// the B200 through a fake OpKernelContext and compared with the ctypes path (tests/test_tf_shim_stub.py).
//
// Each kernel only validates shapes, allocates outputs + one scratch buffer through TF's allocator, fetches
// TF's CUDA stream and calls the launcher; there is deliberately no CPU kernel registration.
#define EIGEN_USE_GPU
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "mrcnn_roi_b200.h"

namespace tf = tensorflow;
using tf::shape_inference::InferenceContext;
using GPUDevice = Eigen::GpuDevice;

namespace {

tf::Status LauncherStatus(int rc, const char* what) {
  if (rc == 0) return tf::Status();
  if (rc < 0) return tf::errors::InvalidArgument(what, ": ", mrcnn_status_string(rc));
  return tf::errors::Internal(what, ": CUDA error ", rc, " (", mrcnn_status_string(rc), ")");
}

void* StreamOf(tf::OpKernelContext* ctx) { return ctx->eigen_device<GPUDevice>().stream(); }

// Attributes are fp32, but the reference computes with the Python floats the user wrote (0.1, 0.33, 1e-3 as float64:
// config.py:90, mrcnn_layers.py:904, utils.py:794).  Recover them: the shortest decimal that reads back as the same
// fp32 is what the user typed (0.1f -> 0.1, not 0.100000001490116).
double AttrAsDouble(float v) {
  char buf[32];
  for (int prec = 1; prec <= 9; ++prec) {
    std::snprintf(buf, sizeof(buf), "%.*g", prec, static_cast<double>(v));
    if (std::strtof(buf, nullptr) == v) break;
  }
  return std::strtod(buf, nullptr);
}

tf::Status Scratch(tf::OpKernelContext* ctx, size_t bytes, tf::Tensor* t) {
  return ctx->allocate_temp(tf::DT_UINT8, tf::TensorShape({static_cast<tf::int64>(bytes)}), t);
}

}  // namespace

// ---- ProposalLayer.call (mrcnn_layers.py:233-269) ------------------------------------------------------
REGISTER_OP("MrcnnProposal")
    .Input("rpn_probs: float")   // [B,A,2]
    .Input("rpn_bbox: float")    // [B,A,4]
    .Input("anchors: float")     // [B,A,4]
    .Output("proposals: float")  // [B,P,4]
    .Output("topk_idx: int32")   // [B,K] anchors selected by tf.nn.top_k (L:246), saved for the gradient
    .Output("keep_idx: int32")   // [B,P] NMS survivors as positions in the top-k list, -1 padded (saved too)
    .Attr("proposal_count: int")
    .Attr("pre_nms_limit: int = 6000")
    .Attr("nms_threshold: float = 0.7")
    .Attr("std_dev: list(float) = [0.1, 0.1, 0.2, 0.2]")
    .SetShapeFn([](InferenceContext* c) {
      int p;
      TF_RETURN_IF_ERROR(c->GetAttr("proposal_count", &p));
      c->set_output(0, c->MakeShape({c->Dim(c->input(0), 0), p, 4}));
      c->set_output(1, c->MakeShape({c->Dim(c->input(0), 0), c->UnknownDim()}));
      c->set_output(2, c->MakeShape({c->Dim(c->input(0), 0), p}));
      return tf::Status();
    });

class MrcnnProposalOp : public tf::OpKernel {
 public:
  explicit MrcnnProposalOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("proposal_count", &p_));
    OP_REQUIRES_OK(c, c->GetAttr("pre_nms_limit", &pre_));
    OP_REQUIRES_OK(c, c->GetAttr("nms_threshold", &thr_));
    OP_REQUIRES_OK(c, c->GetAttr("std_dev", &std_));
    OP_REQUIRES(c, std_.size() == 4, tf::errors::InvalidArgument("std_dev needs 4 values"));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& probs = ctx->input(0);
    const tf::Tensor& bbox = ctx->input(1);
    const tf::Tensor& anchors = ctx->input(2);
    OP_REQUIRES(ctx, probs.dims() == 3 && probs.dim_size(2) == 2, tf::errors::InvalidArgument("rpn_probs [B,A,2]"));
    const int B = probs.dim_size(0), A = probs.dim_size(1);
    OP_REQUIRES(ctx, bbox.shape() == tf::TensorShape({B, A, 4}) && anchors.shape() == bbox.shape(),
                tf::errors::InvalidArgument("rpn_bbox / anchors must be [B,A,4]"));
    tf::Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, p_, 4}), &out));
    const int K = pre_ < A ? pre_ : A;  // L:245
    tf::Tensor* topk = nullptr;
    tf::Tensor* keep = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({B, K}), &topk));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, tf::TensorShape({B, p_}), &keep));
    size_t ws_bytes = 0;
    OP_REQUIRES_OK(ctx, LauncherStatus(mrcnn_proposal_workspace_bytes(B, A, pre_, p_, &ws_bytes), "proposal ws"));
    tf::Tensor ws;
    OP_REQUIRES_OK(ctx, Scratch(ctx, ws_bytes, &ws));
    OP_REQUIRES_OK(ctx, LauncherStatus(
        mrcnn_proposal_forward(probs.flat<float>().data(), bbox.flat<float>().data(), anchors.flat<float>().data(), B,
                               A, pre_, p_, std_.data(), thr_, out->flat<float>().data(),
                               topk->flat<tf::int32>().data(), keep->flat<tf::int32>().data(), nullptr, nullptr,
                               ws.flat<tf::uint8>().data(), ws_bytes, StreamOf(ctx)),
        "mrcnn_proposal_forward"));
  }

 private:
  int p_, pre_;
  float thr_;
  std::vector<float> std_;
};
REGISTER_KERNEL_BUILDER(Name("MrcnnProposal").Device(tf::DEVICE_GPU), MrcnnProposalOp);

// Gradient of ProposalLayer.call w.r.t. rpn_bbox: the reference does not stop it (model.py:155-157,168 back-propagate
// mrcnn_bbox_loss through the proposals: gather L:227 -> clip utils.py:854-869 -> decode utils.py:830-851 -> L:238).
REGISTER_OP("MrcnnProposalGrad")
    .Input("grad_proposals: float")  // [B,P,4]
    .Input("rpn_bbox: float")        // [B,A,4]
    .Input("anchors: float")         // [B,A,4]
    .Input("topk_idx: int32")        // [B,K]
    .Input("keep_idx: int32")        // [B,P]
    .Output("grad_rpn_bbox: float")  // [B,A,4]
    .Attr("std_dev: list(float) = [0.1, 0.1, 0.2, 0.2]")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->input(1));
      return tf::Status();
    });

class MrcnnProposalGradOp : public tf::OpKernel {
 public:
  explicit MrcnnProposalGradOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("std_dev", &std_));
    OP_REQUIRES(c, std_.size() == 4, tf::errors::InvalidArgument("std_dev needs 4 values"));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& grad = ctx->input(0);
    const tf::Tensor& bbox = ctx->input(1);
    const tf::Tensor& anchors = ctx->input(2);
    const tf::Tensor& topk = ctx->input(3);
    const tf::Tensor& keep = ctx->input(4);
    OP_REQUIRES(ctx, bbox.dims() == 3 && bbox.dim_size(2) == 4 && anchors.shape() == bbox.shape(),
                tf::errors::InvalidArgument("rpn_bbox / anchors must be [B,A,4]"));
    OP_REQUIRES(ctx, topk.dims() == 2 && keep.dims() == 2 && topk.dim_size(0) == bbox.dim_size(0) &&
                         keep.dim_size(0) == bbox.dim_size(0),
                tf::errors::InvalidArgument("topk_idx [B,K] and keep_idx [B,P] must have the batch size of rpn_bbox"));
    const int B = bbox.dim_size(0), A = bbox.dim_size(1), K = topk.dim_size(1), P = keep.dim_size(1);
    OP_REQUIRES(ctx, grad.shape() == tf::TensorShape({B, P, 4}), tf::errors::InvalidArgument("grad_proposals [B,P,4]"));
    tf::Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, bbox.shape(), &out));
    OP_REQUIRES_OK(ctx, LauncherStatus(
        mrcnn_proposal_backward(grad.flat<float>().data(), bbox.flat<float>().data(), anchors.flat<float>().data(),
                                topk.flat<tf::int32>().data(), keep.flat<tf::int32>().data(), B, A, K, P, std_.data(),
                                out->flat<float>().data(), StreamOf(ctx)),
        "mrcnn_proposal_backward"));
  }

 private:
  std::vector<float> std_;
};
REGISTER_KERNEL_BUILDER(Name("MrcnnProposalGrad").Device(tf::DEVICE_GPU), MrcnnProposalGradOp);

// ---- PyramidROIAlign.call (mrcnn_layers.py:583-664) and its feature-map gradient ------------------------
REGISTER_OP("MrcnnPyramidRoiAlign")
    .Input("boxes: float")       // [B,N,4]
    .Input("image_meta: float")  // [B,meta]
    .Input("p2: float").Input("p3: float").Input("p4: float").Input("p5: float")  // [B,H,W,C]
    .Output("pooled: float")     // [B,N,ph,pw,C]
    .Output("roi_map: int32")    // [B,N] (saved for the gradient)
    .Attr("pool_height: int").Attr("pool_width: int")
    .Attr("denominator: float = 244.0")
    .Attr("map_mode: int = 0")
    .SetShapeFn([](InferenceContext* c) {
      int ph, pw;
      TF_RETURN_IF_ERROR(c->GetAttr("pool_height", &ph));
      TF_RETURN_IF_ERROR(c->GetAttr("pool_width", &pw));
      c->set_output(0, c->MakeShape({c->Dim(c->input(0), 0), c->Dim(c->input(0), 1), ph, pw, c->Dim(c->input(2), 3)}));
      c->set_output(1, c->MakeShape({c->Dim(c->input(0), 0), c->Dim(c->input(0), 1)}));
      return tf::Status();
    });

class MrcnnPyramidRoiAlignOp : public tf::OpKernel {
 public:
  explicit MrcnnPyramidRoiAlignOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("pool_height", &ph_));
    OP_REQUIRES_OK(c, c->GetAttr("pool_width", &pw_));
    OP_REQUIRES_OK(c, c->GetAttr("denominator", &den_));
    OP_REQUIRES_OK(c, c->GetAttr("map_mode", &mode_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& boxes = ctx->input(0);
    const tf::Tensor& meta = ctx->input(1);
    OP_REQUIRES(ctx, boxes.dims() == 3 && boxes.dim_size(2) == 4, tf::errors::InvalidArgument("boxes [B,N,4]"));
    const int B = boxes.dim_size(0), N = boxes.dim_size(1);
    OP_REQUIRES(ctx, meta.dims() == 2 && meta.dim_size(0) == B && meta.dim_size(1) >= 6,
                tf::errors::InvalidArgument("image_meta [B,meta] with meta >= 6 and the batch size of boxes"));
    const float* maps[4];
    int H[4], W[4];
    OP_REQUIRES(ctx, ctx->input(2).dims() == 4, tf::errors::InvalidArgument("feature maps must be [B,H,W,C]"));
    const int C = ctx->input(2).dim_size(3);
    for (int l = 0; l < 4; ++l) {
      const tf::Tensor& m = ctx->input(2 + l);
      OP_REQUIRES(ctx, m.dims() == 4 && m.dim_size(0) == B && m.dim_size(3) == C,
                  tf::errors::InvalidArgument("feature maps must be [B,H,W,C]"));
      maps[l] = m.flat<float>().data();
      H[l] = m.dim_size(1);
      W[l] = m.dim_size(2);
    }
    tf::Tensor *out = nullptr, *roi_map = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, N, ph_, pw_, C}), &out));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({B, N}), &roi_map));
    size_t ws_bytes = 0;
    OP_REQUIRES_OK(ctx, LauncherStatus(mrcnn_roialign_workspace_bytes(B, N, &ws_bytes), "roialign ws"));
    tf::Tensor ws;
    OP_REQUIRES_OK(ctx, Scratch(ctx, ws_bytes, &ws));
    OP_REQUIRES_OK(ctx, LauncherStatus(
        mrcnn_roialign_forward(boxes.flat<float>().data(), meta.flat<float>().data(), meta.dim_size(1), maps, H, W, C,
                               B, N, ph_, pw_, den_, mode_, out->flat<float>().data(),
                               roi_map->flat<tf::int32>().data(), nullptr, ws.flat<tf::uint8>().data(), ws_bytes,
                               StreamOf(ctx)),
        "mrcnn_roialign_forward"));
  }

 private:
  int ph_, pw_, mode_;
  float den_;
};
REGISTER_KERNEL_BUILDER(Name("MrcnnPyramidRoiAlign").Device(tf::DEVICE_GPU), MrcnnPyramidRoiAlignOp);

REGISTER_OP("MrcnnPyramidRoiAlignGrad")
    .Input("grad: float")        // [B,N,ph,pw,C]
    .Input("boxes: float")
    .Input("roi_map: int32")
    .Input("p2: float").Input("p3: float").Input("p4: float").Input("p5: float")  // shapes only
    .Output("d2: float").Output("d3: float").Output("d4: float").Output("d5: float")
    .SetShapeFn([](InferenceContext* c) {
      for (int l = 0; l < 4; ++l) c->set_output(l, c->input(3 + l));
      return tf::Status();
    });

class MrcnnPyramidRoiAlignGradOp : public tf::OpKernel {
 public:
  using tf::OpKernel::OpKernel;
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& grad = ctx->input(0);
    const tf::Tensor& boxes = ctx->input(1);
    const tf::Tensor& roi_map = ctx->input(2);
    OP_REQUIRES(ctx, grad.dims() == 5, tf::errors::InvalidArgument("grad [B,N,ph,pw,C]"));
    const int B = grad.dim_size(0), N = grad.dim_size(1), ph = grad.dim_size(2), pw = grad.dim_size(3);
    const int C = grad.dim_size(4);
    OP_REQUIRES(ctx, boxes.shape() == tf::TensorShape({B, N, 4}) && roi_map.shape() == tf::TensorShape({B, N}),
                tf::errors::InvalidArgument("boxes [B,N,4] and roi_map [B,N] must match grad [B,N,ph,pw,C]"));
    float* grads[4];
    int H[4], W[4];
    for (int l = 0; l < 4; ++l) {
      const tf::Tensor& fm = ctx->input(3 + l);
      OP_REQUIRES(ctx, fm.dims() == 4 && fm.dim_size(0) == B && fm.dim_size(3) == C,
                  tf::errors::InvalidArgument("feature maps must be [B,H,W,C] with the batch size and channels of grad"));
      tf::Tensor* g = nullptr;
      OP_REQUIRES_OK(ctx, ctx->allocate_output(l, ctx->input(3 + l).shape(), &g));
      grads[l] = g->flat<float>().data();
      H[l] = g->dim_size(1);
      W[l] = g->dim_size(2);
    }
    // deterministic mode: every gradient pixel is summed in CropAndResizeGradImage's sequential order
    size_t ws_bytes = 0;
    OP_REQUIRES_OK(ctx, LauncherStatus(mrcnn_roialign_backward_workspace_bytes(B, N, ph, pw, H, W, C, &ws_bytes),
                                       "mrcnn_roialign_backward_workspace_bytes"));
    tf::Tensor ws;
    OP_REQUIRES_OK(ctx, Scratch(ctx, ws_bytes, &ws));
    OP_REQUIRES_OK(ctx, LauncherStatus(
        mrcnn_roialign_backward(grad.flat<float>().data(), boxes.flat<float>().data(),
                                roi_map.flat<tf::int32>().data(), grads, H, W, C, B, N, ph, pw,
                                ws.flat<tf::uint8>().data(), ws_bytes, StreamOf(ctx)),
        "mrcnn_roialign_backward"));
  }
};
REGISTER_KERNEL_BUILDER(Name("MrcnnPyramidRoiAlignGrad").Device(tf::DEVICE_GPU), MrcnnPyramidRoiAlignGradOp);

// ---- DetectionLayer.call (mrcnn_layers.py:369-524) ------------------------------------------------------
REGISTER_OP("MrcnnDetection")
    .Input("rois: float").Input("mrcnn_class: float").Input("mrcnn_bbox: float").Input("image_meta: float")
    .Output("detections: float")  // [B,max_instances,6]
    .Output("boxes: float")       // [B,max_instances,4] = detections[..., :4] (DetectedBoxesExtraction, L:535-550)
    .Attr("min_confidence: float = 0.7")
    .Attr("use_min_confidence: bool = true")
    .Attr("max_instances: int = 100")
    .Attr("nms_threshold: float = 0.3")
    .Attr("std_dev: list(float) = [0.1, 0.1, 0.2, 0.2]")
    .SetShapeFn([](InferenceContext* c) {
      int d;
      TF_RETURN_IF_ERROR(c->GetAttr("max_instances", &d));
      c->set_output(0, c->MakeShape({c->Dim(c->input(0), 0), d, 6}));
      c->set_output(1, c->MakeShape({c->Dim(c->input(0), 0), d, 4}));
      return tf::Status();
    });

class MrcnnDetectionOp : public tf::OpKernel {
 public:
  explicit MrcnnDetectionOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("min_confidence", &conf_));
    OP_REQUIRES_OK(c, c->GetAttr("use_min_confidence", &use_conf_));
    OP_REQUIRES_OK(c, c->GetAttr("max_instances", &d_));
    OP_REQUIRES_OK(c, c->GetAttr("nms_threshold", &thr_));
    OP_REQUIRES_OK(c, c->GetAttr("std_dev", &std_));
    OP_REQUIRES(c, std_.size() == 4, tf::errors::InvalidArgument("std_dev needs 4 values"));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& rois = ctx->input(0);
    const tf::Tensor& probs = ctx->input(1);
    const tf::Tensor& deltas = ctx->input(2);
    const tf::Tensor& meta = ctx->input(3);
    OP_REQUIRES(ctx, probs.dims() == 3, tf::errors::InvalidArgument("mrcnn_class [B,N,NC]"));
    const int B = probs.dim_size(0), N = probs.dim_size(1), NC = probs.dim_size(2);
    OP_REQUIRES(ctx, rois.shape() == tf::TensorShape({B, N, 4}) && deltas.shape() == tf::TensorShape({B, N, NC, 4}),
                tf::errors::InvalidArgument("rois [B,N,4], mrcnn_bbox [B,N,NC,4]"));
    tf::Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, d_, 6}), &out));
    tf::Tensor* boxes_out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({B, d_, 4}), &boxes_out));
    size_t ws_bytes = 0;
    OP_REQUIRES_OK(ctx, LauncherStatus(mrcnn_detection_workspace_bytes(B, N, NC, &ws_bytes), "detection ws"));
    tf::Tensor ws;
    OP_REQUIRES_OK(ctx, Scratch(ctx, ws_bytes, &ws));
    OP_REQUIRES_OK(ctx, LauncherStatus(
        mrcnn_detection_forward(rois.flat<float>().data(), probs.flat<float>().data(), deltas.flat<float>().data(),
                                meta.flat<float>().data(), meta.dim_size(1), B, N, NC, std_.data(), conf_,
                                use_conf_ ? 1 : 0, d_, thr_, 0, out->flat<float>().data(), nullptr,
                                boxes_out->flat<float>().data(), ws.flat<tf::uint8>().data(), ws_bytes, StreamOf(ctx)),
        "mrcnn_detection_forward"));
  }

 private:
  float conf_, thr_;
  bool use_conf_;
  int d_;
  std::vector<float> std_;
};
REGISTER_KERNEL_BUILDER(Name("MrcnnDetection").Device(tf::DEVICE_GPU), MrcnnDetectionOp);

// ---- DetectionTargetLayer.call (mrcnn_layers.py:313-325, 844-1007) --------------------------------------
REGISTER_OP("MrcnnDetectionTarget")
    .Input("proposals: float").Input("gt_class_ids: int32").Input("gt_boxes: float").Input("gt_masks: bool")
    .Input("rand_keys: int32")    // [B,P] raw 32-bit keys (tf.random.uniform(..., dtype=int32)); stands in for tf.random.shuffle
    .Output("rois: float").Output("class_ids: int32").Output("deltas: float").Output("masks: float")
    .Attr("train_rois_per_image: int = 200")
    .Attr("roi_positive_ratio: float = 0.33")
    .Attr("mask_height: int = 28").Attr("mask_width: int = 28")
    .Attr("use_mini_masks: bool = false")
    .Attr("std_dev: list(float) = [0.1, 0.1, 0.2, 0.2]")
    .SetShapeFn([](InferenceContext* c) {
      int t, mh, mw;
      TF_RETURN_IF_ERROR(c->GetAttr("train_rois_per_image", &t));
      TF_RETURN_IF_ERROR(c->GetAttr("mask_height", &mh));
      TF_RETURN_IF_ERROR(c->GetAttr("mask_width", &mw));
      auto b = c->Dim(c->input(0), 0);
      c->set_output(0, c->MakeShape({b, t, 4}));
      c->set_output(1, c->MakeShape({b, t}));
      c->set_output(2, c->MakeShape({b, t, 4}));
      c->set_output(3, c->MakeShape({b, t, mh, mw}));
      return tf::Status();
    });

class MrcnnDetectionTargetOp : public tf::OpKernel {
 public:
  explicit MrcnnDetectionTargetOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("train_rois_per_image", &t_));
    OP_REQUIRES_OK(c, c->GetAttr("roi_positive_ratio", &ratio_));
    OP_REQUIRES_OK(c, c->GetAttr("mask_height", &mh_));
    OP_REQUIRES_OK(c, c->GetAttr("mask_width", &mw_));
    OP_REQUIRES_OK(c, c->GetAttr("use_mini_masks", &mini_));
    OP_REQUIRES_OK(c, c->GetAttr("std_dev", &std_));
    OP_REQUIRES(c, std_.size() == 4, tf::errors::InvalidArgument("std_dev needs 4 values"));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& props = ctx->input(0);
    const tf::Tensor& cls = ctx->input(1);
    const tf::Tensor& boxes = ctx->input(2);
    const tf::Tensor& masks = ctx->input(3);
    const tf::Tensor& keys = ctx->input(4);
    OP_REQUIRES(ctx, props.dims() == 3 && props.dim_size(1) > 0 && props.dim_size(2) == 4,
                tf::errors::InvalidArgument("roi_assertion: at least one proposal (mrcnn_layers.py:866-868)"));
    OP_REQUIRES(ctx, cls.dims() == 2 && masks.dims() == 4, tf::errors::InvalidArgument("gt_class_ids [B,G], gt_masks [B,H,W,G]"));
    const int B = props.dim_size(0), P = props.dim_size(1), G = cls.dim_size(1);
    const int MH = masks.dim_size(1), MW = masks.dim_size(2);
    OP_REQUIRES(ctx, cls.dim_size(0) == B && boxes.shape() == tf::TensorShape({B, G, 4}) && masks.dim_size(0) == B &&
                         masks.dim_size(3) == G && keys.shape() == tf::TensorShape({B, P}),
                tf::errors::InvalidArgument("expected gt_class_ids [B,G], gt_boxes [B,G,4], gt_masks [B,H,W,G] and rand_keys "
                                            "[B,P] for proposals [B,P,4]"));
    tf::Tensor *rois, *ids, *deltas, *out_masks;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, t_, 4}), &rois));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({B, t_}), &ids));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, tf::TensorShape({B, t_, 4}), &deltas));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, tf::TensorShape({B, t_, mh_, mw_}), &out_masks));
    size_t ws_bytes = 0;
    OP_REQUIRES_OK(ctx, LauncherStatus(mrcnn_detection_target_workspace_bytes(B, P, G, t_, &ws_bytes), "target ws"));
    tf::Tensor ws;
    OP_REQUIRES_OK(ctx, Scratch(ctx, ws_bytes, &ws));
    OP_REQUIRES_OK(ctx, LauncherStatus(
        mrcnn_detection_target_forward(
            props.flat<float>().data(), cls.flat<tf::int32>().data(), boxes.flat<float>().data(),
            reinterpret_cast<const uint8_t*>(masks.flat<bool>().data()),
            reinterpret_cast<const uint32_t*>(keys.flat<tf::int32>().data()), B, P, G, MH, MW, t_,
            // the attr is a float; the reference evaluates int(T * ratio) in Python doubles (L:904): round-trip
            // through the shortest decimal so 0.33f means 0.33
            AttrAsDouble(ratio_), std_.data(), mh_, mw_, mini_ ? 1 : 0, rois->flat<float>().data(),
            ids->flat<tf::int32>().data(), deltas->flat<float>().data(), out_masks->flat<float>().data(), nullptr,
            ws.flat<tf::uint8>().data(), ws_bytes, StreamOf(ctx)),
        "mrcnn_detection_target_forward"));
  }

 private:
  int t_, mh_, mw_;
  float ratio_;
  bool mini_;
  std::vector<float> std_;
};
REGISTER_KERNEL_BUILDER(Name("MrcnnDetectionTarget").Device(tf::DEVICE_GPU), MrcnnDetectionTargetOp);

// ---- utils.build_rpn_targets (utils.py:154-262), batched; called from the input pipeline instead of per image in
// SegmentationDataGenerator (preprocess.py:342-348).  anchors: [A,4] float64 pixel boxes (preprocess.py:82,297).
REGISTER_OP("MrcnnRpnTargets")
    .Input("anchors: double").Input("gt_class_ids: int32").Input("gt_boxes: int32")
    .Input("rand_keys: float")    // [B,A] uniform [0,1): stands in for np.random.choice (utils.py:219,227)
    .Output("rpn_match: int32").Output("rpn_bbox: double").Output("rpn_bbox_f32: float")
    .Attr("rpn_train_anchors_per_image: int = 256")
    .Attr("rpn_bbox_std_dev: list(float) = [0.1, 0.1, 0.2, 0.2]")
    .Attr("eps: float = 0.001")
    .SetShapeFn([](InferenceContext* c) {
      int r;
      TF_RETURN_IF_ERROR(c->GetAttr("rpn_train_anchors_per_image", &r));
      auto b = c->Dim(c->input(1), 0);
      c->set_output(0, c->MakeShape({b, c->Dim(c->input(0), 0), 1}));  // [B,A,1] as the loader stacks it (preprocess.py:369-370)
      c->set_output(1, c->MakeShape({b, r, 4}));
      c->set_output(2, c->MakeShape({b, r, 4}));
      return tf::Status();
    });

class MrcnnRpnTargetsOp : public tf::OpKernel {
 public:
  explicit MrcnnRpnTargetsOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("rpn_train_anchors_per_image", &r_));
    OP_REQUIRES_OK(c, c->GetAttr("rpn_bbox_std_dev", &std_));
    OP_REQUIRES(c, std_.size() == 4, tf::errors::InvalidArgument("rpn_bbox_std_dev needs 4 values"));
    OP_REQUIRES_OK(c, c->GetAttr("eps", &eps_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& anchors = ctx->input(0);
    const tf::Tensor& cls = ctx->input(1);
    const tf::Tensor& boxes = ctx->input(2);
    const tf::Tensor& keys = ctx->input(3);
    OP_REQUIRES(ctx, anchors.dims() == 2 && anchors.dim_size(1) == 4 && cls.dims() == 2 && boxes.dims() == 3,
                tf::errors::InvalidArgument("expected anchors [A,4], gt_class_ids [B,G], gt_boxes [B,G,4]"));
    const int A = anchors.dim_size(0), B = cls.dim_size(0), G = cls.dim_size(1);
    OP_REQUIRES(ctx, boxes.shape() == tf::TensorShape({B, G, 4}) && keys.shape() == tf::TensorShape({B, A}),
                tf::errors::InvalidArgument("expected gt_boxes [B,G,4] and rand_keys [B,A]"));
    tf::Tensor *match, *bbox, *bbox32;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, A, 1}), &match));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({B, r_, 4}), &bbox));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, tf::TensorShape({B, r_, 4}), &bbox32));
    size_t ws_bytes = 0;
    OP_REQUIRES_OK(ctx, LauncherStatus(mrcnn_rpn_targets_workspace_bytes(B, A, G, r_, &ws_bytes), "rpn targets ws"));
    tf::Tensor ws;
    OP_REQUIRES_OK(ctx, Scratch(ctx, ws_bytes, &ws));
    // the loader passes config['rpn_bbox_std_dev'], a FLOAT32 array (config.py:90, preprocess.py:347), and numpy widens
    // it for `rpn_bbox[ix] /= rpn_bbox_std` (utils.py:259): the divisor is (double)0.1f, not 0.1 -- plain widening here
    double sd[4];
    for (int i = 0; i < 4; ++i) sd[i] = static_cast<double>(std_[i]);
    OP_REQUIRES_OK(ctx, LauncherStatus(
        mrcnn_rpn_targets_forward(anchors.flat<double>().data(), cls.flat<tf::int32>().data(),
                                  boxes.flat<tf::int32>().data(), keys.flat<float>().data(), B, A, G, r_, sd,
                                  AttrAsDouble(eps_), match->flat<tf::int32>().data(),
                                  bbox->flat<double>().data(), bbox32->flat<float>().data(), nullptr,
                                  ws.flat<tf::uint8>().data(), ws_bytes, StreamOf(ctx)),
        "mrcnn_rpn_targets_forward"));
  }

 private:
  int r_;
  float eps_;
  std::vector<float> std_;
};
REGISTER_KERNEL_BUILDER(Name("MrcnnRpnTargets").Device(tf::DEVICE_GPU), MrcnnRpnTargetsOp);

// ---- ProposalLayer fed by the per-level RPN head outputs (rpn_graph per pyramid level, model.py:465-478): the Keras
// softmax and the three Concatenate layers are fused into the launch.  Inference path (no gradient registered).
REGISTER_OP("MrcnnProposalLevels")
    .Input("rpn_class_logits: N * float")   // level l: [B,A_l,2] raw logits (reshaped rpn_class_raw)
    .Input("rpn_bbox: N * float")           // level l: [B,A_l,4] raw deltas (reshaped rpn_bbox_pred)
    .Input("anchors: float")                // [B,A,4], A = sum A_l, level-major
    .Output("proposals: float")             // [B,P,4]
    .Output("rpn_probs: float")             // [B,A,2]: the model's `rpn_class` output (concat_rpn_class)
    .Attr("N: int >= 1")
    .Attr("proposal_count: int")
    .Attr("pre_nms_limit: int = 6000")
    .Attr("nms_threshold: float = 0.7")
    .Attr("std_dev: list(float) = [0.1, 0.1, 0.2, 0.2]")
    .SetShapeFn([](InferenceContext* c) {
      int p, n;
      TF_RETURN_IF_ERROR(c->GetAttr("proposal_count", &p));
      TF_RETURN_IF_ERROR(c->GetAttr("N", &n));
      auto anchors = c->input(2 * n);
      c->set_output(0, c->MakeShape({c->Dim(anchors, 0), p, 4}));
      c->set_output(1, c->MakeShape({c->Dim(anchors, 0), c->Dim(anchors, 1), 2}));
      return tf::Status();
    });

class MrcnnProposalLevelsOp : public tf::OpKernel {
 public:
  explicit MrcnnProposalLevelsOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("N", &n_));
    OP_REQUIRES_OK(c, c->GetAttr("proposal_count", &p_));
    OP_REQUIRES_OK(c, c->GetAttr("pre_nms_limit", &pre_));
    OP_REQUIRES_OK(c, c->GetAttr("nms_threshold", &thr_));
    OP_REQUIRES_OK(c, c->GetAttr("std_dev", &std_));
    OP_REQUIRES(c, std_.size() == 4 && n_ <= 8, tf::errors::InvalidArgument("std_dev needs 4 values, N <= 8"));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& anchors = ctx->input(2 * n_);
    OP_REQUIRES(ctx, anchors.dims() == 3 && anchors.dim_size(2) == 4, tf::errors::InvalidArgument("anchors [B,A,4]"));
    const int B = anchors.dim_size(0), A = anchors.dim_size(1);
    std::vector<const float*> logits(n_), bbox(n_);
    std::vector<int> counts(n_);
    int total = 0;
    for (int l = 0; l < n_; ++l) {
      const tf::Tensor& lg = ctx->input(l);
      const tf::Tensor& bb = ctx->input(n_ + l);
      OP_REQUIRES(ctx, lg.dims() == 3 && lg.dim_size(0) == B && lg.dim_size(2) == 2 &&
                           bb.shape() == tf::TensorShape({B, lg.dim_size(1), 4}),
                  tf::errors::InvalidArgument("level tensors must be [B,A_l,2] and [B,A_l,4]"));
      logits[l] = lg.flat<float>().data();
      bbox[l] = bb.flat<float>().data();
      counts[l] = lg.dim_size(1);
      total += counts[l];
    }
    OP_REQUIRES(ctx, total == A, tf::errors::InvalidArgument("level anchor counts must add up to anchors.shape[1]"));
    tf::Tensor *out = nullptr, *probs = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, p_, 4}), &out));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({B, A, 2}), &probs));
    size_t ws_bytes = 0;
    OP_REQUIRES_OK(ctx, LauncherStatus(mrcnn_proposal_levels_workspace_bytes(B, A, pre_, p_, &ws_bytes), "proposal ws"));
    tf::Tensor ws;
    OP_REQUIRES_OK(ctx, Scratch(ctx, ws_bytes, &ws));
    OP_REQUIRES_OK(ctx, LauncherStatus(
        mrcnn_proposal_forward_levels(logits.data(), bbox.data(), counts.data(), n_, anchors.flat<float>().data(), B,
                                      pre_, p_, std_.data(), thr_, out->flat<float>().data(),
                                      probs->flat<float>().data(), nullptr, nullptr, nullptr,
                                      ws.flat<tf::uint8>().data(), ws_bytes, StreamOf(ctx)),
        "mrcnn_proposal_forward_levels"));
  }

 private:
  int n_, p_, pre_;
  float thr_;
  std::vector<float> std_;
};
REGISTER_KERNEL_BUILDER(Name("MrcnnProposalLevels").Device(tf::DEVICE_GPU), MrcnnProposalLevelsOp);
