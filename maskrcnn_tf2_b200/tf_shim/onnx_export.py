"""tf2onnx side of the drop-in: keyword arguments that let the reference's export call sites convert a model
built from the B200 layers.

The reference exports with `maskrcnn_to_onnx(model, model_name, input_spec, kwargs)`, which forwards `kwargs` to
`tf2onnx.convert.from_keras` (src/common/inference_optimize.py:12-20), then `modify_onnx_model` deletes everything
whose tensor name contains the layer names `mrcnn_detection`, `mask_rcnn_inference/roi/`, `roi_align_classifier/`,
`roi_align_mask/` and wires NVIDIA's TensorRT plugins in their place (inference_optimize.py:455-465, 470-560).  With
the B200 layers each of those Keras layers is ONE TF node of a custom op type (tf_shim/mrcnn_roi_ops.cc), named
`<model>/<layer name>/<OpType>`, so the name filter still finds them; the only thing tf2onnx needs is to be told that
the op types are known.  Usage, in the export notebook:

    from onnx_export import tf2onnx_kwargs
    maskrcnn_to_onnx(model, model_name, input_spec, kwargs=tf2onnx_kwargs({'opset': 11}))
    modify_onnx_model(model_path, config)            # unchanged

This module imports neither TensorFlow nor tf2onnx (the handlers only touch the node objects tf2onnx passes in), so
the CPU test suite of this repository checks the table below against the op registrations in mrcnn_roi_ops.cc.
"""

#: ONNX domain the custom nodes are emitted in (an ONNX runtime without the ops can still load the model)
DOMAIN = "ai.mrcnn_roi_b200"

#: TF op type -> (number of layer-result outputs, attributes the node carries).  Inference export only needs the first three;
#: the training-side ops are listed so that a training graph converts too.  Output 0 is always the layer's result;
#: outputs after those are saved for the gradients (top-k / keep indices, roi_map) and must have no consumer.
EXPORTED_OPS = {
    "MrcnnProposal": (1, ("proposal_count", "pre_nms_limit", "nms_threshold", "std_dev")),
    "MrcnnProposalLevels": (2, ("N", "proposal_count", "pre_nms_limit", "nms_threshold", "std_dev")),
    "MrcnnPyramidRoiAlign": (1, ("pool_height", "pool_width", "denominator", "map_mode")),
    "MrcnnDetection": (2, ("min_confidence", "use_min_confidence", "max_instances", "nms_threshold", "std_dev")),
    "MrcnnDetectionTarget": (4, ("train_rois_per_image", "roi_positive_ratio", "mask_height", "mask_width",
                                 "use_mini_masks", "std_dev")),
}


def _handler(ctx, node, name, args):
    """tf2onnx custom-op handler: keep the node as it is (type, attributes, all outputs) and move it into DOMAIN.
    The outputs saved for the gradients stay on the node; in an inference graph nothing consumes them."""
    n_out, _attrs = EXPORTED_OPS[node.type]
    for out in node.output[n_out:]:
        if ctx.find_output_consumers(out):
            raise ValueError(f"{node.type} output {out} is consumed in the exported graph; export the inference model "
                             f"(maskrcnn_to_onnx accepts only 'mask_rcnn_inference', inference_optimize.py:13-14)")
    node.domain = DOMAIN
    return node


def tf2onnx_kwargs(base=None):
    """`kwargs` for maskrcnn_to_onnx: the caller's own entries (opset, ...) plus the custom-op tables."""
    kw = dict(base or {})
    kw["custom_ops"] = {**{op: DOMAIN for op in EXPORTED_OPS}, **kw.get("custom_ops", {})}
    kw["custom_op_handlers"] = {**{op: (_handler, []) for op in EXPORTED_OPS}, **kw.get("custom_op_handlers", {})}
    return kw
