"""The slice of the TensorFlow Lite FlatBuffer schema (`schema.fbs`, file identifier TFL3) that
EfficientDet-Lite detection models use.  The reference loads such files through
tflite_runtime.Interpreter(model_path=...) (track.py:68,93; eval.py:167); its own six
`models/*.tflite` blobs are absent from the checkout (.MISSING_LARGE_BLOBS), and neither
tflite_runtime nor the `flatbuffers` module is installed here, so field ids and enum values are
restated from the public schema [3P-MEM] -- `tests/test_tflite_io.py` can only prove the reader
and the writer against each other, not against a file written by TensorFlow.

Field ids (position in the table declaration):
  Model:        0 version, 1 operator_codes, 2 subgraphs, 3 description, 4 buffers
  OperatorCode: 0 deprecated_builtin_code (i8), 1 custom_code, 2 version, 3 builtin_code (i32)
  SubGraph:     0 tensors, 1 inputs, 2 outputs, 3 operators, 4 name
  Tensor:       0 shape, 1 type, 2 buffer, 3 name, 4 quantization
  Quantization: 0 min, 1 max, 2 scale, 3 zero_point, 4 details_type, 5 details, 6 quantized_dimension
  Operator:     0 opcode_index, 1 inputs, 2 outputs, 3 builtin_options_type, 4 builtin_options,
                5 custom_options, 6 custom_options_format
  Buffer:       0 data
"""

# TensorType
FLOAT32, INT32, UINT8, INT64, INT8 = 0, 2, 3, 4, 9
NP_OF_TYPE = {FLOAT32: 'f32', INT32: 'i32', UINT8: 'u8', INT64: 'i64', INT8: 'i8'}

# BuiltinOperator
ADD, CONCATENATION, CONV_2D, DEPTHWISE_CONV_2D, DEQUANTIZE = 0, 2, 3, 4, 6
LOGISTIC, MAX_POOL_2D, RESHAPE, CUSTOM, RESIZE_NEAREST_NEIGHBOR, QUANTIZE = 14, 17, 22, 32, 97, 114
OP_NAMES = {ADD: 'ADD', CONCATENATION: 'CONCATENATION', CONV_2D: 'CONV_2D',
            DEPTHWISE_CONV_2D: 'DEPTHWISE_CONV_2D', DEQUANTIZE: 'DEQUANTIZE', LOGISTIC: 'LOGISTIC',
            MAX_POOL_2D: 'MAX_POOL_2D', RESHAPE: 'RESHAPE', CUSTOM: 'CUSTOM',
            RESIZE_NEAREST_NEIGHBOR: 'RESIZE_NEAREST_NEIGHBOR', QUANTIZE: 'QUANTIZE'}

# BuiltinOptions union tags
OPT_NONE, OPT_CONV2D, OPT_DEPTHWISE, OPT_POOL2D, OPT_CONCAT, OPT_ADD, OPT_RESHAPE = 0, 1, 2, 5, 10, 11, 17
# QUANTIZE / DEQUANTIZE / LOGISTIC / RESIZE_NEAREST_NEIGHBOR are written without an options table (all
# their options default to 0 / false, which is what these graphs use) and read without one.
#   Conv2DOptions:          0 padding, 1 stride_w, 2 stride_h, 3 fused_activation_function
#   DepthwiseConv2DOptions: 0 padding, 1 stride_w, 2 stride_h, 3 depth_multiplier, 4 fused_activation_function
#   Pool2DOptions:          0 padding, 1 stride_w, 2 stride_h, 3 filter_width, 4 filter_height, 5 fused_activation_function
#   AddOptions:             0 fused_activation_function
#   ConcatenationOptions:   0 axis, 1 fused_activation_function
#   ReshapeOptions:         0 new_shape
#   ResizeNearestNeighborOptions: 0 align_corners, 1 half_pixel_centers

PAD_SAME, PAD_VALID = 0, 1
ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 3

DETECTION_POSTPROCESS = 'TFLite_Detection_PostProcess'
