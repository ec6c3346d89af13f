// ssr_tf_ops.cc — TensorFlow custom-op shim over the C ABI of libssr_b200.so (include/ssr_b200.h).
//
// BASELINE.json north_star: "Python host code calling hand-written sm_100a CUDA through a thin C-ABI loaded as TensorFlow
// custom ops via tf.load_op_library".  The reference is TF 2.2 (requirements.txt:65); TensorFlow cannot be installed in
// the build image (no network), so this file is compiled ONLY where `import tensorflow` works - tf_shim/build.py does
// that and skips loudly otherwise.  UNTESTED HERE: it has never been compiled against TensorFlow headers.
//
// Each op forwards TF-allocated device buffers and TF's compute stream to one ABI call: the library never allocates,
// never synchronises, never copies (the ABI's contract), so the ops are ordinary asynchronous GPU OpKernels.
//   SsrConv2d            Conv2D(padding="same", strides=1) + BiasAdd + activation (+ residual)   model_builder.py:285-293
//   SsrConv2dGradInput   Conv2DBackpropInput  (the same kernel over dZ with the rotated weight image)
//   SsrConv2dGradFilter  Conv2DBackpropFilter + BiasAddGrad (split-K tcgen05)
//   SsrPackWeights       HWIO fp32 kernel -> UMMA operand image (forward or dgrad form)
//   SsrDepthToSpace2     tf.nn.depth_to_space(x, 2)                                              model_builder.py:279
// Gradients are registered on the Python side (tf_shim/ssr_tf.py) so that tape.gradient(loss, model.trainable_variables)
// (sr_model.py:419-447) traverses a model built from these ops.
#include <cstring>

#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#define EIGEN_USE_GPU
#include "tensorflow/core/util/gpu_kernel_helper.h"

#include "../include/ssr_b200.h"

namespace tf = tensorflow;
using tf::shape_inference::InferenceContext;

namespace {

ssr_ctx* Ctx() {   // one context per process and device 0..7, created on first use
  static ssr_ctx* ctx[8] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev > 7) return nullptr;
  if (ctx[dev] == nullptr) ssr_ctx_create(dev, &ctx[dev]);
  return ctx[dev];
}

void* Stream(tf::OpKernelContext* c) { return reinterpret_cast<void*>(c->eigen_gpu_device().stream()); }

#define SSR_REQUIRE(c, rc)                                                                         \
  OP_REQUIRES(c, (rc) == SSR_OK,                                                                   \
              (rc) == SSR_ERR_INVALID ? tf::errors::InvalidArgument(ssr_last_error())              \
                                      : tf::errors::Internal(ssr_last_error()))

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
REGISTER_OP("SsrPackWeights")
    .Input("kernel: float")          // HWIO [kh, kw, cin, cout]
    .Output("packed: uint8")
    .Attr("cin_padded: int")
    .Attr("up: int = 1")
    .Attr("dgrad: bool = false")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->Vector(c->UnknownDim()));
      return tf::Status::OK();
    });

class SsrPackWeightsOp : public tf::OpKernel {
 public:
  explicit SsrPackWeightsOp(tf::OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("cin_padded", &cin_));
    OP_REQUIRES_OK(c, c->GetAttr("up", &up_));
    OP_REQUIRES_OK(c, c->GetAttr("dgrad", &dgrad_));
  }
  void Compute(tf::OpKernelContext* c) override {
    const tf::Tensor& k = c->input(0);
    OP_REQUIRES(c, k.dims() == 4, tf::errors::InvalidArgument("kernel must be HWIO"));
    const int kh = k.dim_size(0), kw = k.dim_size(1), cin = k.dim_size(2), cout = k.dim_size(3);
    size_t bytes = dgrad_ ? ssr_conv2d_packed_bytes_hw(kh, kw, (cout + 15) / 16 * 16, cin, 1)
                          : ssr_conv2d_packed_bytes_hw(kh, kw, cin_, cout, up_);
    OP_REQUIRES(c, bytes > 0, tf::errors::InvalidArgument(ssr_last_error()));
    tf::Tensor* out = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, tf::TensorShape({static_cast<tf::int64>(bytes)}), &out));
    int rc = dgrad_ ? ssr_conv2d_pack_weights_dgrad(Ctx(), k.flat<float>().data(), kh, kw, cin, cout, 0,
                                                    out->flat<tf::uint8>().data(), Stream(c))
                    : ssr_conv2d_pack_weights_hw(Ctx(), k.flat<float>().data(), kh, kw, cin, cin_, cout, up_,
                                                 out->flat<tf::uint8>().data(), Stream(c));
    SSR_REQUIRE(c, rc);
  }

 private:
  int cin_, up_;
  bool dgrad_;
};
REGISTER_KERNEL_BUILDER(Name("SsrPackWeights").Device(tf::DEVICE_GPU), SsrPackWeightsOp);

// ---------------------------------------------------------------------------------------------------------------------
REGISTER_OP("SsrConv2d")
    .Input("x: bfloat16")            // NHWC, channels = in_cstride
    .Input("packed: uint8")
    .Input("bias: float")
    .Input("alpha: float")           // PReLU slopes ([0] when unused)
    .Input("res: bfloat16")          // residual, or a [0] tensor
    .Output("y: bfloat16")
    .Attr("cin: int")
    .Attr("cout: int")
    .Attr("ksize: int = 3")
    .Attr("act: int = 0")            // ssr_act
    .Attr("act_alpha: float = 0.2")
    .Attr("res_beta: float = 1.0")
    .Attr("up: int = 1")
    .SetShapeFn([](InferenceContext* c) {
      int cout, up;
      TF_RETURN_IF_ERROR(c->GetAttr("cout", &cout));
      TF_RETURN_IF_ERROR(c->GetAttr("up", &up));
      tf::shape_inference::ShapeHandle x = c->input(0);
      tf::shape_inference::DimensionHandle h, w;
      TF_RETURN_IF_ERROR(c->Multiply(c->Dim(x, 1), up, &h));
      TF_RETURN_IF_ERROR(c->Multiply(c->Dim(x, 2), up, &w));
      c->set_output(0, c->MakeShape({c->Dim(x, 0), h, w, c->MakeDim(up == 2 ? cout / 4 : cout)}));
      return tf::Status::OK();
    });

class SsrConv2dOp : public tf::OpKernel {
 public:
  explicit SsrConv2dOp(tf::OpKernelConstruction* c) : OpKernel(c) {
    memset(&d_, 0, sizeof(d_));
    int v;
    float f;
    OP_REQUIRES_OK(c, c->GetAttr("cin", &v)); d_.cin = v;
    OP_REQUIRES_OK(c, c->GetAttr("cout", &v)); d_.cout = v;
    OP_REQUIRES_OK(c, c->GetAttr("ksize", &v)); d_.ksize = v;
    OP_REQUIRES_OK(c, c->GetAttr("act", &v)); d_.act = v;
    OP_REQUIRES_OK(c, c->GetAttr("act_alpha", &f)); d_.act_alpha = f;
    OP_REQUIRES_OK(c, c->GetAttr("res_beta", &f)); d_.res_beta = f;
    OP_REQUIRES_OK(c, c->GetAttr("up", &v)); d_.up = v;
  }
  void Compute(tf::OpKernelContext* c) override {
    const tf::Tensor& x = c->input(0);
    OP_REQUIRES(c, x.dims() == 4, tf::errors::InvalidArgument("x must be NHWC"));
    ssr_conv_desc d = d_;
    d.n = x.dim_size(0); d.h = x.dim_size(1); d.w = x.dim_size(2);
    d.in_cstride = x.dim_size(3);
    d.in_cvalid = d.in_cstride;
    const int cs = d.up == 2 ? d.cout / 4 : d.cout;
    d.out_dtype = SSR_BF16; d.out_cstride = cs; d.out_coff = 0;
    const bool has_res = c->input(4).NumElements() > 0;
    d.res_dtype = has_res ? SSR_BF16 : SSR_NONE; d.res_cstride = cs; d.res_coff = 0;
    tf::Tensor* y = nullptr;                                               // TF owns the memory
    OP_REQUIRES_OK(c, c->allocate_output(0, tf::TensorShape({d.n, d.h * d.up, d.w * d.up, cs}), &y));
    const float* alpha = c->input(3).NumElements() > 0 ? c->input(3).flat<float>().data() : nullptr;
    int rc = ssr_conv2d_fwd(Ctx(), &d, x.tensor_data().data(), c->input(1).flat<tf::uint8>().data(),
                            c->input(2).flat<float>().data(), alpha,
                            has_res ? c->input(4).tensor_data().data() : nullptr,
                            const_cast<char*>(y->tensor_data().data()), nullptr, Stream(c));
    SSR_REQUIRE(c, rc);                                                    // SSR_ERR_INVALID <-> ValueError
  }

 private:
  ssr_conv_desc d_;
};
REGISTER_KERNEL_BUILDER(Name("SsrConv2d").Device(tf::DEVICE_GPU), SsrConv2dOp);

// ---------------------------------------------------------------------------------------------------------------------
REGISTER_OP("SsrConv2dGradFilter")
    .Input("x: bfloat16")
    .Input("dz: bfloat16")
    .Output("dkernel: float")        // HWIO
    .Output("dbias: float")
    .Attr("cin: int")
    .Attr("cout: int")
    .Attr("ksize: int = 3")
    .SetShapeFn([](InferenceContext* c) {
      int cin, cout, k;
      TF_RETURN_IF_ERROR(c->GetAttr("cin", &cin));
      TF_RETURN_IF_ERROR(c->GetAttr("cout", &cout));
      TF_RETURN_IF_ERROR(c->GetAttr("ksize", &k));
      c->set_output(0, c->MakeShape({k, k, cin, cout}));
      c->set_output(1, c->Vector(cout));
      return tf::Status::OK();
    });

class SsrConv2dGradFilterOp : public tf::OpKernel {
 public:
  explicit SsrConv2dGradFilterOp(tf::OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("cin", &cin_));
    OP_REQUIRES_OK(c, c->GetAttr("cout", &cout_));
    OP_REQUIRES_OK(c, c->GetAttr("ksize", &k_));
  }
  void Compute(tf::OpKernelContext* c) override {
    const tf::Tensor &x = c->input(0), &dz = c->input(1);
    const int n = x.dim_size(0), h = x.dim_size(1), w = x.dim_size(2);
    tf::Tensor *dk = nullptr, *db = nullptr, ws;
    OP_REQUIRES_OK(c, c->allocate_output(0, tf::TensorShape({k_, k_, cin_, cout_}), &dk));
    OP_REQUIRES_OK(c, c->allocate_output(1, tf::TensorShape({cout_}), &db));
    const size_t wsb = ssr_conv2d_wgrad_workspace_bytes(Ctx(), h, w, cin_, cout_, k_, k_);
    OP_REQUIRES(c, wsb > 0, tf::errors::InvalidArgument(ssr_last_error()));
    OP_REQUIRES_OK(c, c->allocate_temp(tf::DT_UINT8, tf::TensorShape({static_cast<tf::int64>(wsb)}), &ws));  // scratch from TF
    int rc = ssr_conv2d_wgrad_bias(Ctx(), x.tensor_data().data(), x.dim_size(3), 0, cin_, dz.tensor_data().data(),
                                   dz.dim_size(3), 0, cout_, n, h, w, k_, k_, 1.0f, 0, ws.flat<tf::uint8>().data(),
                                   dk->flat<float>().data(), db->flat<float>().data(), 1.0f, 0, Stream(c));
    SSR_REQUIRE(c, rc);
  }

 private:
  int cin_, cout_, k_;
};
REGISTER_KERNEL_BUILDER(Name("SsrConv2dGradFilter").Device(tf::DEVICE_GPU), SsrConv2dGradFilterOp);

// SsrConv2dGradInput needs no kernel of its own: dX = SsrConv2d(dZ, SsrPackWeights(kernel, dgrad=true), bias=0) - see
// ssr_tf.py::_conv_grad.

// ---------------------------------------------------------------------------------------------------------------------
REGISTER_OP("SsrDepthToSpace2")
    .Input("x: T")
    .Output("y: T")
    .Attr("T: {bfloat16, float}")
    .SetShapeFn([](InferenceContext* c) {
      tf::shape_inference::ShapeHandle x = c->input(0);
      tf::shape_inference::DimensionHandle h, w, ch;
      TF_RETURN_IF_ERROR(c->Multiply(c->Dim(x, 1), 2, &h));
      TF_RETURN_IF_ERROR(c->Multiply(c->Dim(x, 2), 2, &w));
      TF_RETURN_IF_ERROR(c->Divide(c->Dim(x, 3), 4, true, &ch));
      c->set_output(0, c->MakeShape({c->Dim(x, 0), h, w, ch}));
      return tf::Status::OK();
    });

template <typename T>
class SsrDepthToSpace2Op : public tf::OpKernel {
 public:
  explicit SsrDepthToSpace2Op(tf::OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(tf::OpKernelContext* c) override {
    const tf::Tensor& x = c->input(0);
    const int n = x.dim_size(0), h = x.dim_size(1), w = x.dim_size(2), ch = x.dim_size(3) / 4;
    tf::Tensor* y = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, tf::TensorShape({n, 2 * h, 2 * w, ch}), &y));
    SSR_REQUIRE(c, ssr_depth_to_space2(x.tensor_data().data(), const_cast<char*>(y->tensor_data().data()), n, h, w, ch,
                                       sizeof(T), Stream(c)));
  }
};
REGISTER_KERNEL_BUILDER(Name("SsrDepthToSpace2").Device(tf::DEVICE_GPU).TypeConstraint<tf::bfloat16>("T"),
                        SsrDepthToSpace2Op<tf::bfloat16>);
REGISTER_KERNEL_BUILDER(Name("SsrDepthToSpace2").Device(tf::DEVICE_GPU).TypeConstraint<float>("T"),
                        SsrDepthToSpace2Op<float>);
