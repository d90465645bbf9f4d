// TEST INFRASTRUCTURE ONLY (oracle/): minimal pybind shim that exposes the
// reference's `altcorr_forward` (declared at /root/reference/src/droid.cpp:193-203,
// defined in /root/reference/src/altcorr_kernel.cu:290-319) without the rest of
// `droid_backends` (which needs Eigen + lietorch and cannot be built here).
// The .cu file is compiled from where it lies under /root/reference; only this
// 20-line shim is ours.
#include <torch/extension.h>
#include <vector>

std::vector<torch::Tensor> altcorr_cuda_forward(
    torch::Tensor fmap1, torch::Tensor fmap2, torch::Tensor coords, int radius);

static std::vector<torch::Tensor> altcorr_forward(
    torch::Tensor fmap1, torch::Tensor fmap2, torch::Tensor coords, int radius) {
  TORCH_CHECK(fmap1.is_contiguous() && fmap2.is_contiguous() && coords.is_contiguous(),
              "altcorr_forward: inputs must be contiguous");
  return altcorr_cuda_forward(fmap1, fmap2, coords, radius);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("altcorr_forward", &altcorr_forward, "reference altcorr forward");
}
