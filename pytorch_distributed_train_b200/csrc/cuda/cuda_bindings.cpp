// CUDA-side bindings (symmetric memory, NVLink collectives, NCCL baseline, sm_100a ops).
#include <torch/extension.h>

namespace py = pybind11;

namespace pdt {
void register_cuda_bindings(py::module_& m) {
  (void)m;
}
}  // namespace pdt
