// TEST INFRASTRUCTURE ONLY (oracle/): binding shim for the UNMODIFIED reference sources.
//
// Builds the reference implementation of the hot path straight from the files where they lie
// under /root/reference/engine/kernels (never copied into this repo) so that tests/ and
// bench.py's `--impl reference` / cpu_baseline leg can call the reference's own code:
//   tpack / tunpack            -> /root/reference/engine/kernels/tpack/tpack.cu:203-255, :429-476
//   quantconv2d_float_input    -> /root/reference/engine/kernels/functions/quantconv2d_float_input.cu:140-220
//   quantlinear_float_input    -> /root/reference/engine/kernels/functions/quantlinear_float_input.cu:120-182
//   quantconv2d                -> /root/reference/engine/kernels/functions/quantconv2d.cu:164-264
//   quantlinear                -> /root/reference/engine/kernels/functions/quantlinear.cu:231-297
// The reference's own pybind.cpp (engine/kernels/pybind.cpp:7-17) registers all 8 ops and would pull
// in the off-path .cu files; this shim registers only the ops this repository rebuilds under a different
// module name so it can be imported next to the product's `quant_engine`.
#include <pybind11/pybind11.h>
#include <torch/extension.h>
#include "tpack/tpack.h"
#include "functions/funcs.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
    m.def("tpack", &tpack, "reference tpack");
    m.def("tunpack", &tunpack, "reference tunpack");
    m.def("quantconv2d_float_input", &quantconv2d_float_input, "reference quantconv2d_float_input");
    m.def("quantlinear_float_input", &quantlinear_float_input, "reference quantlinear_float_input");
    m.def("quantconv2d", &quantconv2d, "reference quantconv2d");
    m.def("quantlinear", &quantlinear, "reference quantlinear");
}
