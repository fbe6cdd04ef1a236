// Thin pybind11 module over the C ABI (include/fpb.h), built by CMake next to where the
// reference builds lemonpy (flamingpy/cpp/lemonpy.cpp, CMakeLists.txt:53-54).  It follows the
// reference's native pattern (lemonpy.cpp:50-78): C-contiguous NumPy arrays in, freshly owned
// arrays out, failures as C++ exceptions -> Python RuntimeError.  All compute stays in libfpb.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <stdexcept>
#include <string>

#include "../../include/fpb.h"

namespace py = pybind11;

static void check(int rc) {
  if (rc != 0) throw std::runtime_error(std::string("fpb error ") + std::to_string(rc) + ": " + fpb_last_error());
}

PYBIND11_MODULE(fpbpy, m) {
  m.doc() = "pybind11 plugin over the B200 Monte-Carlo EC trial backend (C ABI: include/fpb.h)";
  m.def("abi_version", &fpb_abi_version);
  m.def("device_count", [] {
    int n = fpb_device_count();
    if (n < 0) throw std::runtime_error(fpb_last_error());
    return n;
  });
  // handle lifetime and runs take the address of an fpb_config / fpb_handle built by the caller
  // (flamingpy_b200._lib builds the structs with ctypes), so the module stays a pure marshaller
  m.def("create", [](std::uintptr_t cfg_addr) {
    fpb_handle* h = nullptr;
    check(fpb_create(reinterpret_cast<const fpb_config*>(cfg_addr), &h));
    return reinterpret_cast<std::uintptr_t>(h);
  });
  m.def("destroy", [](std::uintptr_t h) { check(fpb_destroy(reinterpret_cast<fpb_handle*>(h))); });
  m.def("run_injected",
        [](std::uintptr_t h, py::array_t<double, py::array::c_style | py::array::forcecast> x,
           py::array_t<uint8_t, py::array::c_style | py::array::forcecast> state, int stage) {
          if (x.ndim() != 2 || state.ndim() != 2 || x.shape(0) != state.shape(0))
            throw std::runtime_error("x must be [T, 2N] and state [T, N]");
          py::gil_scoped_release release;
          check(fpb_run_injected(reinterpret_cast<fpb_handle*>(h), x.shape(0), x.data(), state.data(), 0, stage));
        });
  m.def("run_rng", [](std::uintptr_t h, uint64_t seed, int64_t first_trial, int64_t n_trials, int stage) {
    py::gil_scoped_release release;
    check(fpb_run_rng(reinterpret_cast<fpb_handle*>(h), seed, first_trial, n_trials, stage));
  });
  m.def("run_iid", [](std::uintptr_t h, uint64_t seed, int64_t first_trial, int64_t n_trials, double p_err, int stage) {
    py::gil_scoped_release release;
    check(fpb_run_iid(reinterpret_cast<fpb_handle*>(h), seed, first_trial, n_trials, p_err, stage));
  });
  m.def("run_bits",
        [](std::uintptr_t h, py::array_t<uint32_t, py::array::c_style | py::array::forcecast> bits, int stage) {
          if (bits.ndim() != 2) throw std::runtime_error("bits must be uint32[n_trials][bit_words]");
          py::gil_scoped_release release;
          check(fpb_run_bits(reinterpret_cast<fpb_handle*>(h), bits.shape(0), bits.data(), 0, stage));
        });
  m.def("sizes", [](std::uintptr_t h) {
    fpb_sizes s;
    check(fpb_get_sizes(reinterpret_cast<fpb_handle*>(h), &s));
    py::dict d;
    d["n_trials"] = s.n_trials;
    d["n_complex"] = s.n_complex;
    d["total_qubits"] = s.total_qubits;
    d["total_odd"] = py::make_tuple(s.total_odd[0], s.total_odd[1]);
    d["total_pairs"] = py::make_tuple(s.total_pairs[0], s.total_pairs[1]);
    return d;
  });
  m.def("pair_lengths", [](std::uintptr_t h, int complex_index) {
    fpb_sizes s;
    check(fpb_get_sizes(reinterpret_cast<fpb_handle*>(h), &s));
    if (complex_index < 0 || complex_index >= s.n_complex) throw std::runtime_error("bad complex index");
    py::array_t<double> out(s.total_pairs[complex_index]);
    fpb_outputs o = {};
    o.pair_len[complex_index] = out.mutable_data();
    check(fpb_fetch(reinterpret_cast<fpb_handle*>(h), &o, 0));
    return out;
  });
}
