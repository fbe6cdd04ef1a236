// pybind11 module over the C ABI (include/fpb.h), built next to where the reference builds
// lemonpy (flamingpy/cpp/lemonpy.cpp, CMakeLists.txt:53-54; here: CMakeLists.txt and
// __graft_entry__.build()).  It follows the reference's native pattern (lemonpy.cpp:50-78):
// C-contiguous NumPy arrays in, NumPy arrays out, failures as C++ exceptions -> Python
// RuntimeError, GIL released around device work.  All compute stays in libfpb; this file only
// marshals.  flamingpy_b200/_lib.py routes every TrialEngine call through `Handle` when the module
// is importable (ctypes otherwise), so it is on the product path and in the -m gpu tests.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fpb.h"

namespace py = pybind11;

namespace {

// Failures surface as flamingpy_b200._lib.FpbError (a RuntimeError, the analogue of lemonpy.cpp:53-54)
// once _lib has registered the class; plain RuntimeError before that.  Must be called with the GIL held.
PyObject* g_error_class = nullptr;
void check(int rc) {
  if (rc == 0) return;
  const std::string msg = std::string("fpb error ") + std::to_string(rc) + ": " + fpb_last_error();
  if (g_error_class) {
    PyErr_SetString(g_error_class, msg.c_str());
    throw py::error_already_set();
  }
  throw std::runtime_error(msg);
}

template <typename T>
using carray = py::array_t<T, py::array::c_style | py::array::forcecast>;

// keeps the converted (contiguous, right dtype) arrays of a config alive until fpb_create returns
struct ConfigHolder {
  fpb_config cfg;
  std::vector<py::object> keep;
  template <typename T>
  const T* arr(const py::handle& obj, size_t min_len, const char* what) {
    carray<T> a = carray<T>::ensure(obj);
    if (!a) throw std::runtime_error(std::string("config: ") + what + " is not convertible to the expected dtype");
    if ((size_t)a.size() < min_len) throw std::runtime_error(std::string("config: ") + what + " is too short");
    keep.push_back(a);
    return a.size() ? a.data() : nullptr;
  }
};

template <typename T>
T get(const py::dict& d, const char* key) {
  if (!d.contains(key)) throw std::runtime_error(std::string("config: missing key '") + key + "'");
  return d[key].cast<T>();
}
template <typename T>
T get_or(const py::dict& d, const char* key, T dflt) {
  return d.contains(key) ? d[key].cast<T>() : dflt;
}

// writable, C-contiguous destination of `count` elements of T, or nullptr when the key is absent
template <typename T>
T* dest(const py::dict& bufs, const std::string& key, size_t count) {
  if (!bufs.contains(key.c_str())) return nullptr;
  py::object o = bufs[key.c_str()];
  if (o.is_none()) return nullptr;
  if (!py::isinstance<py::array>(o)) throw std::runtime_error("buffer '" + key + "' is not a NumPy array");
  py::array a = py::reinterpret_borrow<py::array>(o);
  if (!(a.flags() & py::array::c_style) || !a.writeable())
    throw std::runtime_error("buffer '" + key + "' must be C-contiguous and writable");
  if ((size_t)a.itemsize() != sizeof(T)) throw std::runtime_error("buffer '" + key + "' has the wrong item size");
  if ((size_t)a.size() < count) throw std::runtime_error("buffer '" + key + "' too small");
  return static_cast<T*>(a.mutable_data());
}

class Handle {
 public:
  // config: the fields of fpb_config by name; "complexes" is a list of dicts with the fields of
  // fpb_complex_desc (arrays as NumPy arrays or sequences)
  explicit Handle(const py::dict& c) {
    ConfigHolder H;
    fpb_config& cfg = H.cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = FPB_ABI_VERSION;
    cfg.device = get_or<int>(c, "device", 0);
    cfg.n_nodes = get<int>(c, "n_nodes");
    if (cfg.n_nodes <= 0) throw std::runtime_error("config: n_nodes must be positive");
    cfg.adj_indptr = H.arr<int32_t>(c["adj_indptr"], (size_t)cfg.n_nodes + 1, "adj_indptr");
    const size_t nnz = (size_t)cfg.adj_indptr[cfg.n_nodes];
    cfg.adj_indices = H.arr<int32_t>(c["adj_indices"], nnz, "adj_indices");
    cfg.adj_vals = H.arr<int8_t>(c["adj_vals"], nnz, "adj_vals");
    py::list cxs = get<py::list>(c, "complexes");
    cfg.n_complex = (int)cxs.size();
    if (cfg.n_complex < 1 || cfg.n_complex > FPB_MAX_COMPLEX) throw std::runtime_error("config: 1 or 2 complexes");
    for (int i = 0; i < cfg.n_complex; ++i) {
      py::dict d = cxs[i].cast<py::dict>();
      fpb_complex_desc& x = cfg.complex_[i];
      x.n_cubes = get<int>(d, "n_cubes");
      x.n_qubits = get<int>(d, "n_qubits");
      x.n_slots = get<int>(d, "n_slots");
      if (x.n_cubes <= 0 || x.n_qubits <= 0 || x.n_slots <= 0) throw std::runtime_error("config: empty complex");
      x.qubit_node = H.arr<int32_t>(d["qubit_node"], (size_t)x.n_qubits, "qubit_node");
      x.cube_qubit = H.arr<int32_t>(d["cube_qubit"], (size_t)x.n_cubes * 6, "cube_qubit");
      x.ninfo = H.arr<uint16_t>(d["ninfo"], (size_t)x.n_cubes, "ninfo");
      x.sbase = H.arr<int32_t>(d["sbase"], (size_t)x.n_cubes, "sbase");
      const int32_t* doff = H.arr<int32_t>(d["d_off"], 12, "d_off");
      const int32_t* soff = H.arr<int32_t>(d["s_off"], 12, "s_off");
      std::memcpy(x.d_off, doff, sizeof x.d_off);
      std::memcpy(x.s_off, soff, sizeof x.s_off);
      x.slot_qubit = H.arr<int32_t>(d["slot_qubit"], (size_t)x.n_slots, "slot_qubit");
      x.n_low = (int)py::len(d["low_cube"]);
      x.n_high = (int)py::len(d["high_cube"]);
      x.low_cube = H.arr<int32_t>(d["low_cube"], (size_t)x.n_low, "low_cube");
      x.low_slot = H.arr<int32_t>(d["low_slot"], (size_t)x.n_low, "low_slot");
      x.high_cube = H.arr<int32_t>(d["high_cube"], (size_t)x.n_high, "high_cube");
      x.high_slot = H.arr<int32_t>(d["high_slot"], (size_t)x.n_high, "high_slot");
    }
    cfg.delta = get<double>(c, "delta");
    cfg.p_swap = get_or<double>(c, "p_swap", 0.0);
    cfg.weight_method = get_or<int>(c, "weight_method", FPB_WEIGHT_UNIFORM);
    cfg.weight_delta = get_or<double>(c, "weight_delta", cfg.delta);
    cfg.weight_integer = get_or<int>(c, "weight_integer", 0);
    cfg.weight_multiplier = get_or<double>(c, "weight_multiplier", 1.0);
    cfg.label_mode = get_or<int>(c, "label_mode", FPB_MODE_COMPAT);
    cfg.chain_len = get_or<int64_t>(c, "chain_len", 0);
    cfg.delta_step = get_or<double>(c, "delta_step", 0.0);
    cfg.length_mode = get_or<int>(c, "length_mode", FPB_LENGTH_FLOAT);
    cfg.reserve_sms = get_or<int>(c, "reserve_sms", 0);
    check(fpb_create(&cfg, &h_));
    n_nodes_ = cfg.n_nodes;
    ncx_ = cfg.n_complex;
    qtot_ = 0;
    for (int i = 0; i < ncx_; ++i) {
      qtot_ += cfg.complex_[i].n_qubits;
      cubes_[i] = cfg.complex_[i].n_cubes;
      has_bnd_[i] = cfg.complex_[i].n_low > 0;
    }
    bit_words_ = (qtot_ + 31) / 32;
  }
  ~Handle() { close(); }
  Handle(const Handle&) = delete;
  Handle& operator=(const Handle&) = delete;

  void close() {
    if (h_) {
      fpb_destroy(h_);
      h_ = nullptr;
    }
  }
  fpb_handle* h() const {
    if (!h_) throw std::runtime_error("fpb handle is closed");
    return h_;
  }
  std::uintptr_t address() const { return reinterpret_cast<std::uintptr_t>(h()); }

  // ---- runs (include/fpb.h: fpb_run_injected / fpb_run_rng / fpb_run_bits / fpb_run_iid / fpb_run_macro) ----
  void run_injected(carray<double> x, carray<uint8_t> state, int stage) {
    if (x.ndim() != 2 || state.ndim() != 2 || x.shape(0) != state.shape(0) || x.shape(1) != 2 * (py::ssize_t)n_nodes_ ||
        state.shape(1) != (py::ssize_t)n_nodes_)
      throw std::runtime_error("x must be float64[T, 2N] and state uint8[T, N]");
    const int64_t T = x.shape(0);
    const double* xp = x.data();
    const uint8_t* sp = state.data();
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_run_injected(h(), T, xp, sp, 0, stage);
    }
    check(rc);
  }
  void run_rng(uint64_t seed, int64_t first_trial, int64_t n_trials, int stage) {
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_run_rng(h(), seed, first_trial, n_trials, stage);
    }
    check(rc);
  }
  void run_bits(carray<uint32_t> bits, int stage) {
    if (bits.ndim() != 2 || bits.shape(1) != (py::ssize_t)bit_words_)
      throw std::runtime_error("bits must be uint32[n_trials][bit_words]");
    const int64_t T = bits.shape(0);
    const uint32_t* bp = bits.data();
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_run_bits(h(), T, bp, 0, stage);
    }
    check(rc);
  }
  void run_weighted(carray<double> weights, carray<uint32_t> bits, int stage) {
    if (weights.ndim() != 2 || bits.ndim() != 2 || weights.shape(0) != bits.shape(0) ||
        weights.shape(1) != (py::ssize_t)qtot_ || bits.shape(1) != (py::ssize_t)bit_words_)
      throw std::runtime_error("weights must be float64[T, Qtot] and bits uint32[T, bit_words]");
    const int64_t T = weights.shape(0);
    const double* wp = weights.data();
    const uint32_t* bp = bits.data();
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_run_weighted(h(), T, wp, bp, 0, stage);
    }
    check(rc);
  }
  // ---- macronode layer (fpb_set_macro / fpb_run_macro_* / fpb_fetch_macro) ----
  void set_macro(carray<int32_t> partner, carray<int8_t> sign) {
    if (partner.size() != 4 * (py::ssize_t)n_nodes_ || sign.size() != 4 * (py::ssize_t)n_nodes_)
      throw std::runtime_error("partner and sign need 4 N entries");
    check(fpb_set_macro(h(), partner.data(), sign.data()));
  }
  void run_macro_injected(carray<uint8_t> state, carray<float> mean, carray<double> z, int stage) {
    const py::ssize_t M = 4 * (py::ssize_t)n_nodes_;
    if (state.ndim() != 2 || mean.ndim() != 2 || z.ndim() != 2 || state.shape(1) != M || mean.shape(1) != M ||
        z.shape(1) != M || mean.shape(0) != state.shape(0) || z.shape(0) != state.shape(0))
      throw std::runtime_error("state uint8[T, 4N], mean float32[T, 4N] and z float64[T, 4N] expected");
    const int64_t T = state.shape(0);
    const uint8_t* sp = state.data();
    const float* mp = mean.data();
    const double* zp = z.data();
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_run_macro_injected(h(), T, sp, mp, zp, 0, stage);
    }
    check(rc);
  }
  void run_macro_rng(uint64_t seed, int64_t first_trial, int64_t n_trials, int stage) {
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_run_macro_rng(h(), seed, first_trial, n_trials, stage);
    }
    check(rc);
  }
  // reduced lattice of the last macronode run: freshly owned (red_state, bit_val, p_phase_cond, outcome), each [T, N]
  py::tuple fetch_macro() {
    fpb_sizes s;
    check(fpb_get_sizes(h(), &s));
    const py::ssize_t T = (py::ssize_t)s.n_trials, N = (py::ssize_t)n_nodes_;
    py::array_t<uint8_t> rs({T, N}), bv({T, N});
    py::array_t<double> pp({T, N}), oc({T, N});
    uint8_t *rp = rs.mutable_data(), *bp = bv.mutable_data();
    double *ppp = pp.mutable_data(), *op = oc.mutable_data();
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_fetch_macro(h(), rp, bp, ppp, op, 0);
    }
    check(rc);
    return py::make_tuple(rs, bv, pp, oc);
  }
  void run_iid(uint64_t seed, int64_t first_trial, int64_t n_trials, double p_err, int stage) {
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_run_iid(h(), seed, first_trial, n_trials, p_err, stage);
    }
    check(rc);
  }

  py::dict sizes() const {
    fpb_sizes s;
    check(fpb_get_sizes(h(), &s));
    py::dict d;
    d["n_trials"] = s.n_trials;
    d["n_complex"] = s.n_complex;
    d["total_qubits"] = s.total_qubits;
    py::list odd, pairs;
    for (int i = 0; i < FPB_MAX_COMPLEX; ++i) {
      odd.append(s.total_odd[i]);
      pairs.append(s.total_pairs[i]);
    }
    d["total_odd"] = odd;
    d["total_pairs"] = pairs;
    return d;
  }

  // Copy results of the last run into caller-owned arrays (keys as in fpb_outputs; per-complex
  // arrays are named "<field><complex index>", e.g. "pair_len0").  Absent keys are skipped.
  void fetch_into(const py::dict& bufs) {
    fpb_sizes s;
    check(fpb_get_sizes(h(), &s));
    const size_t T = (size_t)s.n_trials;
    fpb_outputs o;
    std::memset(&o, 0, sizeof o);
    o.hom = dest<double>(bufs, "hom", T * qtot_);
    o.weights = dest<double>(bufs, "weights", T * qtot_);
    o.bits = dest<uint32_t>(bufs, "bits", T * bit_words_);
    o.state = dest<uint8_t>(bufs, "state", T * n_nodes_);
    o.x = dest<double>(bufs, "x", T * 2 * n_nodes_);
    for (int i = 0; i < ncx_; ++i) {
      const std::string k = std::to_string(i);
      o.parity[i] = dest<uint32_t>(bufs, "parity" + k, T * ((cubes_[i] + 31) / 32));
      o.odd_count[i] = dest<int32_t>(bufs, "odd_count" + k, T);
      o.odd_ids[i] = dest<int32_t>(bufs, "odd_ids" + k, (size_t)s.total_odd[i]);
      o.pair_len[i] = dest<double>(bufs, "pair_len" + k, (size_t)s.total_pairs[i]);
      if (has_bnd_[i]) {
        o.bnd_len[i] = dest<double>(bufs, "bnd_len" + k, (size_t)s.total_odd[i]);
        o.bnd_side[i] = dest<uint8_t>(bufs, "bnd_side" + k, (size_t)s.total_odd[i]);
      }
    }
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_fetch(h(), &o, 0);
    }
    check(rc);
  }

  // lemonpy-style convenience: freshly owned array of one complex's packed pair lengths
  py::array_t<double> pair_lengths(int complex_index) {
    fpb_sizes s;
    check(fpb_get_sizes(h(), &s));
    if (complex_index < 0 || complex_index >= s.n_complex) throw std::runtime_error("bad complex index");
    py::array_t<double> out(s.total_pairs[complex_index]);
    fpb_outputs o;
    std::memset(&o, 0, sizeof o);
    o.pair_len[complex_index] = out.mutable_data();
    check(fpb_fetch(h(), &o, 0));
    return out;
  }

  py::dict device_outputs() const {
    fpb_outputs o;
    check(fpb_device_outputs(h(), &o));
    auto addr = [](const void* p) { return reinterpret_cast<std::uintptr_t>(p); };
    py::dict d;
    d["hom"] = addr(o.hom);
    d["bits"] = addr(o.bits);
    d["weights"] = addr(o.weights);
    d["state"] = addr(o.state);
    d["x"] = addr(o.x);
    for (int i = 0; i < ncx_; ++i) {
      const std::string k = std::to_string(i);
      d[("parity" + k).c_str()] = addr(o.parity[i]);
      d[("odd_count" + k).c_str()] = addr(o.odd_count[i]);
      d[("odd_ids" + k).c_str()] = addr(o.odd_ids[i]);
      d[("pair_len" + k).c_str()] = addr(o.pair_len[i]);
      d[("bnd_len" + k).c_str()] = addr(o.bnd_len[i]);
      d[("bnd_side" + k).c_str()] = addr(o.bnd_side[i]);
    }
    return d;
  }

  py::dict timings() const {
    fpb_timings t;
    check(fpb_get_timings(h(), &t));
    py::dict d;
    d["gen_ms"] = t.gen_ms;
    d["noise_ms"] = t.noise_ms;
    d["syndrome_ms"] = t.syndrome_ms;
    d["graph_ms"] = t.graph_ms;
    d["total_ms"] = t.total_ms;
    d["launches"] = t.launches;
    d["graph_ctas"] = t.graph_ctas;
    d["graph_warps_per_cta"] = t.graph_warps_per_cta;
    d["graph_smem_bytes"] = t.graph_smem_bytes;
    return d;
  }

  void synchronize() {
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_synchronize(h());
    }
    check(rc);
  }
  void build_table() {
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_build_table(h());
    }
    check(rc);
  }
  void mark(int which) { check(fpb_mark(h(), which)); }
  float elapsed_ms() {
    float ms = 0;
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_elapsed_ms(h(), &ms);
    }
    check(rc);
    return ms;
  }

  // candidate partners of the last run into `out` (int32, at least total_odd * k entries)
  void knn(int complex_index, int k, py::array out) {
    fpb_sizes s;
    check(fpb_get_sizes(h(), &s));
    if (complex_index < 0 || complex_index >= s.n_complex) throw std::runtime_error("bad complex index");
    if (!(out.flags() & py::array::c_style) || !out.writeable() || out.itemsize() != 4 ||
        (int64_t)out.size() < s.total_odd[complex_index] * k)
      throw std::runtime_error("knn: out must be a writable C-contiguous int32 array of total_odd * k entries");
    int32_t* p = static_cast<int32_t*>(out.mutable_data());
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_knn(h(), complex_index, k, p, 0);
    }
    check(rc);
  }

  // sparse hand-off: candidates with reduced lengths, (ids int32[n, k], lens float64[n, k], wmax float64[T])
  py::tuple knn_len(int complex_index, int k) {
    fpb_sizes s;
    check(fpb_get_sizes(h(), &s));
    if (complex_index < 0 || complex_index >= s.n_complex) throw std::runtime_error("bad complex index");
    const py::ssize_t n = (py::ssize_t)s.total_odd[complex_index];
    py::array_t<int32_t> ids({n, (py::ssize_t)k});
    py::array_t<double> lens({n, (py::ssize_t)k});
    py::array_t<double> wmax((py::ssize_t)s.n_trials);
    int32_t* ip = ids.mutable_data();
    double *lp = lens.mutable_data(), *wp = wmax.mutable_data();
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_knn_len(h(), complex_index, k, ip, lp, wp);
    }
    check(rc);
    return py::make_tuple(ids, lens, wmax);
  }
  // the same into caller-owned (e.g. pinned) arrays: ids int32 / lens float64 with >= total_odd * k
  // entries, wmax float64 with >= n_trials entries
  void knn_len_into(int complex_index, int k, py::array ids, py::array lens, py::array wmax) {
    fpb_sizes s;
    check(fpb_get_sizes(h(), &s));
    if (complex_index < 0 || complex_index >= s.n_complex) throw std::runtime_error("bad complex index");
    const int64_t need = s.total_odd[complex_index] * (int64_t)k;
    auto ok = [](const py::array& a, size_t item, int64_t n) {
      return (a.flags() & py::array::c_style) && a.writeable() && (size_t)a.itemsize() == item && (int64_t)a.size() >= n;
    };
    if (!ok(ids, 4, need) || !ok(lens, 8, need) || !ok(wmax, 8, s.n_trials))
      throw std::runtime_error("knn_len_into: buffers must be writable, C-contiguous and large enough");
    int32_t* ip = static_cast<int32_t*>(ids.mutable_data());
    double* lp = static_cast<double*>(lens.mutable_data());
    double* wp = static_cast<double*>(wmax.mutable_data());
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_knn_len(h(), complex_index, k, ip, lp, wp);
    }
    check(rc);
  }
  py::array_t<double> pairs_at(int complex_index, carray<int32_t> trial, carray<int32_t> a, carray<int32_t> b) {
    const int64_t n = trial.size();
    if (a.size() != n || b.size() != n) throw std::runtime_error("pair arrays differ in length");
    py::array_t<double> out((py::ssize_t)n);
    const int32_t *tp = trial.data(), *ap = a.data(), *bp = b.data();
    double* op = out.mutable_data();
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_pairs_at(h(), complex_index, n, tp, ap, bp, op);
    }
    check(rc);
    return out;
  }
  // device-side pricing: (trial, a, b, reduced length) of every pair violating the vertex-dual test
  py::array_t<double> pairs_of_trial(int complex_index, int64_t trial, int64_t n_pairs) {
    py::array_t<double> out((py::ssize_t)n_pairs);
    double* op = out.mutable_data();
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_pairs_of_trial(h(), complex_index, trial, op, n_pairs);
    }
    check(rc);
    return out;
  }
  py::tuple price(int complex_index, carray<int64_t> y, py::object top_o, py::object ztop_o, carray<double> scale,
                  carray<int32_t> wshift, carray<uint8_t> active, int64_t cap) {
    carray<int32_t> top;
    carray<int64_t> ztop;
    const bool have_top = !top_o.is_none();
    if (have_top) {
      top = carray<int32_t>::ensure(top_o);
      ztop = carray<int64_t>::ensure(ztop_o);
      if (!top || !ztop || top.size() < y.size() || ztop.size() < y.size())
        throw std::runtime_error("price: top / ztop need one entry per vertex dual");
    }
    fpb_sizes s;
    check(fpb_get_sizes(h(), &s));
    if (complex_index < 0 || complex_index >= s.n_complex) throw std::runtime_error("bad complex index");
    if (y.size() < s.total_odd[complex_index] || scale.size() < s.n_trials || wshift.size() < s.n_trials ||
        active.size() < s.n_trials)
      throw std::runtime_error("price: y needs total_odd entries, scale / wshift / active one per trial");
    const int64_t* yp = y.data();
    const int32_t* topp = have_top ? top.data() : nullptr;
    const int64_t* ztp = have_top ? ztop.data() : nullptr;
    const double* cp = scale.data();
    const int32_t* mp = wshift.data();
    const uint8_t* ap = active.data();
    for (;;) {
      py::array_t<int32_t> vt((py::ssize_t)cap), va((py::ssize_t)cap), vb((py::ssize_t)cap);
      py::array_t<double> vl((py::ssize_t)cap);
      int32_t *tp = vt.mutable_data(), *up = va.mutable_data(), *wp = vb.mutable_data();
      double* lp = vl.mutable_data();
      int64_t found = 0;
      int rc;
      {
        py::gil_scoped_release release;
        rc = fpb_price(h(), complex_index, yp, topp, ztp, cp, mp, ap, cap, tp, up, wp, lp, &found);
      }
      check(rc);
      if (found <= cap) {
        vt.resize({(py::ssize_t)found});
        va.resize({(py::ssize_t)found});
        vb.resize({(py::ssize_t)found});
        vl.resize({(py::ssize_t)found});
        return py::make_tuple(vt, va, vb, vl);
      }
      cap = found + found / 8 + 64;  // not enough room: once more with what was needed
    }
  }

  // correction paths of matched pairs: XOR-toggles the crossed qubits into flips uint32[T, bit_words]
  void paths_toggle(carray<int32_t> trial, carray<int32_t> cx, carray<int32_t> src, carray<int32_t> dst,
                    py::array flips) {
    const int64_t n = trial.size();
    if (cx.size() != n || src.size() != n || dst.size() != n) throw std::runtime_error("query arrays differ in length");
    fpb_sizes s;
    check(fpb_get_sizes(h(), &s));
    if (!(flips.flags() & py::array::c_style) || !flips.writeable() || flips.itemsize() != 4 ||
        (int64_t)flips.size() < s.n_trials * (int64_t)bit_words_)
      throw std::runtime_error("paths_toggle: flips must be a writable C-contiguous uint32[T, bit_words] array");
    const int32_t *tp = trial.data(), *cp = cx.data(), *sp = src.data(), *dp = dst.data();
    uint32_t* fp = static_cast<uint32_t*>(flips.mutable_data());
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_paths_toggle(h(), n, tp, cp, sp, dp, fp);
    }
    check(rc);
  }

  // ordered correction paths: (qubits int32[n, max_len], lengths int32[n]) - freshly owned arrays
  py::tuple paths_list(carray<int32_t> trial, carray<int32_t> cx, carray<int32_t> src, carray<int32_t> dst,
                       int max_len) {
    const int64_t n = trial.size();
    if (cx.size() != n || src.size() != n || dst.size() != n) throw std::runtime_error("query arrays differ in length");
    if (max_len < 1) throw std::runtime_error("max_len must be positive");
    py::array_t<int32_t> qubits({(py::ssize_t)n, (py::ssize_t)max_len});
    py::array_t<int32_t> length((py::ssize_t)n);
    const int32_t *tp = trial.data(), *cp = cx.data(), *sp = src.data(), *dp = dst.data();
    int32_t *qp = qubits.mutable_data(), *lp = length.mutable_data();
    int rc;
    {
      py::gil_scoped_release release;
      rc = fpb_paths_list(h(), n, tp, cp, sp, dp, max_len, qp, lp);
    }
    check(rc);
    return py::make_tuple(qubits, length);
  }

 private:
  fpb_handle* h_ = nullptr;
  size_t n_nodes_ = 0, qtot_ = 0, bit_words_ = 0;
  int ncx_ = 0;
  int cubes_[FPB_MAX_COMPLEX] = {0, 0};
  bool has_bnd_[FPB_MAX_COMPLEX] = {false, false};
};

}  // namespace

PYBIND11_MODULE(fpbpy, m) {
  m.doc() = "pybind11 plugin over the B200 Monte-Carlo EC trial backend (C ABI: include/fpb.h)";
  m.def("abi_version", &fpb_abi_version);
  m.def("set_error_class", [](py::object cls) {
    Py_XDECREF(g_error_class);
    g_error_class = cls.is_none() ? nullptr : cls.inc_ref().ptr();
  });
  m.def("device_count", [] {
    int n = fpb_device_count();
    if (n < 0) throw std::runtime_error(fpb_last_error());
    return n;
  });
  py::class_<Handle>(m, "Handle")
      .def(py::init<const py::dict&>(), py::arg("config"))
      .def("close", &Handle::close)
      .def("address", &Handle::address)
      .def("run_injected", &Handle::run_injected, py::arg("x"), py::arg("state"), py::arg("stage") = FPB_STAGE_GRAPH)
      .def("run_rng", &Handle::run_rng, py::arg("seed"), py::arg("first_trial"), py::arg("n_trials"),
           py::arg("stage") = FPB_STAGE_GRAPH)
      .def("run_bits", &Handle::run_bits, py::arg("bits"), py::arg("stage") = FPB_STAGE_GRAPH)
      .def("run_weighted", &Handle::run_weighted, py::arg("weights"), py::arg("bits"), py::arg("stage") = FPB_STAGE_GRAPH)
      .def("set_macro", &Handle::set_macro)
      .def("run_macro_injected", &Handle::run_macro_injected, py::arg("state"), py::arg("mean"), py::arg("z"),
           py::arg("stage") = FPB_STAGE_GRAPH)
      .def("run_macro_rng", &Handle::run_macro_rng, py::arg("seed"), py::arg("first_trial"), py::arg("n_trials"),
           py::arg("stage") = FPB_STAGE_GRAPH)
      .def("fetch_macro", &Handle::fetch_macro)
      .def("run_iid", &Handle::run_iid, py::arg("seed"), py::arg("first_trial"), py::arg("n_trials"), py::arg("p_err"),
           py::arg("stage") = FPB_STAGE_GRAPH)
      .def("sizes", &Handle::sizes)
      .def("fetch_into", &Handle::fetch_into, py::arg("bufs"))
      .def("pair_lengths", &Handle::pair_lengths, py::arg("complex_index") = 0)
      .def("device_outputs", &Handle::device_outputs)
      .def("timings", &Handle::timings)
      .def("synchronize", &Handle::synchronize)
      .def("build_table", &Handle::build_table)
      .def("mark", &Handle::mark)
      .def("elapsed_ms", &Handle::elapsed_ms)
      .def("knn", &Handle::knn, py::arg("complex_index"), py::arg("k"), py::arg("out"))
      .def("knn_len", &Handle::knn_len, py::arg("complex_index"), py::arg("k"))
      .def("knn_len_into", &Handle::knn_len_into)
      .def("pairs_at", &Handle::pairs_at)
      .def("pairs_of_trial", &Handle::pairs_of_trial)
      .def("price", &Handle::price, py::arg("complex_index"), py::arg("y"), py::arg("top"), py::arg("ztop"),
           py::arg("scale"), py::arg("wshift"),
           py::arg("active"), py::arg("cap") = 65536)
      .def("paths_toggle", &Handle::paths_toggle)
      .def("paths_list", &Handle::paths_list);
}
