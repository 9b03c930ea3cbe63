// pymodule.cpp -- `_pbkpm`: the pybind11 binding of libpbkpm.so, i.e. the B200 counterpart of the reference's
// `_pybinding.kpm(...)` factory and `_pybinding.KPM` class (cppmodule/src/kpm.cpp:8-37,76-102) together with the
// cpb::KPM facade they wrap (cppcore/src/KPM.cpp:7-148).  Host C++ only: every calculation is one call through the
// C ABI of include/pbkpm.h with the GIL released (cppmodule/include/wrappers.hpp:15), errors come back as the same
// Python exception types pybind11 gives the reference's C++ exceptions (std::invalid_argument -> ValueError,
// std::runtime_error / std::logic_error -> RuntimeError).
//
// The model is duck-typed exactly like cpb::KPM reads it: `model.hamiltonian` (scipy CSR: f32 / c64 / f64 / c128,
// cppmodule/src/model.cpp:29-31), `model.system.{find_nearest, to_hamiltonian_indices, positions, expanded_positions,
// expanded_positions}` (cppmodule/src/system.cpp:85-94), `model.eval()`, `model.is_multiorbital`.  The sublattice range of
// calc_spatial_ldos (System::sublattice_range, cppcore/src/system/System.cpp:50-66, not bound by the reference) comes from
// `pybinding_b200.chebyshev.sublattice_range`, which reads `system.sublattices` on a real pybinding System.
#include <pybind11/pybind11.h>
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/stl.h>

#include <complex>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pbkpm.h"

namespace py = pybind11;
using namespace py::literals;
using cd = std::complex<double>;

namespace {

[[noreturn]] void rethrow(int status, pbk_ctx* ctx) {
    char const* m = pbk_last_error(ctx);
    std::string msg = (m && *m) ? m : ("pbkpm error " + std::to_string(status));
    if (status == PBK_INVALID_ARGUMENT) throw std::invalid_argument(msg);   // -> ValueError
    throw std::runtime_error(msg);                                          // -> RuntimeError
}
inline void check(int status, pbk_ctx* ctx) { if (status != PBK_OK) rethrow(status, ctx); }

template<class T> using carray = py::array_t<T, py::array::c_style | py::array::forcecast>;

// ---- kernels (cppmodule/src/kpm.cpp:68-74, cppcore/src/kpm/Kernel.cpp) ---------------------------------
struct Kernel {
    int id = PBK_JACKSON;
    double lambda_value = 4.0;

    py::array_t<double> damping_coefficients(int n) const {
        py::array_t<double> out(n);
        check(pbk_kernel_damping(id, lambda_value, n, out.mutable_data()), nullptr);
        return out;
    }
    int required_num_moments(double scaled_broadening) const {
        int32_t m = 0;
        check(pbk_kernel_required_num_moments(id, lambda_value, scaled_broadening, &m), nullptr);
        return m;
    }
};

Kernel kernel_from_object(py::object const& k) {
    if (k.is_none()) return Kernel{};
    if (py::isinstance<Kernel>(k)) return k.cast<Kernel>();
    if (py::isinstance<py::str>(k)) {
        auto const name = k.cast<std::string>();
        if (name == "default" || name == "jackson") return Kernel{PBK_JACKSON, 4.0};
        if (name == "lorentz") return Kernel{PBK_LORENTZ, 4.0};
        if (name == "dirichlet") return Kernel{PBK_DIRICHLET, 4.0};
        throw std::invalid_argument("unknown KPM kernel '" + name + "'");
    }
    // kernel objects of the ctypes layer (pybinding_b200.chebyshev.KPMKernel) carry .kind / .lambda_value
    if (py::hasattr(k, "kind")) return Kernel{k.attr("kind").cast<int>(), py::hasattr(k, "lambda_value") ? k.attr("lambda_value").cast<double>() : 4.0};
    throw std::invalid_argument("kernel must be a KPMKernel");
}

struct Stats {  // kpm::Stats as exposed by cppmodule/src/kpm.cpp:50-66, plus the GPU counters of pbk_stats
    pbk_stats s{};
    // Stats::ops (cppcore/src/kpm/Stats.cpp:53-58): 2 operations per non-zero, 5 per vector element; same formula as
    // pybinding_b200.chebyshev.KPMStats.ops
    double ops() const { return s.moments_time > 0 ? s.multiplier * (2.0 * static_cast<double>(s.nnz) + 5.0 * static_cast<double>(s.vec)) / s.moments_time : 0.0; }
};

// ---- the KPM object --------------------------------------------------------------------------------------
class KPM {
public:
    KPM(py::object model_, std::pair<float, float> energy_range, Kernel kernel_, std::string const& matrix_format, bool optimal_size,
        bool interleaved, float lanczos_precision, py::object progress_callback, int device, int max_batch, int locality_tile)
        : kernel(kernel_) {
        pbk_config cfg{};
        cfg.min_energy = energy_range.first;
        cfg.max_energy = energy_range.second;
        cfg.kernel = kernel.id;
        cfg.lambda_value = kernel.lambda_value;
        cfg.optimal_size = optimal_size;
        cfg.interleaved = interleaved;
        cfg.matrix_format = matrix_format == "ELL";
        cfg.lanczos_precision = lanczos_precision;
        cfg.max_batch = max_batch;
        cfg.locality_tile = locality_tile;
        check(pbk_create(&ctx, device, &cfg), nullptr);
        try {
            if (!progress_callback.is_none()) {
                progress = progress_callback;
                check(pbk_set_progress_callback(ctx, &KPM::progress_trampoline, this), ctx);
            }
            set_model(std::move(model_));
        } catch (...) {
            pbk_destroy(ctx);
            ctx = nullptr;
            throw;
        }
    }
    KPM(KPM const&) = delete;
    KPM& operator=(KPM const&) = delete;
    ~KPM() { if (ctx) pbk_destroy(ctx); }

    // cpb::KPM::set_model (KPM.cpp:10-13): re-evaluates the model, uploads the Hamiltonian, resets the bounds
    void set_model(py::object m) {
        if (py::hasattr(m, "eval")) {
            py::object e = m.attr("eval")();
            if (!e.is_none()) m = e;
        }
        py::object h = m.attr("hamiltonian").attr("tocsr")();
        if (!h.attr("has_sorted_indices").cast<bool>()) { h = h.attr("copy")(); h.attr("sort_indices")(); }
        py::dtype const dt = h.attr("dtype").cast<py::dtype>();
        int dtype = -1;
        if (dt.is(py::dtype::of<float>())) dtype = PBK_F32;
        else if (dt.is(py::dtype::of<std::complex<float>>())) dtype = PBK_C64;
        else if (dt.is(py::dtype::of<double>())) dtype = PBK_F64;
        else if (dt.is(py::dtype::of<cd>())) dtype = PBK_C128;
        else throw py::type_error("unsupported Hamiltonian dtype");
        auto const shape = h.attr("shape").cast<std::pair<int64_t, int64_t>>();
        if (h.attr("nnz").cast<int64_t>() >= (int64_t{1} << 31)) throw std::invalid_argument("the Hamiltonian has too many non-zeros for int32 indices");
        carray<int32_t> indptr(h.attr("indptr")), indices(h.attr("indices"));
        py::array data = py::array::ensure(h.attr("data"), py::array::c_style);
        {
            py::gil_scoped_release nogil;
            check(pbk_set_hamiltonian(ctx, dtype, shape.first, indptr.data(), indices.data(), data.data()), ctx);
        }
        model = std::move(m);
        size = shape.first;
        is_complex = dtype == PBK_C64 || dtype == PBK_C128;
    }
    py::object get_model() const { return model; }
    py::object system() const { return model.attr("system"); }

    std::pair<double, double> scaling_factors() {
        double a = 0, b = 0;
        { py::gil_scoped_release nogil; check(pbk_scaling_factors(ctx, &a, &b), ctx); }
        return {a, b};
    }

    // cpb::KPM::moments (KPM.cpp:19-45)
    py::array_t<cd> moments(int num_moments, py::object alpha_, py::object beta_, py::object op_) {
        carray<cd> alpha(py::array::ensure(alpha_).attr("ravel")());
        carray<cd> beta = beta_.is_none() ? carray<cd>(0) : carray<cd>(py::array::ensure(beta_).attr("ravel")());
        bool has_op = false;
        int64_t op_rows = 0;
        carray<int32_t> ip(0), ix(0);
        carray<cd> od(0);
        bool op_size_ok = true;
        if (!op_.is_none()) {
            auto const shp = op_.attr("shape").cast<std::pair<int64_t, int64_t>>();
            int64_t const op_size = shp.first * shp.second;
            op_size_ok = op_size == 0 || (shp.first == size && shp.second == size);
            has_op = op_size > 1;
            if (has_op && op_size_ok) {
                py::object op = op_.attr("tocsr")();
                op.attr("sort_indices")();
                op_rows = shp.first;
                ip = carray<int32_t>(op.attr("indptr"));
                ix = carray<int32_t>(op.attr("indices"));
                od = carray<cd>(op.attr("data"));
            }
        }
        auto const mismatch = [](char const* name) {
            throw std::runtime_error(std::string("Size mismatch between the model Hamiltonian and the given argument '") + name + "'");
        };
        if (alpha.size() != size) mismatch("alpha");
        if (beta.size() != 0 && beta.size() != size) mismatch("beta");
        if (!op_size_ok) mismatch("operator");
        if (!is_complex) {
            auto const has_imag = [](carray<cd> const& a) {
                for (py::ssize_t i = 0; i < a.size(); ++i) if (a.data()[i].imag() != 0) return true;
                return false;
            };
            auto const complaint = [](char const* name) {
                throw std::runtime_error(std::string("The model Hamiltonian is real, but the given argument '") + name + "' is complex");
            };
            if (has_imag(alpha)) complaint("alpha");
            if (has_imag(beta)) complaint("beta");
            if (has_op && has_imag(od)) complaint("operator");
        }
        py::array_t<cd> out(num_moments);
        {
            py::gil_scoped_release nogil;
            check(pbk_moments(ctx, num_moments, alpha.data(), beta.size() ? beta.data() : nullptr, op_rows,
                              op_rows ? ip.data() : nullptr, op_rows ? ix.data() : nullptr, op_rows ? od.data() : nullptr,
                              out.mutable_data()), ctx);
        }
        return out;
    }

    // cpb::KPM::calc_greens / calc_greens_vector (KPM.cpp:47-62)
    std::vector<py::array_t<cd>> calc_greens_vector(int64_t row, std::vector<int32_t> const& cols, carray<double> energy, double broadening) {
        if (row < 0 || row >= size) throw std::logic_error("KPM::calc_greens(i,j): invalid value for i or j.");
        for (auto c : cols) if (c < 0 || c >= size) throw std::logic_error("KPM::calc_greens(i,j): invalid value for i or j.");
        std::vector<cd> flat(cols.size() * static_cast<size_t>(energy.size()));
        {
            py::gil_scoped_release nogil;
            check(pbk_calc_greens(ctx, static_cast<int32_t>(row), cols.data(), static_cast<int32_t>(cols.size()), energy.data(),
                                  static_cast<int32_t>(energy.size()), broadening, flat.data()), ctx);
        }
        std::vector<py::array_t<cd>> out;
        for (size_t i = 0; i < cols.size(); ++i) {
            py::array_t<cd> a(energy.size());
            std::memcpy(a.mutable_data(), flat.data() + i * energy.size(), sizeof(cd) * energy.size());
            out.push_back(std::move(a));
        }
        return out;
    }
    py::array_t<cd> calc_greens(int64_t row, int64_t col, carray<double> energy, double broadening) {
        return calc_greens_vector(row, {static_cast<int32_t>(col)}, std::move(energy), broadening)[0];
    }

    // cpb::KPM::calc_ldos (KPM.cpp:64-74): energy x orbitals, column-major like ArrayXXdCM
    py::array_t<double> ldos_indices(std::vector<int32_t> const& idx, carray<double> const& energy, double broadening) {
        py::array_t<double, py::array::f_style> out({energy.size(), static_cast<py::ssize_t>(idx.size())});
        {
            py::gil_scoped_release nogil;
            check(pbk_calc_ldos(ctx, energy.data(), static_cast<int32_t>(energy.size()), broadening, idx.data(),
                                static_cast<int32_t>(idx.size()), out.mutable_data()), ctx);
        }
        return out;
    }
    py::object calc_ldos(carray<double> energy, double broadening, py::object position, std::string const& sublattice, bool reduce) {
        py::object sys = system();
        py::object site = sys.attr("find_nearest")(position, sublattice);
        auto const idx = py::array_t<int32_t, py::array::c_style | py::array::forcecast>(py::module_::import("numpy").attr("atleast_1d")(
                             sys.attr("to_hamiltonian_indices")(site)));
        std::vector<int32_t> v(idx.data(), idx.data() + idx.size());
        py::array_t<double> out = ldos_indices(v, energy, broadening);
        if (reduce && v.size() > 1) return out.attr("sum")("axis"_a = 1, "keepdims"_a = true);
        return std::move(out);
    }
    // cpb::KPM::calc_spatial_ldos (KPM.cpp:76-101)
    py::array_t<double> calc_spatial_ldos(carray<double> energy, double broadening, py::object shape, std::string const& sublattice) {
        if (py::hasattr(model, "is_multiorbital") && model.attr("is_multiorbital").cast<bool>()) {
            throw std::runtime_error("This function doesn't currently support multi-orbital models");
        }
        py::object sys = system();
        py::object np = py::module_::import("numpy");
        py::object pos = sys.attr("positions");
        py::object contains = np.attr("asarray")(shape.attr("contains")(pos.attr("x"), pos.attr("y"), pos.attr("z")));
        auto const range = py::module_::import("pybinding_b200.chebyshev").attr("sublattice_range")(sys, sublattice).cast<std::pair<int64_t, int64_t>>();
        auto const mask = carray<bool>(contains);
        std::vector<int32_t> idx;
        for (int64_t i = range.first; i < range.second; ++i) if (mask.data()[i]) idx.push_back(static_cast<int32_t>(i));
        if (idx.empty()) throw std::runtime_error("calc_spatial_ldos: the shape contains no sites");
        return ldos_indices(idx, energy, broadening);
    }

    // cpb::KPM::calc_dos (KPM.cpp:103-110)
    py::array_t<double> calc_dos(carray<double> energy, double broadening, int num_random) {
        py::array_t<double> out(energy.size());
        {
            py::gil_scoped_release nogil;
            check(pbk_calc_dos(ctx, energy.data(), static_cast<int32_t>(energy.size()), broadening, num_random, out.mutable_data()), ctx);
        }
        return out;
    }

    // cpb::KPM::calc_conductivity (KPM.cpp:112-146)
    py::array_t<double> calc_conductivity(carray<double> mu, double broadening, double temperature, std::string const& direction,
                                          int num_random, int num_points) {
        auto const valid = [](char c) { return c == 'x' || c == 'y' || c == 'z'; };
        if (direction.size() != 2 || !valid(direction[0]) || !valid(direction[1])) {
            throw std::logic_error("Invalid direction: must be 'xx', 'xy', 'zz', or similar.");
        }
        py::object sys = system();
        bool const multi = py::hasattr(model, "is_multiorbital") && model.attr("is_multiorbital").cast<bool>();
        py::object pos = multi ? sys.attr("expanded_positions") : sys.attr("positions");
        auto const axis = [&](char c) { return carray<float>(pos.attr(std::string(1, c).c_str())); };
        carray<float> left = axis(direction[0]), right = axis(direction[1]);
        std::vector<cd> sigma(static_cast<size_t>(mu.size()));
        {
            py::gil_scoped_release nogil;
            check(pbk_calc_conductivity(ctx, left.data(), right.data(), mu.data(), static_cast<int32_t>(mu.size()), broadening,
                                        temperature, num_random, num_points, sigma.data()), ctx);
        }
        py::array_t<double> out(mu.size());
        for (py::ssize_t i = 0; i < mu.size(); ++i) out.mutable_data()[i] = sigma[static_cast<size_t>(i)].real();
        return out;
    }

    // both take the context's mutex, which a calculation running on another thread holds while its progress callback
    // waits for the GIL: never wait for that mutex with the GIL held
    std::string report(bool shortform) {
        std::vector<char> buf(4096);
        { py::gil_scoped_release nogil; check(pbk_report(ctx, shortform, buf.data(), static_cast<int64_t>(buf.size())), ctx); }
        return buf.data();
    }
    Stats stats() {
        Stats s;
        { py::gil_scoped_release nogil; check(pbk_get_stats(ctx, &s.s), ctx); }
        return s;
    }

    // raw (undamped) moments of the compute-strategy level, for parity tests
    py::array_t<cd> moments_dos(int num_moments, int num_random) {
        py::array_t<cd> out(num_moments);
        { py::gil_scoped_release nogil; check(pbk_moments_dos(ctx, num_moments, num_random, out.mutable_data()), ctx); }
        return out;
    }

    Kernel kernel;

private:
    static void progress_trampoline(int64_t delta, int64_t total, void* user) {
        auto* self = static_cast<KPM*>(user);
        py::gil_scoped_acquire gil;
        self->progress(delta, total);
    }

    pbk_ctx* ctx = nullptr;
    py::object model;
    py::object progress;
    int64_t size = 0;
    bool is_complex = false;
};

// Deferred result (cppmodule/include/thread.hpp:9-43): computes on first use of `.result`, `.compute()` may be called
// from a worker thread; `.solver` is the KPM object (pybinding/parallel.py reads `.solver.report()`).
class DeferredLdos {
public:
    DeferredLdos(py::object solver_, std::function<py::object()> fn_) : solver(std::move(solver_)), fn(std::move(fn_)) {}
    // one computation even when several threads ask: the first caller runs the job, the others wait for it with the
    // GIL released (the job itself needs the GIL to enter calc_ldos)
    void compute() {
        std::unique_lock<std::mutex> lock(mutex, std::defer_lock);
        { py::gil_scoped_release nogil; lock.lock(); }
        if (!done) { value = fn(); done = true; }
    }
    py::object result() { compute(); return value; }
    py::object solver;
private:
    std::function<py::object()> fn;
    py::object value;
    std::mutex mutex;
    bool done = false;
};

} // anonymous namespace

PYBIND11_MODULE(_pbkpm, m) {
    m.doc() = "pybind11 binding of libpbkpm.so: the B200 KPM engine behind pybinding's _pybinding.kpm / KPM interface";
    m.attr("abi_version") = pbk_version();

    py::class_<Kernel>(m, "KPMKernel")
        .def("damping_coefficients", &Kernel::damping_coefficients, "num_moments"_a)
        .def("required_num_moments", &Kernel::required_num_moments, "scaled_broadening"_a)
        .def_readonly("kernel_id", &Kernel::id)
        .def_readonly("lambda_value", &Kernel::lambda_value);
    m.def("jackson_kernel", [] { return Kernel{PBK_JACKSON, 4.0}; });
    m.def("lorentz_kernel", [](double lambda_value) {
        if (lambda_value <= 0) throw std::invalid_argument("Lorentz kernel: lambda must be positive.");
        return Kernel{PBK_LORENTZ, lambda_value};
    }, "lambda_value"_a = 4.0);
    m.def("dirichlet_kernel", [] { return Kernel{PBK_DIRICHLET, 4.0}; });

    py::class_<Stats>(m, "KPMStats")
        .def_property_readonly("num_moments", [](Stats const& s) { return s.s.num_moments; })
        .def_property_readonly("uses_full_system", [](Stats const& s) { return s.s.uses_full_system != 0; })
        .def_property_readonly("nnz", [](Stats const& s) { return s.s.nnz; })
        .def_property_readonly("opt_nnz", [](Stats const& s) { return s.s.opt_nnz; })
        .def_property_readonly("vec", [](Stats const& s) { return s.s.vec; })
        .def_property_readonly("opt_vec", [](Stats const& s) { return s.s.opt_vec; })
        .def_property_readonly("matrix_memory", [](Stats const& s) { return s.s.matrix_memory; })
        .def_property_readonly("vector_memory", [](Stats const& s) { return s.s.vector_memory; })
        .def_property_readonly("eps", [](Stats const& s) { return s.s.eps; })
        .def_property_readonly("ops", &Stats::ops)
        .def_property_readonly("hamiltonian_time", [](Stats const& s) { return s.s.hamiltonian_time; })
        .def_property_readonly("moments_time", [](Stats const& s) { return s.s.moments_time; })
        .def_property_readonly("kernel_launches", [](Stats const& s) { return s.s.kernel_launches; })
        .def_property_readonly("step_launches", [](Stats const& s) { return s.s.step_launches; })
        .def_property_readonly("bulk_launches", [](Stats const& s) { return s.s.bulk_launches; })
        .def_property_readonly("res_launches", [](Stats const& s) { return s.s.res_launches; })
        .def_property_readonly("persist_launches", [](Stats const& s) { return s.s.persist_launches; })
        .def_property_readonly("graph_launches", [](Stats const& s) { return s.s.graph_launches; })
        .def_property_readonly("step_ms", [](Stats const& s) { return s.s.step_ms; })
        .def_property_readonly("step_bytes", [](Stats const& s) { return s.s.step_bytes; })
        .def_property_readonly("moments_device_ms", [](Stats const& s) { return s.s.moments_device_ms; })
        .def_property_readonly("batch", [](Stats const& s) { return s.s.batch; });

    py::class_<DeferredLdos, std::shared_ptr<DeferredLdos>>(m, "DeferredXd")
        .def("compute", &DeferredLdos::compute)
        .def_readonly("solver", &DeferredLdos::solver)
        .def_property_readonly("result", &DeferredLdos::result);

    py::class_<KPM, std::shared_ptr<KPM>>(m, "KPM")
        .def("moments", &KPM::moments, "num_moments"_a, "alpha"_a, "beta"_a = py::none(), "op"_a = py::none())
        .def("calc_greens", &KPM::calc_greens, "i"_a, "j"_a, "energy"_a, "broadening"_a)
        .def("calc_greens", &KPM::calc_greens_vector, "i"_a, "j"_a, "energy"_a, "broadening"_a)
        .def("calc_dos", &KPM::calc_dos, "energy"_a, "broadening"_a, "num_random"_a)
        .def("calc_conductivity", &KPM::calc_conductivity, "chemical_potential"_a, "broadening"_a, "temperature"_a,
             "direction"_a = "xx", "num_random"_a = 1, "num_points"_a = 1000)
        .def("calc_ldos", &KPM::calc_ldos, "energy"_a, "broadening"_a, "position"_a, "sublattice"_a = "", "reduce"_a = true)
        .def("calc_spatial_ldos", &KPM::calc_spatial_ldos, "energy"_a, "broadening"_a, "shape"_a, "sublattice"_a = "")
        .def("deferred_ldos", [](py::object self, py::object energy, double broadening, py::object position, std::string sublattice) {
            py::object e = py::module_::import("numpy").attr("array")(energy, "dtype"_a = "float64");
            return std::make_shared<DeferredLdos>(self, [self, e, broadening, position, sublattice] {
                return self.attr("calc_ldos")(e, broadening, position, sublattice);
            });
        }, "energy"_a, "broadening"_a, "position"_a, "sublattice"_a = "")
        .def("moments_dos", &KPM::moments_dos, "num_moments"_a, "num_random"_a)
        .def("report", &KPM::report, "shortform"_a = false)
        .def_property("model", &KPM::get_model, &KPM::set_model)
        .def_property_readonly("system", &KPM::system)
        .def_property_readonly("scaling_factors", &KPM::scaling_factors)
        .def_readonly("kernel", &KPM::kernel)
        .def_property_readonly("stats", &KPM::stats);

    // same keywords as _pybinding.kpm (cppmodule/src/kpm.cpp:8-37); num_threads is accepted and ignored
    // (the reference's thread pool over SIMD batches is replaced by batched device passes), device / max_batch /
    // locality_tile are the extra knobs of pbk_config
    auto factory = [](py::object model, py::object energy_range, py::object kernel, std::string matrix_format, bool optimal_size,
                      bool interleaved, float lanczos_precision, py::object /*num_threads*/, py::object progress_callback, int device,
                      int max_batch, int locality_tile) {
        std::pair<float, float> er{0.f, 0.f};
        if (!energy_range.is_none()) {
            auto const seq = energy_range.cast<py::sequence>();
            if (py::len(seq) != 2) throw std::invalid_argument("energy_range must be a (min, max) pair");
            er = {seq[0].cast<float>(), seq[1].cast<float>()};
        }
        return std::make_shared<KPM>(std::move(model), er, kernel_from_object(kernel), matrix_format, optimal_size, interleaved,
                                     lanczos_precision, std::move(progress_callback), device, max_batch, locality_tile);
    };
    for (char const* name : {"kpm", "kpm_cuda"}) {
        m.def(name, factory, "model"_a, "energy_range"_a = py::make_tuple(0.f, 0.f), "kernel"_a = py::none(), "matrix_format"_a = "ELL",
              "optimal_size"_a = true, "interleaved"_a = true, "lanczos_precision"_a = 0.002f, "num_threads"_a = py::none(),
              "progress_callback"_a = py::none(), "device"_a = 0, "max_batch"_a = 0, "locality_tile"_a = 0);
    }
}
