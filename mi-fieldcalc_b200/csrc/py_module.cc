// py_module.cc -- Python module `mi_fieldcalc`: the pybind11 subset of the reference
// (python/py_mi_fieldcalc.cc:189-207 of mi-fieldcalc) on top of the drop-in C++ API
// (include/mi_fieldcalc/FieldCalculations.h -> libmi-fieldcalc.so.0 -> libfcb200.so -> sm_100a kernels).
//
// Same conventions as the reference module:
//   * every array argument is converted to a C-contiguous float32 numpy array; all arrays must be
//     2-D and of equal shape, otherwise the function returns None (reference :82-84);
//   * nx = shape[0], ny = shape[1] (reference :89 -- the operators of the subset are point-wise, so
//     the order does not matter);
//   * the ValuesDefined flag handed to the operator is SOME_DEFINED (:90); it is not returned;
//   * the GIL is released for the call (:75); a `false` from the operator becomes None (:92-93).
// All 15 functions of the reference module are exported.  showalterIndex leaves output points with an
// undefined input untouched (FC.cc:966-968); like the reference module, the output array is freshly
// allocated, so those points hold whatever the allocation held.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <string>
#include <utility>

#include "mi_fieldcalc/FieldCalculations.h"

namespace py = pybind11;
namespace fc = miutil::fieldcalc;

namespace {

typedef py::array_t<float, py::array::c_style | py::array::forcecast> farray;

// shape bookkeeping over a mixed argument pack: arrays must agree with the first one, scalars are ignored
struct Shape
{
  bool ok = true;
  py::ssize_t d0 = -1, d1 = -1;
  void see(const farray& a)
  {
    if (a.ndim() != 2) {
      ok = false;
      return;
    }
    if (d0 < 0) {
      d0 = a.shape(0);
      d1 = a.shape(1);
    } else if (a.shape(0) != d0 || a.shape(1) != d1) {
      ok = false;
    }
  }
  template <class T>
  void see(const T&)
  {
  }
};

inline const float* raw(const farray& a) { return a.data(); }
template <class T>
inline const T& raw(const T& v)
{
  return v;
}

// op(nx, ny, args..., out, fDefined, undef) with numpy in, numpy (or None) out
template <class Op, class... Args>
py::object apply(Op op, float undef, const Args&... args)
{
  Shape shape;
  int dummy[] = {(shape.see(args), 0)...};
  (void)dummy;
  if (!shape.ok || shape.d0 < 0)
    return py::none();
  py::array_t<float> out({shape.d0, shape.d1});
  float* dst = out.mutable_data();
  miutil::ValuesDefined defined = miutil::SOME_DEFINED;
  bool good;
  {
    py::gil_scoped_release nogil;
    good = op((int)shape.d0, (int)shape.d1, raw(args)..., dst, defined, undef);
  }
  if (!good)
    return py::none();
  return std::move(out);
}

} // namespace

PYBIND11_MODULE(mi_fieldcalc, m)
{
  m.doc() = "mi-fieldcalc FieldCalculations on NVIDIA B200 (sm_100a CUDA kernels behind the reference's API)";
  py::enum_<miutil::ValuesDefined>(m, "ValuesDefined")
      .value("ALL_DEFINED", miutil::ALL_DEFINED)
      .value("NONE_DEFINED", miutil::NONE_DEFINED)
      .value("SOME_DEFINED", miutil::SOME_DEFINED);

  m.def("windCooling", [](farray t, farray u, farray v, int compute, float undef) { return apply(fc::windCooling, undef, t, u, v, compute); });
  m.def("vesselIcingOverland", [](farray airtemp, farray seatemp, farray u, farray v, farray sal, farray aice, float undef) {
    return apply(fc::vesselIcingOverland, undef, airtemp, seatemp, u, v, sal, aice);
  });
  m.def("vesselIcingMertins", [](farray airtemp, farray seatemp, farray u, farray v, farray sal, farray aice, float undef) {
    return apply(fc::vesselIcingMertins, undef, airtemp, seatemp, u, v, sal, aice);
  });
  m.def("vesselIcingModStall", [](farray sal, farray wave, farray x_wind, farray y_wind, farray airtemp, farray rh, farray sst, farray p, farray Pw,
                                  farray aice, farray depth, float vs, float alpha, float zmin, float zmax, float undef) {
    return apply(fc::vesselIcingModStall, undef, sal, wave, x_wind, y_wind, airtemp, rh, sst, p, Pw, aice, depth, vs, alpha, zmin, zmax);
  });
  m.def("vesselIcingMincog", [](farray sal, farray wave, farray x_wind, farray y_wind, farray airtemp, farray rh, farray sst, farray p, farray Pw,
                                farray aice, farray depth, float vs, float alpha, float zmin, float zmax, int alt, float undef) {
    return apply(fc::vesselIcingMincog, undef, sal, wave, x_wind, y_wind, airtemp, rh, sst, p, Pw, aice, depth, vs, alpha, zmin, zmax, alt);
  });
  m.def("kIndex", [](farray t500, farray t700, farray rh700, farray t850, farray rh850, float p500, float p700, float p850, int compute, float undef) {
    return apply(fc::kIndex, undef, t500, t700, rh700, t850, rh850, p500, p700, p850, compute);
  });
  m.def("ductingIndex", [](farray t850, farray rh850, float p850, int compute, float undef) { return apply(fc::ductingIndex, undef, t850, rh850, p850, compute); });
  m.def("showalterIndex", [](farray t500, farray t850, farray rh850, float p500, float p850, int compute, float undef) {
    return apply(fc::showalterIndex, undef, t500, t850, rh850, p500, p850, compute);
  });
  m.def("boydenIndex", [](farray t700, farray z700, farray z1000, float p700, float p1000, int compute, float undef) {
    return apply(fc::boydenIndex, undef, t700, z700, z1000, p700, p1000, compute);
  });
  m.def("sweatIndex", [](farray t850, farray t500, farray td850, farray td500, farray u850, farray v850, farray u500, farray v500, float undef) {
    return apply(fc::sweatIndex, undef, t850, t500, td850, td500, u850, v850, u500, v500);
  });
  m.def("seaSoundSpeed", [](farray t, farray s, float z, int compute, float undef) { return apply(fc::seaSoundSpeed, undef, t, s, z, compute); });
  m.def("cvtemp", [](farray tinp, int compute, float undef) { return apply(fc::cvtemp, undef, tinp, compute); });
  m.def("cvhum", [](farray t, farray huminp, const std::string& unit, int compute, float undef) { return apply(fc::cvhum, undef, t, huminp, unit, compute); });
  m.def("abshum", [](farray t, farray rhum, float undef) { return apply(fc::abshum, undef, t, rhum); });
  m.def("underCooledRain", [](farray precip, farray snow, farray tk, float precipMin, float snowRateMax, float tcMax, float undef) {
    return apply(fc::underCooledRain, undef, precip, snow, tk, precipMin, snowRateMax, tcMax);
  });
}
