// oracle/ref_wrap.cc -- TEST INFRASTRUCTURE, not product code.
//
// extern "C" adapter around the UNMODIFIED reference (mi-fieldcalc, compiled in place from
// /root/reference/src/mi_fieldcalc/*.cc by oracle/Makefile into oracle/_ref/). It exposes every
// entry of include/fcb200_api.inc with the prefix `fcref_` so that tests can drive the real
// reference, the C restatement (fco_*) and the CUDA product (fcb200_*) through one signature.
// Nothing here computes anything: each function forwards to miutil::fieldcalc::<name>
// (src/mi_fieldcalc/FieldCalculations.h) and converts `ValuesDefined&` <-> `int*`.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load the resulting library.

#include "mi_fieldcalc/FieldCalculations.h"

#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace fc = miutil::fieldcalc;
using miutil::ValuesDefined;

namespace {
struct Flag
{
  int* p;
  ValuesDefined v;
  explicit Flag(int* fp)
      : p(fp)
      , v(static_cast<ValuesDefined>(*fp))
  {
  }
  ~Flag() { *p = static_cast<int>(v); }
  operator ValuesDefined&() { return v; }
};

std::vector<float*> as_vector(const float* const* fields, int n)
{
  std::vector<float*> v(n > 0 ? n : 0);
  for (int i = 0; i < n; ++i)
    v[i] = const_cast<float*>(fields[i]);
  return v;
}

std::vector<ValuesDefined> as_flags(const int* f, int n)
{
  std::vector<ValuesDefined> v(n > 0 ? n : 0);
  for (int i = 0; i < n; ++i)
    v[i] = static_cast<ValuesDefined>(f[i]);
  return v;
}
} // namespace

extern "C" {

int fcref_openmp_threads()
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int fcref_pleveltemp(int nx, int ny, const float* tinp, float p, const char* unit, int compute, float* tout, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::pleveltemp(nx, ny, tinp, p, unit, compute, tout, f, undef);
}

int fcref_plevelhum(int nx, int ny, const float* t, const float* huminp, float p, const char* unit, int compute, float* humout, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::plevelhum(nx, ny, t, huminp, p, unit, compute, humout, f, undef);
}

int fcref_hleveltemp(int nx, int ny, const float* tinp, const float* ps, float alevel, float blevel, const char* unit, int compute, float* tout,
                     int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::hleveltemp(nx, ny, tinp, ps, alevel, blevel, unit, compute, tout, f, undef);
}

int fcref_hlevelthe(int nx, int ny, const float* t, const float* q, const float* ps, float alevel, float blevel, int compute, float* the, int* fDefined,
                    float undef)
{
  Flag f(fDefined);
  return fc::hlevelthe(nx, ny, t, q, ps, alevel, blevel, compute, the, f, undef);
}

int fcref_hlevelhum(int nx, int ny, const float* t, const float* huminp, const float* ps, float alevel, float blevel, const char* unit, int compute,
                    float* humout, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::hlevelhum(nx, ny, t, huminp, ps, alevel, blevel, unit, compute, humout, f, undef);
}

int fcref_hlevelducting(int nx, int ny, const float* t, const float* h, const float* ps, float alevel, float blevel, int compute, float* duct,
                        int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::hlevelducting(nx, ny, t, h, ps, alevel, blevel, compute, duct, f, undef);
}

int fcref_hlevelpressure(int nx, int ny, const float* ps, float alevel, float blevel, float* p, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::hlevelpressure(nx, ny, ps, alevel, blevel, p, f, undef);
}

int fcref_aleveltemp(int nx, int ny, const float* tinp, const float* p, const char* unit, int compute, float* tout, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::aleveltemp(nx, ny, tinp, p, unit, compute, tout, f, undef);
}

int fcref_alevelthe(int nx, int ny, const float* t, const float* q, const float* p, int compute, float* the, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::alevelthe(nx, ny, t, q, p, compute, the, f, undef);
}

int fcref_alevelhum(int nx, int ny, const float* t, const float* huminp, const float* p, const char* unit, int compute, float* humout, int* fDefined,
                    float undef)
{
  Flag f(fDefined);
  return fc::alevelhum(nx, ny, t, huminp, p, unit, compute, humout, f, undef);
}

int fcref_alevelducting(int nx, int ny, const float* t, const float* h, const float* p, int compute, float* duct, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::alevelducting(nx, ny, t, h, p, compute, duct, f, undef);
}

int fcref_ilevelgwind(int nx, int ny, const float* mpot, const float* xmapr, const float* ymapr, const float* fcoriolis, float* ug, float* vg,
                      int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::ilevelgwind(nx, ny, mpot, xmapr, ymapr, fcoriolis, ug, vg, f, undef);
}

int fcref_relvort(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, float* rvort, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::relvort(nx, ny, u, v, xmapr, ymapr, rvort, f, undef);
}

int fcref_absvort(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, const float* fcoriolis, float* avort,
                  int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::absvort(nx, ny, u, v, xmapr, ymapr, fcoriolis, avort, f, undef);
}

int fcref_divergence(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, float* diverg, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::divergence(nx, ny, u, v, xmapr, ymapr, diverg, f, undef);
}

int fcref_advection(int nx, int ny, const float* fld, const float* u, const float* v, const float* xmapr, const float* ymapr, float hours, float* advec,
                    int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::advection(nx, ny, fld, u, v, xmapr, ymapr, hours, advec, f, undef);
}

int fcref_gradient(int nx, int ny, const float* field, const float* xmapr, const float* ymapr, int compute, float* fgrad, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::gradient(nx, ny, field, xmapr, ymapr, compute, fgrad, f, undef);
}

int fcref_shapiro2_filter(int nx, int ny, float* field, float* fsmooth, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::shapiro2_filter(nx, ny, field, fsmooth, f, undef);
}

int fcref_windCooling(int nx, int ny, const float* t, const float* u, const float* v, int compute, float* dtcool, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::windCooling(nx, ny, t, u, v, compute, dtcool, f, undef);
}

int fcref_thermalFrontParameter(int nx, int ny, const float* t, const float* xmapr, const float* ymapr, float* tfp, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::thermalFrontParameter(nx, ny, t, xmapr, ymapr, tfp, f, undef);
}

int fcref_momentumXcoordinate(int nx, int ny, const float* v, const float* xmapr, const float* fcoriolis, float fcoriolisMin, float* mxy, int* fDefined,
                              float undef)
{
  Flag f(fDefined);
  return fc::momentumXcoordinate(nx, ny, v, xmapr, fcoriolis, fcoriolisMin, mxy, f, undef);
}

int fcref_momentumYcoordinate(int nx, int ny, const float* u, const float* ymapr, const float* fcoriolis, float fcoriolisMin, float* nxy, int* fDefined,
                              float undef)
{
  Flag f(fDefined);
  return fc::momentumYcoordinate(nx, ny, u, ymapr, fcoriolis, fcoriolisMin, nxy, f, undef);
}

int fcref_jacobian(int nx, int ny, const float* field1, const float* field2, const float* xmapr, const float* ymapr, float* fjacobian, int* fDefined,
                   float undef)
{
  Flag f(fDefined);
  return fc::jacobian(nx, ny, field1, field2, xmapr, ymapr, fjacobian, f, undef);
}

int fcref_vesselIcingOverland(int nx, int ny, const float* airtemp, const float* seatemp, const float* u, const float* v, const float* sal,
                              const float* aice, float* icing, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::vesselIcingOverland(nx, ny, airtemp, seatemp, u, v, sal, aice, icing, f, undef);
}

int fcref_vesselIcingMertins(int nx, int ny, const float* airtemp, const float* seatemp, const float* u, const float* v, const float* sal,
                             const float* aice, float* icing, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::vesselIcingMertins(nx, ny, airtemp, seatemp, u, v, sal, aice, icing, f, undef);
}

int fcref_vesselIcingModStall(int nx, int ny, const float* sal, const float* wave, const float* x_wind, const float* y_wind, const float* airtemp,
                              const float* rh, const float* sst, const float* p, const float* Pw, const float* aice, const float* depth, float vs,
                              float alpha, float zmin, float zmax, float* icing, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::vesselIcingModStall(nx, ny, sal, wave, x_wind, y_wind, airtemp, rh, sst, p, Pw, aice, depth, vs, alpha, zmin, zmax, icing, f, undef);
}

int fcref_vesselIcingMincog(int nx, int ny, const float* sal, const float* wave, const float* x_wind, const float* y_wind, const float* airtemp,
                            const float* rh, const float* sst, const float* p, const float* Pw, const float* aice, const float* depth, float vs,
                            float alpha, float zmin, float zmax, int alt, float* icing, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::vesselIcingMincog(nx, ny, sal, wave, x_wind, y_wind, airtemp, rh, sst, p, Pw, aice, depth, vs, alpha, zmin, zmax, alt, icing, f, undef);
}

int fcref_fieldOPERfield(int compute, int nx, int ny, const float* field1, const float* field2, float* fres, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::fieldOPERfield(compute, nx, ny, field1, field2, fres, f, undef);
}

int fcref_meanValue(int nx, int ny, const float* const* fields, int nfields, const int* fDefinedIn, float* fres, int* fDefinedOut, float undef)
{
  Flag f(fDefinedOut);
  return fc::meanValue(nx, ny, as_vector(fields, nfields), as_flags(fDefinedIn, nfields), fres, f, undef);
}

int fcref_stddevValue(int nx, int ny, const float* const* fields, int nfields, const int* fDefinedIn, float* fres, int* fDefinedOut, float undef)
{
  Flag f(fDefinedOut);
  return fc::stddevValue(nx, ny, as_vector(fields, nfields), as_flags(fDefinedIn, nfields), fres, f, undef);
}

int fcref_extremeValue(int compute, int nx, int ny, const float* const* fields, int nfields, float* fres, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::extremeValue(compute, nx, ny, as_vector(fields, nfields), fres, f, undef);
}

int fcref_probability(int compute, int nx, int ny, const float* const* fields, int nfields, const int* fDefinedIn, const float* limits, int nlimits,
                      float* fres, int* fDefinedOut, float undef)
{
  Flag f(fDefinedOut);
  // the reference reads limits[0] before validating the size (FieldCalculations.cc:2824);
  // keep one readable element so that the nlimits == 0 rejection path can be driven safely
  std::vector<float> lim(limits, limits + (nlimits > 0 ? nlimits : 0));
  lim.reserve(2);
  return fc::probability(compute, nx, ny, as_vector(fields, nfields), as_flags(fDefinedIn, nfields), lim, fres, f, undef);
}

// ---- the rest of the reference's Python subset
int fcref_kIndex(int nx, int ny, const float* t500, const float* t700, const float* rh700, const float* t850, const float* rh850, float p500, float p700,
                 float p850, int compute, float* kfield, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::kIndex(nx, ny, t500, t700, rh700, t850, rh850, p500, p700, p850, compute, kfield, f, undef);
}

int fcref_ductingIndex(int nx, int ny, const float* t850, const float* rh850, float p850, int compute, float* duct, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::ductingIndex(nx, ny, t850, rh850, p850, compute, duct, f, undef);
}

int fcref_showalterIndex(int nx, int ny, const float* t500, const float* t850, const float* rh850, float p500, float p850, int compute, float* sfield,
                         int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::showalterIndex(nx, ny, t500, t850, rh850, p500, p850, compute, sfield, f, undef);
}

int fcref_boydenIndex(int nx, int ny, const float* t700, const float* z700, const float* z1000, float p700, float p1000, int compute, float* bfield,
                      int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::boydenIndex(nx, ny, t700, z700, z1000, p700, p1000, compute, bfield, f, undef);
}

int fcref_sweatIndex(int nx, int ny, const float* t850, const float* t500, const float* td850, const float* td500, const float* u850, const float* v850,
                     const float* u500, const float* v500, float* sindex, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::sweatIndex(nx, ny, t850, t500, td850, td500, u850, v850, u500, v500, sindex, f, undef);
}

int fcref_seaSoundSpeed(int nx, int ny, const float* t, const float* s, float z, int compute, float* soundspeed, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::seaSoundSpeed(nx, ny, t, s, z, compute, soundspeed, f, undef);
}

int fcref_cvtemp(int nx, int ny, const float* tinp, int compute, float* tout, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::cvtemp(nx, ny, tinp, compute, tout, f, undef);
}

int fcref_cvhum(int nx, int ny, const float* t, const float* huminp, const char* unit, int compute, float* humout, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::cvhum(nx, ny, t, huminp, std::string(unit), compute, humout, f, undef);
}

int fcref_abshum(int nx, int ny, const float* t, const float* rhum, float* abshumout, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::abshum(nx, ny, t, rhum, abshumout, f, undef);
}

int fcref_underCooledRain(int nx, int ny, const float* precip, const float* snow, const float* tk, float precipMin, float snowRateMax, float tcMax,
                          float* undercooled, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::underCooledRain(nx, ny, precip, snow, tk, precipMin, snowRateMax, tcMax, undercooled, f, undef);
}

// ---- the rest of SURVEY.md 8f rank 1
int fcref_plevelthe(int nx, int ny, const float* t, const float* rh, float p, int compute, float* the, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::plevelthe(nx, ny, t, rh, p, compute, the, f, undef);
}
int fcref_pleveldz2tmean(int nx, int ny, const float* z1, const float* z2, float p1, float p2, int compute, float* tmean, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::pleveldz2tmean(nx, ny, z1, z2, p1, p2, compute, tmean, f, undef);
}
int fcref_plevelducting(int nx, int ny, const float* t, const float* h, float p, int compute, float* duct, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::plevelducting(nx, ny, t, h, p, compute, duct, f, undef);
}
int fcref_vectorabs(int nx, int ny, const float* u, const float* v, float* ff, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::vectorabs(nx, ny, u, v, ff, f, undef);
}
int fcref_pressure2FlightLevel(int nx, int ny, const float* pressure, float* flightlevel, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::pressure2FlightLevel(nx, ny, pressure, flightlevel, f, undef);
}
int fcref_values2classes(int nx, int ny, const float* fvalue, float* fclass, const float* values, int nvalues, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::values2classes(nx, ny, fvalue, fclass, std::vector<float>(values, values + (nvalues > 0 ? nvalues : 0)), f, undef);
}
#define FCREF_VOID2(name)                                                                                                                            \
  int fcref_##name(int nx, int ny, const float* field1, const float* field2, float* fres, int* fDefined, float undef)                               \
  {                                                                                                                                                  \
    Flag f(fDefined);                                                                                                                                \
    fc::name(nx, ny, field1, field2, fres, f, undef);                                                                                                \
    return 1;                                                                                                                                        \
  }
#define FCREF_VOID1C(name)                                                                                                                           \
  int fcref_##name(int nx, int ny, const float* field, float value, float* fres, int* fDefined, float undef)                                        \
  {                                                                                                                                                  \
    Flag f(fDefined);                                                                                                                                \
    fc::name(nx, ny, field, value, fres, f, undef);                                                                                                  \
    return 1;                                                                                                                                        \
  }
#define FCREF_VOID1(name)                                                                                                                            \
  int fcref_##name(int nx, int ny, const float* field, float* fres, int* fDefined, float undef)                                                      \
  {                                                                                                                                                  \
    Flag f(fDefined);                                                                                                                                \
    fc::name(nx, ny, field, fres, f, undef);                                                                                                         \
    return 1;                                                                                                                                        \
  }
FCREF_VOID2(minvalueFields)
FCREF_VOID2(maxvalueFields)
FCREF_VOID1C(minvalueFieldConst)
FCREF_VOID1C(maxvalueFieldConst)
FCREF_VOID1(absvalueField)
FCREF_VOID1(log10Field)
FCREF_VOID1(pow10Field)
FCREF_VOID1(logField)
FCREF_VOID1(expField)
FCREF_VOID1C(powerField)
FCREF_VOID1C(replaceUndefined)
FCREF_VOID1C(replaceDefined)
int fcref_fieldOPERconstant(int compute, int nx, int ny, const float* field, float value, float* fres, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::fieldOPERconstant(compute, nx, ny, field, value, fres, f, undef);
}
int fcref_constantOPERfield(int compute, int nx, int ny, float value, const float* field, float* fres, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::constantOPERfield(compute, nx, ny, value, field, fres, f, undef);
}
int fcref_sumFields(int nx, int ny, const float* const* fields, int nfields, float* fres, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::sumFields(nx, ny, as_vector(fields, nfields), fres, f, undef);
}
int fcref_snow_in_cm(int nx, int ny, const float* snow_water, const float* tk2m, const float* td2m, float* snow_cm, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::snow_in_cm(nx, ny, snow_water, tk2m, td2m, snow_cm, f, undef);
}

// ---- geostrophic stencil siblings
int fcref_plevelgwind_xcomp(int nx, int ny, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis, float* ug, int* fDefined,
                            float undef)
{
  Flag f(fDefined);
  return fc::plevelgwind_xcomp(nx, ny, z, xmapr, ymapr, fcoriolis, ug, f, undef);
}
int fcref_plevelgwind_ycomp(int nx, int ny, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis, float* vg, int* fDefined,
                            float undef)
{
  if (nx < 3 || ny < 3) // the reference has no such test and runs fillEdges out of bounds: never drive it there
    return 0;
  Flag f(fDefined);
  return fc::plevelgwind_ycomp(nx, ny, z, xmapr, ymapr, fcoriolis, vg, f, undef);
}
int fcref_plevelgvort(int nx, int ny, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis, float* gvort, int* fDefined,
                      float undef)
{
  Flag f(fDefined);
  return fc::plevelgvort(nx, ny, z, xmapr, ymapr, fcoriolis, gvort, f, undef);
}

int fcref_plevelqvector(int nx, int ny, const float* z, const float* t, const float* xmapr, const float* ymapr, const float* fcoriolis, float p, int compute,
                        float* qcomp, int* fDefined, float undef)
{
  Flag f(fDefined);
  return fc::plevelqvector(nx, ny, z, t, xmapr, ymapr, fcoriolis, p, compute, qcomp, f, undef);
}

// ---- neighbourhood functions; parameter combinations for which the reference indexes out of
// bounds are never driven into it (same rejections as the restatement)
int fcref_neighbourProbFunctions(int nx, int ny, const float* field, const float* constants, int nconstants, int compute, float* fres, int* fDefined,
                                 float undef)
{
  if (nconstants >= 2) {
    const int range = (int)constants[1];
    if (range < 0 || range > nx || range > ny)
      return 0;
  }
  Flag f(fDefined);
  return fc::neighbourProbFunctions(nx, ny, field, std::vector<float>(constants, constants + (nconstants > 0 ? nconstants : 0)), compute, fres, f, undef);
}
int fcref_neighbourFunctions(int nx, int ny, const float* field, const float* constants, int nconstants, int compute, float* fres, int* fDefined,
                             float undef)
{
  if (!(nconstants < 1 || (nconstants < 2 && compute > 3))) {
    int range = 3, step = 3, limit = 0;
    if (compute < 4) {
      range = (int)constants[0];
      if (nconstants == 2)
        step = (int)constants[1];
    } else {
      limit = (int)constants[0];
      range = (int)constants[1];
      if (nconstants == 3)
        step = (int)constants[2];
    }
    const int nwin = (2 * range + 1) * (2 * range + 1);
    const int ii = (int)((float)nwin * limit / 100);
    if (range >= 1 && step >= 1 && (step / 2 > range || (compute == 4 && (ii < 0 || ii >= nwin))))
      return 0;
  }
  Flag f(fDefined);
  return fc::neighbourFunctions(nx, ny, field, std::vector<float>(constants, constants + (nconstants > 0 ? nconstants : 0)), compute, fres, f, undef);
}

} // extern "C"
