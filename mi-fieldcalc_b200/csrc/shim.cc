// shim.cc -- libmi-fieldcalc.so.0 drop-in: the reference's C++ free-function API
// (namespace miutil::fieldcalc, src/mi_fieldcalc/FieldCalculations.h:113-303 of the reference) on top
// of the C-ABI of include/fcb200.h.  Nothing is computed here: every function converts
// `ValuesDefined&` <-> `int*`, `std::string` -> `const char*`, `std::vector` -> pointer + count and
// forwards to fcb200_<name>.  The exported (mangled) symbols are exactly the reference's, so an
// application linked against mi-fieldcalc picks this library up unchanged.
//
// Error policy: the reference cannot fail at run time.  A runtime failure here (no CUDA device,
// CUDA error) must not look like "arguments rejected", so it throws std::runtime_error; with
// FCB200_ON_ERROR=return in the environment it prints the message and returns false instead.
#include "mi_fieldcalc/FieldCalculations.h"

#include "mi_fieldcalc/MetConstants.h"

#include "fcb200.h"

#include <cmath>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

const float fieldUndef = 1.0e35f;

namespace miutil {

const float UNDEF = 1.0e35f;

ValuesDefined checkDefined(const float* data, size_t n)
{ // reference FieldDefined.cc:36-60: defined means `value < UNDEF` (NaN is therefore undefined)
  bool some_defined = false, some_undefined = false;
  for (size_t i = 0; i < n && !(some_defined && some_undefined); ++i) {
    if (data[i] < UNDEF)
      some_defined = true;
    else
      some_undefined = true;
  }
  if (some_defined && some_undefined)
    return SOME_DEFINED;
  return some_defined ? ALL_DEFINED : NONE_DEFINED;
}

ValuesDefined checkDefined(size_t n_undefined, size_t n)
{ // reference FieldDefined.cc:62-70
  if (n_undefined == 0)
    return ALL_DEFINED;
  return (n_undefined == n) ? NONE_DEFINED : SOME_DEFINED;
}

ValuesDefined combineDefined(ValuesDefined a, ValuesDefined b)
{ // reference FieldDefined.cc:72-83
  if (a == ALL_DEFINED)
    return b;
  if (a == NONE_DEFINED)
    return NONE_DEFINED;
  return (b != ALL_DEFINED) ? b : SOME_DEFINED;
}

// openmp_tools.h:58 of the reference: the thread count of its OpenMP loops.  There are no host loops here: the serial
// build's answer (openmp_tools.cc:78).
int compute_num_threads(long) { return 1; }

// ---- OUT OF SCOPE, ABI FILLER: host scalars that the drop-in library must export so that `nm -D`
// matches the reference's (tests/test_cpp_api.py); nothing on the GPU path calls them and they are not counted as product work.
// MetConstants.cc:51-131 of the reference: ICAO standard atmosphere, seven layers of constant lapse rate up to
// 84.852 km.  Layer k starts at height H[k] (km) with temperature T[k] (K) and pressure P[k] (hPa) and has the
// temperature gradient L[k] (K/km).  Inside a layer: p/P = (1 + dh L/T)^(-g/(L R)) for L != 0, exp(-dh g/(R T)) for L = 0.
namespace constants {
namespace {
const double kG = 9.80665, kR = 287.05287;
const int kLayers = 8;
const double kL[kLayers - 1] = {-6.5, 0, +1.0, +2.8, 0, -2.8, -2.0};
const double kH[kLayers] = {0, 11, 20, 32, 47, 51, 71, 84.852};
const double kT[kLayers] = {288.15, 216.65, 216.65, 228.65, 270.65, 270.65, 214.65, 186.946};
const double kP[kLayers] = {1013.15,           226.29806486313493, 54.743370958898005,  8.679301101236328,
                            1.1089482781849516, 0.6693192180209551, 0.0395600169484907, 0.0037334345211142398};
} // namespace

double ICAO_geo_altitude_from_pressure(double pressure)
{
  int k = 1;
  while (k < kLayers && pressure < kP[k])
    ++k;
  if (k >= kLayers)
    return 1000 * (kH[kLayers - 1] + 1); // above the standard atmosphere
  --k;
  const double lapse = kL[k] / 1000, base = kH[k] * 1000, ratio = pressure / kP[k];
  if (lapse != 0)
    return (kT[k] / lapse) * (std::pow(ratio, -(lapse * kR) / kG) - 1) + base;
  return base - std::log(ratio) * (kR * kT[k]) / kG;
}

double ICAO_pressure_from_geo_altitude(double altitude)
{
  const double km = altitude / 1000;
  int k = 1;
  while (k < kLayers && km > kH[k])
    ++k;
  if (k >= kLayers)
    return kP[kLayers - 1] - 1; // above the standard atmosphere
  --k;
  const double lapse = kL[k] / 1000, dh = altitude - kH[k] * 1000;
  const double factor = (lapse != 0) ? std::pow(1 + dh * lapse / kT[k], -kG / (lapse * kR)) : std::exp(-dh * kG / (kR * kT[k]));
  return kP[k] * factor;
}

int FL_from_geo_altitude(double a) { return 5 * (int)round(a * ft_per_m / 500); }

double geo_altitude_from_FL(double fl) { return fl * 100 / ft_per_m; }

} // namespace constants

namespace fieldcalc {

// FC.cc:298-301 (an exported helper of the reference, although not declared in its header)
float bad_hlevel(float a, float b) { return (a < 0.0) || (b < 0.0) || (a == 0.0 && b == 0.0) || (b > 1.0); }

namespace {

// result of a C-ABI call -> the reference's bool, flag copied back
bool done(int rc, const int& flag, ValuesDefined& fDefined) // `flag` by reference: it is written by the call that produces `rc`
{
  if (rc < 0) {
    const char* msg = fcb200_last_error();
    const char* mode = std::getenv("FCB200_ON_ERROR");
    if (mode && std::strcmp(mode, "return") == 0) {
      std::fprintf(stderr, "mi-fieldcalc (B200): %s\n", msg && *msg ? msg : "runtime error");
      return false;
    }
    throw std::runtime_error(msg && *msg ? msg : "mi-fieldcalc (B200): runtime error");
  }
  // Deferred mode (fcb200_begin_deferred, thread-local state shared with libfcb200) applies to the C-ABI only, where the
  // caller owns the flag storage.  Here the flag lives in the calling wrapper's stack frame and the reference's contract
  // is "output and flag are final on return": drain now, while `flag` is alive, and stay in deferred mode for the
  // caller's own fcb200_* calls.
  if (rc >= 0 && fcb200_in_deferred()) {
    if (fcb200_end_deferred() < 0) {
      fcb200_begin_deferred();
      return done(-1, flag, fDefined);
    }
    fcb200_begin_deferred();
  }
  fDefined = static_cast<ValuesDefined>(flag);
  return rc == 1;
}

std::vector<int> to_ints(const std::vector<ValuesDefined>& v)
{
  std::vector<int> r(v.size());
  for (size_t i = 0; i < v.size(); ++i)
    r[i] = static_cast<int>(v[i]);
  return r;
}

} // namespace

void copy_field(float* fout, const float* fin, size_t fsize)
{ // host helper of the reference (FieldCalculations.cc:318-322)
  if (fout != fin)
    std::memcpy(fout, fin, sizeof(float) * fsize);
}

#define FCB_FLAG int f = static_cast<int>(fDefined)

bool pleveltemp(int nx, int ny, const float* tinp, float p, const std::string& unit, int compute, float* tout, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_pleveltemp(nx, ny, tinp, p, unit.c_str(), compute, tout, &f, undef), f, fDefined);
}

bool plevelhum(int nx, int ny, const float* t, const float* huminp, float p, const std::string& unit, int compute, float* humout,
               ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_plevelhum(nx, ny, t, huminp, p, unit.c_str(), compute, humout, &f, undef), f, fDefined);
}

bool hleveltemp(int nx, int ny, const float* tinp, const float* ps, float alevel, float blevel, const std::string& unit, int compute, float* tout,
                ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_hleveltemp(nx, ny, tinp, ps, alevel, blevel, unit.c_str(), compute, tout, &f, undef), f, fDefined);
}

bool hlevelthe(int nx, int ny, const float* t, const float* q, const float* ps, float alevel, float blevel, int compute, float* the,
               ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_hlevelthe(nx, ny, t, q, ps, alevel, blevel, compute, the, &f, undef), f, fDefined);
}

bool hlevelhum(int nx, int ny, const float* t, const float* huminp, const float* ps, float alevel, float blevel, const std::string& unit, int compute,
               float* humout, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_hlevelhum(nx, ny, t, huminp, ps, alevel, blevel, unit.c_str(), compute, humout, &f, undef), f, fDefined);
}

bool hlevelducting(int nx, int ny, const float* t, const float* h, const float* ps, float alevel, float blevel, int compute, float* duct,
                   ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_hlevelducting(nx, ny, t, h, ps, alevel, blevel, compute, duct, &f, undef), f, fDefined);
}

bool hlevelpressure(int nx, int ny, const float* ps, float alevel, float blevel, float* p, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_hlevelpressure(nx, ny, ps, alevel, blevel, p, &f, undef), f, fDefined);
}

bool aleveltemp(int nx, int ny, const float* tinp, const float* p, const std::string& unit, int compute, float* tout, ValuesDefined& fDefined,
                float undef)
{
  FCB_FLAG;
  return done(fcb200_aleveltemp(nx, ny, tinp, p, unit.c_str(), compute, tout, &f, undef), f, fDefined);
}

bool alevelthe(int nx, int ny, const float* t, const float* q, const float* p, int compute, float* the, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_alevelthe(nx, ny, t, q, p, compute, the, &f, undef), f, fDefined);
}

bool alevelhum(int nx, int ny, const float* t, const float* huminp, const float* p, const std::string& unit, int compute, float* humout,
               ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_alevelhum(nx, ny, t, huminp, p, unit.c_str(), compute, humout, &f, undef), f, fDefined);
}

bool alevelducting(int nx, int ny, const float* t, const float* h, const float* p, int compute, float* duct, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_alevelducting(nx, ny, t, h, p, compute, duct, &f, undef), f, fDefined);
}

bool ilevelgwind(int nx, int ny, const float* mpot, const float* xmapr, const float* ymapr, const float* fcoriolis, float* ug, float* vg,
                 ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_ilevelgwind(nx, ny, mpot, xmapr, ymapr, fcoriolis, ug, vg, &f, undef), f, fDefined);
}

bool relvort(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, float* rvort, ValuesDefined& fDefined,
             float undef)
{
  FCB_FLAG;
  return done(fcb200_relvort(nx, ny, u, v, xmapr, ymapr, rvort, &f, undef), f, fDefined);
}

bool absvort(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, const float* fcoriolis, float* avort,
             ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_absvort(nx, ny, u, v, xmapr, ymapr, fcoriolis, avort, &f, undef), f, fDefined);
}

bool divergence(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, float* diverg, ValuesDefined& fDefined,
                float undef)
{
  FCB_FLAG;
  return done(fcb200_divergence(nx, ny, u, v, xmapr, ymapr, diverg, &f, undef), f, fDefined);
}

bool advection(int nx, int ny, const float* fld, const float* u, const float* v, const float* xmapr, const float* ymapr, float hours, float* advec,
               ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_advection(nx, ny, fld, u, v, xmapr, ymapr, hours, advec, &f, undef), f, fDefined);
}

bool gradient(int nx, int ny, const float* field, const float* xmapr, const float* ymapr, int compute, float* fgrad, ValuesDefined& fDefined,
              float undef)
{
  FCB_FLAG;
  return done(fcb200_gradient(nx, ny, field, xmapr, ymapr, compute, fgrad, &f, undef), f, fDefined);
}

bool shapiro2_filter(int nx, int ny, float* field, float* fsmooth, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_shapiro2_filter(nx, ny, field, fsmooth, &f, undef), f, fDefined);
}

bool kIndex(int nx, int ny, const float* t500, const float* t700, const float* rh700, const float* t850, const float* rh850, float p500, float p700,
            float p850, int compute, float* kfield, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_kIndex(nx, ny, t500, t700, rh700, t850, rh850, p500, p700, p850, compute, kfield, &f, undef), f, fDefined);
}

bool ductingIndex(int nx, int ny, const float* t850, const float* rh850, float p850, int compute, float* duct, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_ductingIndex(nx, ny, t850, rh850, p850, compute, duct, &f, undef), f, fDefined);
}

bool showalterIndex(int nx, int ny, const float* t500, const float* t850, const float* rh850, float p500, float p850, int compute, float* sfield,
                    ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_showalterIndex(nx, ny, t500, t850, rh850, p500, p850, compute, sfield, &f, undef), f, fDefined);
}

bool boydenIndex(int nx, int ny, const float* t700, const float* z700, const float* z1000, float p700, float p1000, int compute, float* bfield,
                 ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_boydenIndex(nx, ny, t700, z700, z1000, p700, p1000, compute, bfield, &f, undef), f, fDefined);
}

bool sweatIndex(int nx, int ny, const float* t850, const float* t500, const float* td850, const float* td500, const float* u850, const float* v850,
                const float* u500, const float* v500, float* sindex, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_sweatIndex(nx, ny, t850, t500, td850, td500, u850, v850, u500, v500, sindex, &f, undef), f, fDefined);
}

bool seaSoundSpeed(int nx, int ny, const float* t, const float* s, float z, int compute, float* soundspeed, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_seaSoundSpeed(nx, ny, t, s, z, compute, soundspeed, &f, undef), f, fDefined);
}

bool cvtemp(int nx, int ny, const float* tinp, int compute, float* tout, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_cvtemp(nx, ny, tinp, compute, tout, &f, undef), f, fDefined);
}

bool cvhum(int nx, int ny, const float* t, const float* huminp, const std::string& unit, int compute, float* humout, ValuesDefined& fDefined,
           float undef)
{
  FCB_FLAG;
  return done(fcb200_cvhum(nx, ny, t, huminp, unit.c_str(), compute, humout, &f, undef), f, fDefined);
}

bool abshum(int nx, int ny, const float* t, const float* rhum, float* abshumout, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_abshum(nx, ny, t, rhum, abshumout, &f, undef), f, fDefined);
}

bool underCooledRain(int nx, int ny, const float* precip, const float* snow, const float* tk, float precipMin, float snowRateMax, float tcMax,
                     float* undercooled, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_underCooledRain(nx, ny, precip, snow, tk, precipMin, snowRateMax, tcMax, undercooled, &f, undef), f, fDefined);
}

// ---- pressure-level siblings, element functions, field arithmetic
bool plevelthe(int nx, int ny, const float* t, const float* rh, float p, int compute, float* the, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_plevelthe(nx, ny, t, rh, p, compute, the, &f, undef), f, fDefined);
}

bool pleveldz2tmean(int nx, int ny, const float* z1, const float* z2, float p1, float p2, int compute, float* tmean, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_pleveldz2tmean(nx, ny, z1, z2, p1, p2, compute, tmean, &f, undef), f, fDefined);
}

bool plevelducting(int nx, int ny, const float* t, const float* h, float p, int compute, float* duct, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_plevelducting(nx, ny, t, h, p, compute, duct, &f, undef), f, fDefined);
}

bool vectorabs(int nx, int ny, const float* u, const float* v, float* ff, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_vectorabs(nx, ny, u, v, ff, &f, undef), f, fDefined);
}

bool pressure2FlightLevel(int nx, int ny, const float* pressure, float* flightlevel, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_pressure2FlightLevel(nx, ny, pressure, flightlevel, &f, undef), f, fDefined);
}

bool values2classes(int nx, int ny, const float* fvalue, float* fclass, const std::vector<float>& values, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_values2classes(nx, ny, fvalue, fclass, values.data(), (int)values.size(), &f, undef), f, fDefined);
}

#define FCB_SHIM_VOID2(name)                                                                                                                         \
  void name(int nx, int ny, const float* field1, const float* field2, float* fres, ValuesDefined& fDefined, float undef)                             \
  {                                                                                                                                                  \
    FCB_FLAG;                                                                                                                                        \
    done(fcb200_##name(nx, ny, field1, field2, fres, &f, undef), f, fDefined);                                                                       \
  }
#define FCB_SHIM_VOID1C(name)                                                                                                                        \
  void name(int nx, int ny, const float* field, const float value, float* fres, ValuesDefined& fDefined, float undef)                                \
  {                                                                                                                                                  \
    FCB_FLAG;                                                                                                                                        \
    done(fcb200_##name(nx, ny, field, value, fres, &f, undef), f, fDefined);                                                                         \
  }
#define FCB_SHIM_VOID1(name)                                                                                                                         \
  void name(int nx, int ny, const float* field, float* fres, ValuesDefined& fDefined, float undef)                                                   \
  {                                                                                                                                                  \
    FCB_FLAG;                                                                                                                                        \
    done(fcb200_##name(nx, ny, field, fres, &f, undef), f, fDefined);                                                                                \
  }
FCB_SHIM_VOID2(minvalueFields)
FCB_SHIM_VOID1C(minvalueFieldConst)
FCB_SHIM_VOID2(maxvalueFields)
FCB_SHIM_VOID1C(maxvalueFieldConst)
FCB_SHIM_VOID1(absvalueField)
FCB_SHIM_VOID1(log10Field)
FCB_SHIM_VOID1(pow10Field)
FCB_SHIM_VOID1(logField)
FCB_SHIM_VOID1(expField)
FCB_SHIM_VOID1C(powerField)
FCB_SHIM_VOID1C(replaceUndefined)
FCB_SHIM_VOID1C(replaceDefined)

bool fieldOPERconstant(int compute, int nx, int ny, const float* field, float value, float* fres, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_fieldOPERconstant(compute, nx, ny, field, value, fres, &f, undef), f, fDefined);
}

bool constantOPERfield(int compute, int nx, int ny, float value, const float* field, float* fres, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_constantOPERfield(compute, nx, ny, value, field, fres, &f, undef), f, fDefined);
}

bool sumFields(int nx, int ny, const std::vector<float*>& fields, float* fres, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_sumFields(nx, ny, fields.data(), (int)fields.size(), fres, &f, undef), f, fDefined);
}

bool snow_in_cm(int nx, int ny, const float* snow_water, const float* tk2m, const float* td2m, float* snow_cm, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_snow_in_cm(nx, ny, snow_water, tk2m, td2m, snow_cm, &f, undef), f, fDefined);
}

bool plevelgwind_xcomp(int nx, int ny, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis, float* ug, ValuesDefined& fDefined,
                       float undef)
{
  FCB_FLAG;
  return done(fcb200_plevelgwind_xcomp(nx, ny, z, xmapr, ymapr, fcoriolis, ug, &f, undef), f, fDefined);
}

bool plevelgwind_ycomp(int nx, int ny, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis, float* vg, ValuesDefined& fDefined,
                       float undef)
{
  FCB_FLAG;
  return done(fcb200_plevelgwind_ycomp(nx, ny, z, xmapr, ymapr, fcoriolis, vg, &f, undef), f, fDefined);
}

bool plevelgvort(int nx, int ny, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis, float* gvort, ValuesDefined& fDefined,
                 float undef)
{
  FCB_FLAG;
  return done(fcb200_plevelgvort(nx, ny, z, xmapr, ymapr, fcoriolis, gvort, &f, undef), f, fDefined);
}

bool plevelqvector(int nx, int ny, const float* z, const float* t, const float* xmapr, const float* ymapr, const float* fcoriolis, float p, int compute,
                   float* qcomp, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_plevelqvector(nx, ny, z, t, xmapr, ymapr, fcoriolis, p, compute, qcomp, &f, undef), f, fDefined);
}

bool neighbourProbFunctions(int nx, int ny, const float* field, const std::vector<float>& constants, int compute, float* fres, ValuesDefined& fDefined,
                            float undef)
{
  FCB_FLAG;
  return done(fcb200_neighbourProbFunctions(nx, ny, field, constants.data(), (int)constants.size(), compute, fres, &f, undef), f, fDefined);
}

bool neighbourFunctions(int nx, int ny, const float* field, const std::vector<float>& constants, int compute, float* fres, ValuesDefined& fDefined,
                        float undef)
{
  FCB_FLAG;
  return done(fcb200_neighbourFunctions(nx, ny, field, constants.data(), (int)constants.size(), compute, fres, &f, undef), f, fDefined);
}

bool windCooling(int nx, int ny, const float* t, const float* u, const float* v, int compute, float* dtcool, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_windCooling(nx, ny, t, u, v, compute, dtcool, &f, undef), f, fDefined);
}

bool thermalFrontParameter(int nx, int ny, const float* t, const float* xmapr, const float* ymapr, float* tfp, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_thermalFrontParameter(nx, ny, t, xmapr, ymapr, tfp, &f, undef), f, fDefined);
}

bool momentumXcoordinate(int nx, int ny, const float* v, const float* xmapr, const float* fcoriolis, float fcoriolisMin, float* mxy,
                         ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_momentumXcoordinate(nx, ny, v, xmapr, fcoriolis, fcoriolisMin, mxy, &f, undef), f, fDefined);
}

bool momentumYcoordinate(int nx, int ny, const float* u, const float* ymapr, const float* fcoriolis, float fcoriolisMin, float* nxy,
                         ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_momentumYcoordinate(nx, ny, u, ymapr, fcoriolis, fcoriolisMin, nxy, &f, undef), f, fDefined);
}

bool jacobian(int nx, int ny, const float* field1, const float* field2, const float* xmapr, const float* ymapr, float* fjacobian,
              ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_jacobian(nx, ny, field1, field2, xmapr, ymapr, fjacobian, &f, undef), f, fDefined);
}

bool vesselIcingOverland(int nx, int ny, const float* airtemp, const float* seatemp, const float* u, const float* v, const float* sal,
                         const float* aice, float* icing, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_vesselIcingOverland(nx, ny, airtemp, seatemp, u, v, sal, aice, icing, &f, undef), f, fDefined);
}

bool vesselIcingMertins(int nx, int ny, const float* airtemp, const float* seatemp, const float* u, const float* v, const float* sal,
                        const float* aice, float* icing, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_vesselIcingMertins(nx, ny, airtemp, seatemp, u, v, sal, aice, icing, &f, undef), f, fDefined);
}

bool vesselIcingModStall(int nx, int ny, const float* sal, const float* wave, const float* x_wind, const float* y_wind, const float* airtemp,
                         const float* rh, const float* sst, const float* p, const float* Pw, const float* aice, const float* depth, const float vs,
                         const float alpha, const float zmin, const float zmax, float* icing, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_vesselIcingModStall(nx, ny, sal, wave, x_wind, y_wind, airtemp, rh, sst, p, Pw, aice, depth, vs, alpha, zmin, zmax, icing, &f, undef),
              f, fDefined);
}

bool vesselIcingMincog(int nx, int ny, const float* sal, const float* wave, const float* x_wind, const float* y_wind, const float* airtemp,
                       const float* rh, const float* sst, const float* p, const float* Pw, const float* aice, const float* depth, const float vs,
                       const float alpha, const float zmin, const float zmax, const int alt, float* icing, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(
      fcb200_vesselIcingMincog(nx, ny, sal, wave, x_wind, y_wind, airtemp, rh, sst, p, Pw, aice, depth, vs, alpha, zmin, zmax, alt, icing, &f, undef), f,
      fDefined);
}

bool fieldOPERfield(int compute, int nx, int ny, const float* field1, const float* field2, float* fres, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_fieldOPERfield(compute, nx, ny, field1, field2, fres, &f, undef), f, fDefined);
}

bool meanValue(int nx, int ny, const std::vector<float*>& fields, const std::vector<ValuesDefined>& fDefinedIn, float* fres,
               ValuesDefined& fDefinedOut, float undef)
{
  int f = static_cast<int>(fDefinedOut);
  const std::vector<int> fin = to_ints(fDefinedIn);
  return done(fcb200_meanValue(nx, ny, fields.data(), (int)fields.size(), fin.data(), fres, &f, undef), f, fDefinedOut);
}

bool stddevValue(int nx, int ny, const std::vector<float*>& fields, const std::vector<ValuesDefined>& fDefinedIn, float* fres,
                 ValuesDefined& fDefinedOut, float undef)
{
  int f = static_cast<int>(fDefinedOut);
  const std::vector<int> fin = to_ints(fDefinedIn);
  return done(fcb200_stddevValue(nx, ny, fields.data(), (int)fields.size(), fin.data(), fres, &f, undef), f, fDefinedOut);
}

bool extremeValue(int compute, int nx, int ny, const std::vector<float*>& fields, float* fres, ValuesDefined& fDefined, float undef)
{
  FCB_FLAG;
  return done(fcb200_extremeValue(compute, nx, ny, fields.data(), (int)fields.size(), fres, &f, undef), f, fDefined);
}

bool probability(int compute, int nx, int ny, const std::vector<float*>& fields, const std::vector<ValuesDefined>& fDefinedIn,
                 const std::vector<float>& limits, float* fres, ValuesDefined& fDefinedOut, float undef)
{
  int f = static_cast<int>(fDefinedOut);
  const std::vector<int> fin = to_ints(fDefinedIn);
  return done(fcb200_probability(compute, nx, ny, fields.data(), (int)fields.size(), fin.data(), limits.data(), (int)limits.size(), fres, &f, undef), f,
              fDefinedOut);
}

} // namespace fieldcalc
} // namespace miutil
