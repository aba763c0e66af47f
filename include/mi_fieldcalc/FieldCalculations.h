/* mi_fieldcalc/FieldCalculations.h -- drop-in replacement header (B200 build).
 *
 * Declares the hot path of the reference's src/mi_fieldcalc/FieldCalculations.h with IDENTICAL
 * signatures (namespace miutil::fieldcalc, argument order nx, ny, inputs, scalars with `compute`
 * last, output, fDefined, undef -- reference FieldCalculations.h:102-107), implemented by
 * mi-fieldcalc_b200/csrc/shim.cc on top of the C-ABI include/fcb200.h.  The line of the reference
 * declaration is given as "ref:<line>".
 *
 * Differences a caller can observe:
 *   - every field pointer may be host OR device (CUDA) memory;
 *   - a runtime failure (no GPU, CUDA error) throws std::runtime_error -- the reference cannot fail
 *     at run time; set FCB200_ON_ERROR=return to get `false` plus a message on stderr instead;
 *   - functions of the reference that are not declared here are not part of this build (linking
 *     fails loudly instead of silently running something else).
 */
#ifndef MI_FIELDCALC_FIELDCALCULATIONS_H
#define MI_FIELDCALC_FIELDCALCULATIONS_H

#include "FieldDefined.h"

#include <cmath>
#include <cstddef>
#include <string>
#include <vector>

namespace miutil {
namespace fieldcalc {

/* defined = not NaN and not the undefined value (ref:42-45) */
inline bool is_defined(float in, float undef)
{
  return !std::isnan(in) && in != undef;
}

/* is_defined(allDefined, v1, ..., vk, undef): true if allDefined, or if every v is defined.
 * One variadic template stands in for the reference's ten fixed-arity overloads (ref:47-98). */
namespace detail {
template <typename Last>
inline float last_of(Last last)
{
  return last;
}
template <typename First, typename... Rest>
inline float last_of(First, Rest... rest)
{
  return last_of(rest...);
}
} // namespace detail

inline bool is_defined(bool allDefined, float in1, float undef)
{
  return allDefined || is_defined(in1, undef);
}
inline bool is_defined(bool allDefined, float in1, float in2, float undef)
{
  return allDefined || (is_defined(in1, undef) && is_defined(in2, undef));
}
template <typename... F>
inline bool is_defined(bool allDefined, float in1, float in2, float in3, F... more_then_undef)
{
  // the last argument is the undefined value; every other one is a field value
  const float undef = detail::last_of(more_then_undef...);
  const float vals[] = {in1, in2, in3, static_cast<float>(more_then_undef)...};
  if (allDefined)
    return true;
  for (std::size_t k = 0; k + 1 < sizeof(vals) / sizeof(vals[0]); ++k)
    if (!is_defined(vals[k], undef))
      return false;
  return true;
}

void copy_field(float* fout, const float* fin, size_t fsize); /* ref:100 */

/* ---- pressure levels: scalar p (hPa) ---- */
bool pleveltemp(int nx, int ny, const float* tinp, float p, const std::string& unit, int compute, float* tout, ValuesDefined& fDefined,
                float undef); /* ref:113 */
bool plevelhum(int nx, int ny, const float* t, const float* huminp, float p, const std::string& unit, int compute, float* humout,
               ValuesDefined& fDefined, float undef); /* ref:117 */

/* ---- hybrid model levels: p = alevel + blevel * ps ---- */
bool hleveltemp(int nx, int ny, const float* tinp, const float* ps, float alevel, float blevel, const std::string& unit, int compute, float* tout,
                ValuesDefined& fDefined, float undef); /* ref:154 */
bool hlevelthe(int nx, int ny, const float* t, const float* q, const float* ps, float alevel, float blevel, int compute, float* the,
               ValuesDefined& fDefined, float undef); /* ref:157 */
bool hlevelhum(int nx, int ny, const float* t, const float* huminp, const float* ps, float alevel, float blevel, const std::string& unit, int compute,
               float* humout, ValuesDefined& fDefined, float undef); /* ref:160 */
bool hlevelducting(int nx, int ny, const float* t, const float* h, const float* ps, float alevel, float blevel, int compute, float* duct,
                   ValuesDefined& fDefined, float undef); /* ref:163 */
bool hlevelpressure(int nx, int ny, const float* ps, float alevel, float blevel, float* p, ValuesDefined& fDefined, float undef); /* ref:166 */

/* ---- atmospheric model levels: p is a field ---- */
bool aleveltemp(int nx, int ny, const float* tinp, const float* p, const std::string& unit, int compute, float* tout, ValuesDefined& fDefined,
                float undef); /* ref:172 */
bool alevelthe(int nx, int ny, const float* t, const float* q, const float* p, int compute, float* the, ValuesDefined& fDefined,
               float undef); /* ref:174 */
bool alevelhum(int nx, int ny, const float* t, const float* huminp, const float* p, const std::string& unit, int compute, float* humout,
               ValuesDefined& fDefined, float undef); /* ref:176 */
bool alevelducting(int nx, int ny, const float* t, const float* h, const float* p, int compute, float* duct, ValuesDefined& fDefined,
                   float undef); /* ref:179 */

/* ---- isentropic level ---- */
bool ilevelgwind(int nx, int ny, const float* mpot, const float* xmapr, const float* ymapr, const float* fcoriolis, float* ug, float* vg,
                 ValuesDefined& fDefined, float undef); /* ref:185 */

/* ---- level independent: five-point map-ratio stencils, smoother, wind chill ---- */
bool relvort(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, float* rvort, ValuesDefined& fDefined,
             float undef); /* ref:206 */
bool absvort(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, const float* fcoriolis, float* avort,
             ValuesDefined& fDefined, float undef); /* ref:208 */
bool divergence(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, float* diverg, ValuesDefined& fDefined,
                float undef); /* ref:211 */
bool advection(int nx, int ny, const float* f, const float* u, const float* v, const float* xmapr, const float* ymapr, float hours, float* advec,
               ValuesDefined& fDefined, float undef); /* ref:213 */
bool gradient(int nx, int ny, const float* field, const float* xmapr, const float* ymapr, int compute, float* fgrad, ValuesDefined& fDefined,
              float undef); /* ref:216 */
bool shapiro2_filter(int nx, int ny, float* field, float* fsmooth, ValuesDefined& fDefined, float undef); /* ref:218 */
bool windCooling(int nx, int ny, const float* t, const float* u, const float* v, int compute, float* dtcool, ValuesDefined& fDefined,
                 float undef); /* ref:220 */
bool thermalFrontParameter(int nx, int ny, const float* t, const float* xmapr, const float* ymapr, float* tfp, ValuesDefined& fDefined,
                           float undef); /* ref:225 */
bool momentumXcoordinate(int nx, int ny, const float* v, const float* xmapr, const float* fcoriolis, float fcoriolisMin, float* mxy,
                         ValuesDefined& fDefined, float undef); /* ref:229 */
bool momentumYcoordinate(int nx, int ny, const float* u, const float* ymapr, const float* fcoriolis, float fcoriolisMin, float* nxy,
                         ValuesDefined& fDefined, float undef); /* ref:232 */
bool jacobian(int nx, int ny, const float* field1, const float* field2, const float* xmapr, const float* ymapr, float* fjacobian,
              ValuesDefined& fDefined, float undef); /* ref:235 */

/* ---- vessel icing (all temperatures in degrees Celsius) ---- */
bool vesselIcingOverland(int nx, int ny, const float* airtemp, const float* seatemp, const float* u, const float* v, const float* sal,
                         const float* aice, float* icing, ValuesDefined& fDefined, float undef); /* ref:238 */
bool vesselIcingMertins(int nx, int ny, const float* airtemp, const float* seatemp, const float* u, const float* v, const float* sal,
                        const float* aice, float* icing, ValuesDefined& fDefined, float undef); /* ref:241 */
bool vesselIcingModStall(int nx, int ny, const float* sal, const float* wave, const float* x_wind, const float* y_wind, const float* airtemp,
                         const float* rh, const float* sst, const float* p, const float* Pw, const float* aice, const float* depth, const float vs,
                         const float alpha, const float zmin, const float zmax, float* icing, ValuesDefined& fDefined, float undef); /* ref:244 */
bool vesselIcingMincog(int nx, int ny, const float* sal, const float* wave, const float* x_wind, const float* y_wind, const float* airtemp,
                       const float* rh, const float* sst, const float* p, const float* Pw, const float* aice, const float* depth, const float vs,
                       const float alpha, const float zmin, const float zmax, const int alt, float* icing, ValuesDefined& fDefined,
                       float undef); /* ref:248 */

/* ---- fixed-level stability indices and level-independent conversions (the rest of the Python subset) ---- */
bool kIndex(int nx, int ny, const float* t500, const float* t700, const float* rh700, const float* t850, const float* rh850, float p500,
            float p700, float p850, int compute, float* kfield, ValuesDefined& fDefined, float undef); /* ref:136 */
bool ductingIndex(int nx, int ny, const float* t850, const float* rh850, float p850, int compute, float* duct, ValuesDefined& fDefined,
                  float undef); /* ref:139 */
bool showalterIndex(int nx, int ny, const float* t500, const float* t850, const float* rh850, float p500, float p850, int compute,
                    float* sfield, ValuesDefined& fDefined, float undef); /* ref:141 */
bool boydenIndex(int nx, int ny, const float* t700, const float* z700, const float* z1000, float p700, float p1000, int compute,
                 float* bfield, ValuesDefined& fDefined, float undef); /* ref:144 */
bool sweatIndex(int nx, int ny, const float* t850, const float* t500, const float* td850, const float* td500, const float* u850,
                const float* v850, const float* u500, const float* v500, float* sindex, ValuesDefined& fDefined, float undef); /* ref:147 */
bool seaSoundSpeed(int nx, int ny, const float* t, const float* s, float z, int compute, float* soundspeed, ValuesDefined& fDefined,
                   float undef); /* ref:192 */
bool cvtemp(int nx, int ny, const float* tinp, int compute, float* tout, ValuesDefined& fDefined, float undef); /* ref:198 */
bool cvhum(int nx, int ny, const float* t, const float* huminp, const std::string& unit, int compute, float* humout, ValuesDefined& fDefined,
           float undef); /* ref:200 */
bool abshum(int nx, int ny, const float* t, const float* rhum, float* abshumout, ValuesDefined& fDefined, float undef); /* ref:202 */
bool underCooledRain(int nx, int ny, const float* precip, const float* snow, const float* tk, float precipMin, float snowRateMax,
                     float tcMax, float* undercooled, ValuesDefined& fDefined, float undef); /* ref:222 */

/* ---- pressure-level siblings, element functions, field arithmetic (the rest of SURVEY.md 8f rank 1) ---- */
bool plevelthe(int nx, int ny, const float* t, const float* rh, float p, int compute, float* the, ValuesDefined& fDefined, float undef); /* ref:115 */
bool pleveldz2tmean(int nx, int ny, const float* z1, const float* z2, float p1, float p2, int compute, float* tmean, ValuesDefined& fDefined,
                    float undef); /* ref:120 */
bool plevelducting(int nx, int ny, const float* t, const float* h, float p, int compute, float* duct, ValuesDefined& fDefined,
                   float undef); /* ref:125 */
bool vectorabs(int nx, int ny, const float* u, const float* v, float* ff, ValuesDefined& fDefined, float undef); /* ref:204 */
bool pressure2FlightLevel(int nx, int ny, const float* pressure, float* flightlevel, ValuesDefined& fDefined, float undef); /* ref:227 */
bool values2classes(int nx, int ny, const float* fvalue, float* fclass, const std::vector<float>& values, ValuesDefined& fDefined,
                    float undef); /* ref:252 */
void minvalueFields(int nx, int ny, const float* field1, const float* field2, float* fres, ValuesDefined& fDefined, float undef); /* ref:254 */
void minvalueFieldConst(int nx, int ny, const float* field1, const float value, float* fres, ValuesDefined& fDefined, float undef); /* ref:256 */
void maxvalueFields(int nx, int ny, const float* field1, const float* field2, float* fres, ValuesDefined& fDefined, float undef); /* ref:258 */
void maxvalueFieldConst(int nx, int ny, const float* field1, const float value, float* fres, ValuesDefined& fDefined, float undef); /* ref:260 */
void absvalueField(int nx, int ny, const float* field, float* fres, ValuesDefined& fDefined, float undef); /* ref:262 */
void log10Field(int nx, int ny, const float* field, float* fres, ValuesDefined& fDefined, float undef); /* ref:264 */
void pow10Field(int nx, int ny, const float* field, float* fres, ValuesDefined& fDefined, float undef); /* ref:266 */
void logField(int nx, int ny, const float* field, float* fres, ValuesDefined& fDefined, float undef); /* ref:268 */
void expField(int nx, int ny, const float* field, float* fres, ValuesDefined& fDefined, float undef); /* ref:270 */
void powerField(int nx, int ny, const float* field, float value, float* fres, ValuesDefined& fDefined, float undef); /* ref:272 */
void replaceUndefined(int nx, int ny, const float* field, float value, float* fres, ValuesDefined& fDefined, float undef); /* ref:274 */
void replaceDefined(int nx, int ny, const float* field, float value, float* fres, ValuesDefined& fDefined, float undef); /* ref:276 */
bool fieldOPERconstant(int compute, int nx, int ny, const float* field, float value, float* fres, ValuesDefined& fDefined,
                       float undef); /* ref:280 */
bool constantOPERfield(int compute, int nx, int ny, float value, const float* field, float* fres, ValuesDefined& fDefined,
                       float undef); /* ref:282 */
bool sumFields(int nx, int ny, const std::vector<float*>& fields, float* fres, ValuesDefined& fDefined, float undef); /* ref:284 */
bool snow_in_cm(int nx, int ny, const float* snow_water, const float* tk2m, const float* td2m, float* snow_cm, ValuesDefined& fDefined,
                float undef); /* ref:303 */

/* ---- geostrophic wind and vorticity in a pressure level ---- */
bool plevelgwind_xcomp(int nx, int ny, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis, float* ug,
                       ValuesDefined& fDefined, float undef); /* ref:127 */
bool plevelgwind_ycomp(int nx, int ny, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis, float* vg,
                       ValuesDefined& fDefined, float undef); /* ref:130 */
bool plevelgvort(int nx, int ny, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis, float* gvort,
                 ValuesDefined& fDefined, float undef); /* ref:133 */
bool plevelqvector(int nx, int ny, const float* z, const float* t, const float* xmapr, const float* ymapr, const float* fcoriolis, float p,
                   int compute, float* qcomp, ValuesDefined& fDefined, float undef); /* ref:122 */

/* ---- field arithmetic (compute first) ---- */
bool fieldOPERfield(int compute, int nx, int ny, const float* field1, const float* field2, float* fres, ValuesDefined& fDefined,
                    float undef); /* ref:278 */

/* ---- ensemble reductions over a list of member fields ---- */
bool meanValue(int nx, int ny, const std::vector<float*>& fields, const std::vector<ValuesDefined>& fDefinedIn, float* fres,
               ValuesDefined& fDefinedOut, float undef); /* ref:286 */
bool stddevValue(int nx, int ny, const std::vector<float*>& fields, const std::vector<ValuesDefined>& fDefinedIn, float* fres,
                 ValuesDefined& fDefinedOut, float undef); /* ref:289 */
bool extremeValue(int compute, int nx, int ny, const std::vector<float*>& fields, float* fres, ValuesDefined& fDefined, float undef); /* ref:292 */
bool probability(int compute, int nx, int ny, const std::vector<float*>& fields, const std::vector<ValuesDefined>& fDefinedIn,
                 const std::vector<float>& limits, float* fres, ValuesDefined& fDefinedOut, float undef); /* ref:294 */

/* ---- neighbourhood statistics ---- */
bool neighbourProbFunctions(int nx, int ny, const float* field, const std::vector<float>& constants, int compute, float* fres,
                            ValuesDefined& fDefined, float undef); /* ref:297 */
bool neighbourFunctions(int nx, int ny, const float* field, const std::vector<float>& constants, int compute, float* fres, ValuesDefined& fDefined,
                        float undef); /* ref:300 */

} // namespace fieldcalc
} // namespace miutil

#endif // MI_FIELDCALC_FIELDCALCULATIONS_H
