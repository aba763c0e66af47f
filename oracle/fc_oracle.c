/* oracle/fc_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C, single-threaded CPU restatement of the reference algorithm for the
 * FieldCalculations hot path (metno/mi-fieldcalc v0.1.9).  It exists so that the CUDA
 * product can be checked on a GPU box where /root/reference is absent.
 *
 * PARITY PINNED: this restatement is checked bit-for-bit (values, undefined mask and
 * ValuesDefined flag) against the unmodified reference compiled into oracle/_ref
 * (tests/test_oracle_vs_ref.py) and against the reference's own golden vectors
 * (test/FieldCalculationsTest.cc:70-305 restated in tests/test_golden.py).
 *
 * Citations: FC.cc = /root/reference/src/mi_fieldcalc/FieldCalculations.cc,
 * VI.cc = .../FieldCalculationsVesselIcing.cc, MC.h/.cc = .../MetConstants.{h,cc},
 * FC.h = .../FieldCalculations.h, FD.cc = .../FieldDefined.cc.
 *
 * The arithmetic mirrors the reference's C++ expression types: sub-expressions that
 * contain a double literal are evaluated in double and rounded to float once on
 * assignment; everything else is float.  Build with -ffp-contract=off (no FMA), like
 * the reference (-mavx2 without -mfma).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 */
#include "fc_oracle.h"

#include <limits.h>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

enum { ALL_DEFINED = 0, NONE_DEFINED = 1, SOME_DEFINED = 2 }; /* FieldDefined.h:41 */

/* ---------------------------------------------------------------- MC.h:39-59 constants */
static const float K_R = 287.f, K_CP = 1004.f, K_P0 = 1000.f, K_T0 = (float)273.15;
static const float K_EPS = (float)0.622, K_XLH = (float)2.501e+6;
static const float K_RHMIN = (float)0.02, K_RHMAX = (float)1.00;
#define K_P0INV ((float)(1. / K_P0))
#define K_KAPPA (K_R / K_CP)

#define N_EWT 41
static const float EWT[N_EWT] = {.000034, .000089, .000220, .000517, .001155, .002472, .005080, .01005, .01921, .03553, .06356,
                                 .1111,   .1891,   .3139,   .5088,   .8070,   1.2540,  1.9118,  2.8627, 4.2148, 6.1078, 8.7192,
                                 12.272,  17.044,  23.373,  31.671,  42.430,  56.236,  73.777,  95.855, 123.40, 157.46, 199.26,
                                 250.16,  311.69,  385.56,  473.67,  578.09,  701.13,  845.28,  1013.25};

/* ---------------------------------------------------------------- FC.h:42-98, FD.cc:62-70 */
static inline int is_def(float x, float undef)
{
  return !isnan(x) && x != undef;
}

static int check_defined(size_t n_undefined, size_t n)
{
  if (n_undefined == 0)
    return ALL_DEFINED;
  if (n_undefined == n)
    return NONE_DEFINED;
  return SOME_DEFINED;
}

/* ---------------------------------------------------------------- MC.h:61-84, MC.cc:37-45 */
typedef struct
{
  float x;
  int l;
} ewt_t;

static inline ewt_t ewt_make(float t_celsius)
{
  ewt_t e;
  e.x = (float)((t_celsius + 100.) * 0.2);
  /* int(x) of the reference is cvttss2si on x86: NaN and |x| >= 2^31 give INT_MIN */
  if (isnan(e.x) || e.x >= 2147483648.f || e.x < -2147483648.f)
    e.l = INT_MIN;
  else
    e.l = (int)e.x;
  return e;
}

static inline int ewt_defined(ewt_t e)
{
  return e.l >= 0 && e.l < N_EWT - 1;
}

static inline float ewt_value(ewt_t e)
{
  return EWT[e.l] + (EWT[e.l + 1] - EWT[e.l]) * (e.x - (float)e.l);
}

static inline float ewt_inverse(ewt_t e, float et)
{
  int ll = e.l;
  while (ll > 0 && ll < N_EWT - 1 && EWT[ll] > et)
    ll--;
  const float r = (et - EWT[ll]) / (EWT[ll + 1] - EWT[ll]);
  return (float)(-100. + ((float)ll + r) * 5.);
}

/* ---------------------------------------------------------------- FC.cc:186-316 scalar kernels */
static inline float clamp_rh(float rh)
{
  if (rh < K_RHMIN)
    return K_RHMIN;
  if (rh > K_RHMAX)
    return K_RHMAX;
  return rh;
}

static inline float pidcp_from_p(float p)
{
  return powf(p * K_P0INV, K_KAPPA);
}

static inline float pi_from_p(float p)
{
  return K_CP * pidcp_from_p(p);
}

static inline float p_hlevel(float ps, float a, float b)
{
  return a + b * ps;
}

static inline int bad_hlevel(float a, float b)
{
  return (a < 0.0) || (b < 0.0) || (a == 0.0 && b == 0.0) || (b > 1.0);
}

#define EWT_OR_UNDEF(e, tc)                                                                                                                          \
  const ewt_t e = ewt_make(tc);                                                                                                                      \
  if (!ewt_defined(e)) {                                                                                                                             \
    *nu += 1;                                                                                                                                        \
    return undef;                                                                                                                                    \
  }

static inline float t_thesat(float tk, float p, float pi, float undef, size_t* nu)
{
  EWT_OR_UNDEF(e, tk - K_T0)
  const float qsat = K_EPS * ewt_value(e) / p;
  return (K_CP * tk + K_XLH * qsat) / pi;
}

static inline float th_thesat(float th, float p, float pi, float undef, size_t* nu)
{
  EWT_OR_UNDEF(e, th * pi / K_CP - K_T0)
  const float qsat = K_EPS * ewt_value(e) / p;
  return th + K_XLH * qsat / pi;
}

static inline float tk_q_rh(float tk, float q, float p, float undef, size_t* nu)
{
  EWT_OR_UNDEF(e, tk - K_T0)
  const float qsat = K_EPS * ewt_value(e) / p;
  return (float)(100. * q / qsat);
}

static inline float tk_rh_q(float tk, float rh, float p, float undef, size_t* nu)
{
  EWT_OR_UNDEF(e, tk - K_T0)
  const float qsat = K_EPS * ewt_value(e) / p;
  return (float)(0.01 * rh * qsat);
}

static inline float tk_q_td(float tk, float q, float p, float tdconv, float undef, size_t* nu)
{
  EWT_OR_UNDEF(e, tk - K_T0)
  const float et = ewt_value(e);
  const float qsat = K_EPS * et / p;
  const float rh = clamp_rh(q / qsat);
  const float etd = rh * et;
  return ewt_inverse(e, etd) + tdconv;
}

static inline float tk_rh_td(float tk, float rh100, float tdconv, float undef, size_t* nu)
{
  EWT_OR_UNDEF(e, tk - K_T0)
  const float et = ewt_value(e);
  const float rh = clamp_rh((float)(0.01 * rh100));
  const float etd = rh * et;
  return ewt_inverse(e, etd) + tdconv;
}

static inline float tk_q_duct(float tk, float q, float p)
{
  return (float)(77.6 * (p / tk) + 373000. * (q * p) / (K_EPS * tk * tk));
}

static inline float tk_rh_duct(float tk, float q, float p, float undef, size_t* nu)
{
  EWT_OR_UNDEF(e, tk - K_T0)
  const float et = ewt_value(e);
  const float rh = clamp_rh((float)(q * 0.01));
  return (float)(77.6 * (p / tk) + 373000. * rh * et / (tk * tk));
}

/* unit -> compute remapping shared by the *leveltemp functions (FC.cc:340-345, 1060-1065, 1322-1327) */
static int temp_compute(int compute, const char* unit)
{
  if (compute < 3) {
    if (strcmp(unit, "celsius") == 0)
      return 1;
    if (strcmp(unit, "kelvin") == 0)
      return 2;
  }
  return compute;
}

/* unit -> compute remapping shared by the *levelhum functions (FC.cc:422-425, 1174-1177, 1417-1420) */
static int hum_compute(int compute, const char* unit)
{
  if (compute > 8 && strcmp(unit, "celsius") == 0)
    return compute - 4;
  if (compute > 4 && compute <= 8 && strcmp(unit, "kelvin") == 0)
    return compute + 4;
  return compute;
}

/* border ring := nearest interior value, columns then rows (FC.cc:59-74) */
static void fill_edges(int nx, int ny, float* f)
{
  for (int y = 1; y < ny - 1; ++y) {
    f[y * nx] = f[y * nx + 1];
    f[y * nx + nx - 1] = f[y * nx + nx - 2];
  }
  for (int x = 0; x < nx; ++x) {
    f[x] = f[x + nx];
    f[(ny - 1) * nx + x] = f[(ny - 2) * nx + x];
  }
}

/* ================================================================== pressure levels */

int fco_pleveltemp(int nx, int ny, const float* tinp, float p, const char* unit, int compute, float* tout, int* fDefined, float undef)
{ /* FC.cc:328-367 */
  if (p <= 0)
    return 0;
  compute = temp_compute(compute, unit);
  if (compute < 1 || compute > 5)
    return 0;
  const float pidcp = pidcp_from_p(p), pi = pidcp * K_CP;
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (size_t i = 0; i < n; ++i) {
    const float f = tinp[i];
    if (!(all || is_def(f, undef))) {
      tout[i] = undef;
      nundef += 1;
      continue;
    }
    switch (compute) {
    case 1:
      tout[i] = f * pidcp - K_T0;
      break;
    case 2:
      tout[i] = f * pidcp;
      break;
    case 3:
      tout[i] = f / pidcp;
      break;
    case 4:
      tout[i] = t_thesat(f, p, pi, undef, &nundef);
      break;
    default:
      tout[i] = th_thesat(f, p, pi, undef, &nundef);
      break;
    }
  }
  if (compute >= 4) /* c1-3 go through unaryFunctionField, which leaves the flag alone (FC.cc:94-106) */
    *fDefined = check_defined(nundef, n);
  return 1;
}

/* one point of the 12-mode humidity conversion in the a/h-level numbering (FC.cc:1189-1209, 1430-1450);
 * `tk` is already converted from potential temperature for the even modes */
static inline float hum_point_ah(int compute, float tk, float hum, float p, float tdconv, float undef, size_t* nu)
{
  switch (compute) {
  case 1:
  case 2:
    return tk_q_rh(tk, hum, p, undef, nu);
  case 3:
  case 4:
    return tk_rh_q(tk, hum, p, undef, nu);
  case 5:
  case 6:
  case 9:
  case 10:
    return tk_q_td(tk, hum, p, tdconv, undef, nu);
  default: /* 7, 8, 11, 12 */
    return tk_rh_td(tk, hum, tdconv, undef, nu);
  }
}

int fco_plevelhum(int nx, int ny, const float* t, const float* huminp, float p, const char* unit, int compute, float* humout, int* fDefined,
                  float undef)
{ /* FC.cc:400-464; note the p-level numbering: 5,6,9,10 = RH -> Td and 7,8,11,12 = q -> Td */
  if (p <= 0 || compute <= 0 || compute >= 13)
    return 0;
  compute = hum_compute(compute, unit);
  const size_t n = (size_t)nx * ny;
  const int rh_td = (compute == 5 || compute == 6 || compute == 9 || compute == 10);
  if (p == undef && !rh_td) {
    for (size_t i = 0; i < n; ++i)
      humout[i] = undef;
    *fDefined = NONE_DEFINED;
    return 1;
  }
  const float pi = pi_from_p(p);
  const float tconv = (compute % 2 == 0) ? (pi / K_CP) : 1;
  const float tdconv = (compute >= 9) ? K_T0 : 0;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(t[i], undef) && is_def(huminp[i], undef))) {
      const float tk = t[i] * tconv;
      if (compute <= 2)
        humout[i] = tk_q_rh(tk, huminp[i], p, undef, &nundef);
      else if (compute <= 4)
        humout[i] = tk_rh_q(tk, huminp[i], p, undef, &nundef);
      else if (rh_td)
        humout[i] = tk_rh_td(tk, huminp[i], tdconv, undef, &nundef);
      else
        humout[i] = tk_q_td(tk, huminp[i], p, tdconv, undef, &nundef);
    } else {
      humout[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, n);
  return 1;
}

/* ================================================================== hybrid + atmospheric levels
 * The h-level functions evaluate p = alevel + blevel*ps[i] per point (FC.cc:303-306) and then do
 * what the a-level functions do with p[i]; the differences are called out where they exist. */

static int xleveltemp(int hybrid, size_t n, const float* tinp, const float* pin, float a, float b, int compute, float* tout, int* fDefined,
                      float undef)
{ /* FC.cc:1074-1097 (hybrid), 1329-1352 (atmospheric) */
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(tinp[i], undef) && is_def(pin[i], undef))) {
      const float p = hybrid ? p_hlevel(pin[i], a, b) : pin[i];
      const float pidcp = pidcp_from_p(p);
      if (compute == 1)
        tout[i] = tinp[i] * pidcp - K_T0;
      else if (compute == 2)
        tout[i] = tinp[i] * pidcp;
      else if (compute == 3)
        tout[i] = tinp[i] / pidcp;
      else if (compute == 4)
        tout[i] = t_thesat(tinp[i], p, K_CP * pidcp, undef, &nundef);
      else if (compute == 5)
        tout[i] = th_thesat(tinp[i], p, K_CP * pidcp, undef, &nundef);
      /* any other compute: hleveltemp leaves the point untouched (no validation, FC.cc:1080-1090) */
    } else {
      tout[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, n);
  return 1;
}

int fco_hleveltemp(int nx, int ny, const float* tinp, const float* ps, float alevel, float blevel, const char* unit, int compute, float* tout,
                   int* fDefined, float undef)
{ /* FC.cc:1046-1098 */
  compute = temp_compute(compute, unit);
  if (bad_hlevel(alevel, blevel))
    return 0;
  return xleveltemp(1, (size_t)nx * ny, tinp, ps, alevel, blevel, compute, tout, fDefined, undef);
}

int fco_aleveltemp(int nx, int ny, const float* tinp, const float* p, const char* unit, int compute, float* tout, int* fDefined, float undef)
{ /* FC.cc:1310-1353 */
  if (compute <= 0 || compute >= 6)
    return 0;
  compute = temp_compute(compute, unit);
  return xleveltemp(0, (size_t)nx * ny, tinp, p, 0, 0, compute, tout, fDefined, undef);
}

static int xlevelthe(int hybrid, size_t n, const float* t, const float* q, const float* pin, float a, float b, int compute, float* the,
                     int* fDefined, float undef)
{ /* FC.cc:1125-1142, 1375-1391 */
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(t[i], undef) && is_def(q[i], undef) && is_def(pin[i], undef))) {
      const float p = hybrid ? p_hlevel(pin[i], a, b) : pin[i];
      const float pi = pi_from_p(p);
      if (compute == 1)
        the[i] = (t[i] * K_CP + q[i] * K_XLH) / pi;
      else if (compute == 2)
        the[i] = t[i] + q[i] * K_XLH / pi;
    } else {
      the[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, n);
  return 1;
}

int fco_hlevelthe(int nx, int ny, const float* t, const float* q, const float* ps, float alevel, float blevel, int compute, float* the,
                  int* fDefined, float undef)
{ /* FC.cc:1100-1143: no validation of compute */
  if (bad_hlevel(alevel, blevel))
    return 0;
  return xlevelthe(1, (size_t)nx * ny, t, q, ps, alevel, blevel, compute, the, fDefined, undef);
}

int fco_alevelthe(int nx, int ny, const float* t, const float* q, const float* p, int compute, float* the, int* fDefined, float undef)
{ /* FC.cc:1355-1392 */
  if (compute != 1 && compute != 2)
    return 0;
  return xlevelthe(0, (size_t)nx * ny, t, q, p, 0, 0, compute, the, fDefined, undef);
}

int fco_hlevelhum(int nx, int ny, const float* t, const float* huminp, const float* ps, float alevel, float blevel, const char* unit, int compute,
                  float* humout, int* fDefined, float undef)
{ /* FC.cc:1145-1217 */
  if (compute <= 0 || compute >= 13)
    return 0;
  if (bad_hlevel(alevel, blevel))
    return 0;
  compute = hum_compute(compute, unit);
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  const float tdconv = (compute >= 9) ? K_T0 : 0;
  const int need_p = !(compute == 7 || compute == 11);
  size_t nundef = 0;
  for (size_t i = 0; i < n; ++i) {
    if ((all || (is_def(t[i], undef) && is_def(huminp[i], undef))) && (!need_p || all || ps[i] != undef)) {
      const float p = need_p ? p_hlevel(ps[i], alevel, blevel) : 0;
      float tk = t[i];
      if (compute % 2 == 0)
        tk = t[i] * pidcp_from_p(p);
      humout[i] = hum_point_ah(compute, tk, huminp[i], p, tdconv, undef, &nundef);
    } else {
      humout[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, n);
  return 1;
}

int fco_alevelhum(int nx, int ny, const float* t, const float* huminp, const float* p, const char* unit, int compute, float* humout, int* fDefined,
                  float undef)
{ /* FC.cc:1394-1458; quirk at :1429 -- p[i] is only tested when compute is 7 or 11, i.e. when it is not used */
  if (compute <= 0 || compute >= 13)
    return 0;
  compute = hum_compute(compute, unit);
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  const float tdconv = (compute >= 9) ? K_T0 : 0;
  const int p_unused = (compute == 7 || compute == 11);
  size_t nundef = 0;
  for (size_t i = 0; i < n; ++i) {
    if ((all || (is_def(t[i], undef) && is_def(huminp[i], undef))) && (!p_unused || all || p[i] != undef)) {
      float tk = t[i];
      if (compute % 2 == 0)
        tk = t[i] * pidcp_from_p(p[i]);
      humout[i] = hum_point_ah(compute, tk, huminp[i], p[i], tdconv, undef, &nundef);
    } else {
      humout[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, n);
  return 1;
}

static void xlevelducting(int hybrid, size_t n, const float* t, const float* h, const float* pin, float a, float b, int compute, float* duct,
                          int all, float undef, size_t* nundef)
{ /* FC.cc:1256-1271, 1490-1503 */
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(t[i], undef) && is_def(h[i], undef) && is_def(pin[i], undef))) {
      const float p = hybrid ? p_hlevel(pin[i], a, b) : pin[i];
      float tk = t[i];
      if (compute % 2 == 0)
        tk *= pidcp_from_p(p);
      if (compute == 1 || compute == 2)
        duct[i] = tk_q_duct(tk, h[i], p);
      else if (compute == 3 || compute == 4)
        duct[i] = tk_rh_duct(tk, h[i], p, undef, nundef);
    } else {
      duct[i] = undef;
      *nundef += 1;
    }
  }
}

int fco_hlevelducting(int nx, int ny, const float* t, const float* h, const float* ps, float alevel, float blevel, int compute, float* duct,
                      int* fDefined, float undef)
{ /* FC.cc:1219-1274 */
  if (bad_hlevel(alevel, blevel))
    return 0;
  const size_t n = (size_t)nx * ny;
  size_t nundef = 0;
  xlevelducting(1, n, t, h, ps, alevel, blevel, compute, duct, *fDefined == ALL_DEFINED, undef, &nundef);
  *fDefined = check_defined(nundef, n);
  return 1;
}

int fco_alevelducting(int nx, int ny, const float* t, const float* h, const float* p, int compute, float* duct, int* fDefined, float undef)
{ /* FC.cc:1460-1505: the flag is never updated and compute is never validated */
  size_t nundef = 0;
  xlevelducting(0, (size_t)nx * ny, t, h, p, 0, 0, compute, duct, *fDefined == ALL_DEFINED, undef, &nundef);
  return 1;
}

int fco_hlevelpressure(int nx, int ny, const float* ps, float alevel, float blevel, float* p, int* fDefined, float undef)
{ /* FC.cc:1276-1304 */
  if (bad_hlevel(alevel, blevel))
    return 0;
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || is_def(ps[i], undef))
      p[i] = p_hlevel(ps[i], alevel, blevel);
    else {
      p[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, n);
  return 1;
}

/* ================================================================== stencil family
 * All loops run over the FLAT index range [nx, N-nx) -- edge columns included, with their
 * wrapped i-1 / i+1 neighbours -- count undefined points there, derive the flag from
 * N - 2*nx, and only then overwrite the border ring with fill_edges. */

#define DEF4(a, b, c, d) (all || (is_def(a, undef) && is_def(b, undef) && is_def(c, undef) && is_def(d, undef)))

int fco_ilevelgwind(int nx, int ny, const float* mpot, const float* xmapr, const float* ymapr, const float* fcoriolis, float* ug, float* vg,
                    int* fDefined, float undef)
{ /* FC.cc:1511-1549; the flag denominator is N, not N-2nx (:1543) */
  if (nx < 3 || ny < 3)
    return 0;
  const int n = nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (int i = nx; i < n - nx; ++i) {
    if (DEF4(mpot[i - nx], mpot[i - 1], mpot[i + 1], mpot[i + nx])) {
      ug[i] = (float)(-0.5 * ymapr[i] * (mpot[i + nx] - mpot[i - nx]) / fcoriolis[i]);
      vg[i] = (float)(0.5 * xmapr[i] * (mpot[i + 1] - mpot[i - 1]) / fcoriolis[i]);
    } else {
      ug[i] = undef;
      vg[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, (size_t)n);
  fill_edges(nx, ny, ug);
  fill_edges(nx, ny, vg);
  return 1;
}

int fco_relvort(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, float* rvort, int* fDefined, float undef)
{ /* FC.cc:1843-1873 */
  if (nx < 3 || ny < 3)
    return 0;
  const int n = nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (int i = nx; i < n - nx; ++i) {
    if (DEF4(v[i - 1], v[i + 1], u[i - nx], u[i + nx]))
      rvort[i] = (float)(0.5 * xmapr[i] * (v[i + 1] - v[i - 1]) - 0.5 * ymapr[i] * (u[i + nx] - u[i - nx]));
    else {
      rvort[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, (size_t)(n - 2 * nx));
  fill_edges(nx, ny, rvort);
  return 1;
}

int fco_absvort(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, const float* fcoriolis, float* avort,
                int* fDefined, float undef)
{ /* FC.cc:1875-1908 */
  if (nx < 3 || ny < 3)
    return 0;
  const int n = nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (int i = nx; i < n - nx; ++i) {
    if (DEF4(v[i - 1], v[i + 1], u[i - nx], u[i + nx]))
      avort[i] = (float)(0.5 * xmapr[i] * (v[i + 1] - v[i - 1]) - 0.5 * ymapr[i] * (u[i + nx] - u[i - nx]) + fcoriolis[i]);
    else {
      avort[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, (size_t)(n - 2 * nx));
  fill_edges(nx, ny, avort);
  return 1;
}

int fco_divergence(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, float* diverg, int* fDefined,
                   float undef)
{ /* FC.cc:1910-1940; quirk: the definedness test reads v[i+-1], u[i+-nx] (as relvort) but the value reads u[i+-1], v[i+-nx] */
  if (nx < 3 || ny < 3)
    return 0;
  const int n = nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (int i = nx; i < n - nx; ++i) {
    if (DEF4(v[i - 1], v[i + 1], u[i - nx], u[i + nx]))
      diverg[i] = (float)(0.5 * xmapr[i] * (u[i + 1] - u[i - 1]) + 0.5 * ymapr[i] * (v[i + nx] - v[i - nx]));
    else {
      diverg[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, (size_t)(n - 2 * nx));
  fill_edges(nx, ny, diverg);
  return 1;
}

int fco_advection(int nx, int ny, const float* f, const float* u, const float* v, const float* xmapr, const float* ymapr, float hours,
                  float* advec, int* fDefined, float undef)
{ /* FC.cc:1942-1983 */
  if (nx < 3 || ny < 3)
    return 0;
  const float scale = (float)(-3600. * hours);
  const int n = nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (int i = nx; i < n - nx; ++i) {
    if (all || (is_def(u[i], undef) && is_def(v[i], undef) && is_def(f[i - nx], undef) && is_def(f[i - 1], undef) && is_def(f[i + 1], undef) &&
                is_def(f[i + nx], undef)))
      advec[i] = (float)((u[i] * 0.5 * xmapr[i] * (f[i + 1] - f[i - 1]) + v[i] * 0.5 * ymapr[i] * (f[i + nx] - f[i - nx])) * scale);
    else {
      advec[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, (size_t)(n - 2 * nx));
  fill_edges(nx, ny, advec);
  return 1;
}

int fco_gradient(int nx, int ny, const float* field, const float* xmapr, const float* ymapr, int compute, float* fgrad, int* fDefined,
                 float undef)
{ /* FC.cc:1985-2074; c=1 loops over [1, N-1) but the flag still uses N-2nx (:2068) */
  if (nx < 3 || ny < 3)
    return 0;
  const int n = nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  switch (compute) {
  case 1:
    for (int i = 1; i < n - 1; ++i) {
      if (all || (is_def(field[i - 1], undef) && is_def(field[i + 1], undef)))
        fgrad[i] = (float)(0.5 * xmapr[i] * (field[i + 1] - field[i - 1]));
      else {
        fgrad[i] = undef;
        nundef += 1;
      }
    }
    break;
  case 2:
    for (int i = nx; i < n - nx; ++i) {
      if (all || (is_def(field[i - nx], undef) && is_def(field[i + nx], undef)))
        fgrad[i] = (float)(0.5 * ymapr[i] * (field[i + nx] - field[i - nx]));
      else {
        fgrad[i] = undef;
        nundef += 1;
      }
    }
    break;
  case 3:
    for (int i = nx; i < n - nx; ++i) {
      if (DEF4(field[i - nx], field[i - 1], field[i + 1], field[i + nx])) {
        const float dfdx = (float)(0.5 * xmapr[i] * (field[i + 1] - field[i - 1]));
        const float dfdy = (float)(0.5 * ymapr[i] * (field[i + nx] - field[i - nx]));
        fgrad[i] = sqrtf(dfdx * dfdx + dfdy * dfdy);
      } else {
        fgrad[i] = undef;
        nundef += 1;
      }
    }
    break;
  case 4:
    for (int i = nx; i < n - nx; ++i) {
      if (DEF4(field[i - nx], field[i - 1], field[i + 1], field[i + nx]) && (all || is_def(field[i], undef))) {
        const float d2fdx = (float)(field[i - 1] - 2.0 * field[i] + field[i + 1]);
        const float d2fdy = (float)(field[i - nx] - 2.0 * field[i] + field[i + nx]);
        fgrad[i] = (float)(4.0 * (0.25 * xmapr[i] * xmapr[i] * d2fdx + 0.25 * ymapr[i] * ymapr[i] * d2fdy));
      } else {
        fgrad[i] = undef;
        nundef += 1;
      }
    }
    break;
  default:
    return 0;
  }
  *fDefined = check_defined(nundef, (size_t)(n - 2 * nx));
  fill_edges(nx, ny, fgrad);
  return 1;
}

int fco_shapiro2_filter(int nx, int ny, float* field, float* fsmooth, int* fDefined, float undef)
{ /* FC.cc:2076-2179 */
  if (nx < 3 || ny < 3)
    return 0;
  const int n = nx * ny;
  float* f1 = fsmooth;
  if (field != fsmooth)
    memcpy(fsmooth, field, sizeof(float) * (size_t)n);
  float* f2 = (float*)malloc(sizeof(float) * (size_t)n);
  const int all = *fDefined == ALL_DEFINED;
  if (all) {
    /* :2108-2131, weights +0.25 then -0.25; float sum of the neighbours, the rest in double */
    float s = 0.25f;
    for (int it = 0; it < 2; ++it) {
      for (int i = 1; i < n - 1; ++i)
        f2[i] = (float)(f1[i] + s * (f1[i - 1] + f1[i + 1] - 2. * f1[i]));
      for (int y = 0; y < ny; ++y) {
        f2[y * nx] = f1[y * nx];
        f2[y * nx + nx - 1] = f1[y * nx + nx - 1];
      }
      for (int i = nx; i < n - nx; ++i)
        f1[i] = (float)(f2[i] + s * (f2[i - nx] + f2[i + nx] - 2. * f2[i]));
      for (int x = 0; x < nx; ++x) {
        f1[x] = f2[x];
        f1[n - nx + x] = f2[n - nx + x];
      }
      s = -0.25f;
    }
  } else {
    /* :2133-2172, per-point weights from the ORIGINAL field; both iterations use +0.25;
     * all-float arithmetic (the literal 2 is an int) */
    float* s1 = (float*)malloc(sizeof(float) * (size_t)n);
    float* s2 = (float*)malloc(sizeof(float) * (size_t)n);
    for (int i = 1; i < n - 1; ++i)
      s1[i] = (is_def(f1[i - 1], undef) && is_def(f1[i], undef) && is_def(f1[i + 1], undef)) ? 0.25f : 0.f;
    for (int i = nx; i < n - nx; ++i)
      s2[i] = (is_def(f1[i - nx], undef) && is_def(f1[i], undef) && is_def(f1[i + nx], undef)) ? 0.25f : 0.f;
    for (int it = 0; it < 2; ++it) {
      for (int i = 1; i < n - 1; ++i)
        f2[i] = f1[i] + s1[i] * (f1[i - 1] + f1[i + 1] - 2 * f1[i]);
      for (int y = 0; y < ny; ++y) {
        f2[y * nx] = f1[y * nx];
        f2[y * nx + nx - 1] = f1[y * nx + nx - 1];
      }
      for (int i = nx; i < n - nx; ++i)
        f1[i] = f2[i] + s2[i] * (f2[i - nx] + f2[i + nx] - 2 * f2[i]);
      for (int x = 0; x < nx; ++x) {
        f1[x] = f2[x];
        f1[n - nx + x] = f2[n - nx + x];
      }
    }
    free(s1);
    free(s2);
  }
  free(f2);
  *fDefined = ALL_DEFINED; /* :2176 */
  return 1;
}

int fco_windCooling(int nx, int ny, const float* t, const float* u, const float* v, int compute, float* dtcool, int* fDefined, float undef)
{ /* FC.cc:2181-2229; the flag is never updated */
  if (compute != 1 && compute != 2)
    return 0;
  const float tconv = (compute == 1) ? K_T0 : 0.f;
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(t[i], undef) && is_def(u[i], undef) && is_def(v[i], undef))) {
      const float tc = t[i] - tconv;
      const float ff = (float)(sqrtf(u[i] * u[i] + v[i] * v[i]) * 3.6);
      const float ffpow = powf(ff, (float)0.16);
      float d = (float)(13.12 + 0.6215 * tc - 11.37 * ffpow + 0.3965 * tc * ffpow);
      if (d > 0.)
        d = 0.f;
      dtcool[i] = d;
    } else
      dtcool[i] = undef;
  }
  return 1;
}

int fco_thermalFrontParameter(int nx, int ny, const float* tx, const float* xmapr, const float* ymapr, float* tfp, int* fDefined, float undef)
{ /* FC.cc:2266-2309: gradient(c=3) into a scratch field -- with its fill_edges and its flag update --
   * then a second five-point pass whose allDefined comes from the first pass's OUTPUT flag */
  const int n = nx * ny;
  if (nx < 3 || ny < 3)
    return 0;
  float* ad = (float*)malloc(sizeof(float) * (size_t)n);
  if (!fco_gradient(nx, ny, tx, xmapr, ymapr, 3, ad, fDefined, undef)) {
    free(ad);
    return 0;
  }
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (int i = nx; i < n - nx; ++i) {
    const int ok = all || (is_def(tx[i - nx], undef) && is_def(tx[i - 1], undef) && is_def(tx[i + 1], undef) && is_def(tx[i + nx], undef) &&
                           is_def(ad[i - nx], undef) && is_def(ad[i - 1], undef) && is_def(ad[i], undef) && is_def(ad[i + 1], undef) &&
                           is_def(ad[i + nx], undef));
    if (ok && ad[i] != 0) {
      const float dadx = (float)(0.5 * xmapr[i] * (ad[i + 1] - ad[i - 1]));
      const float dady = (float)(0.5 * ymapr[i] * (ad[i + nx] - ad[i - nx]));
      const float dtdxa = (float)(0.5 * xmapr[i] * (tx[i + 1] - tx[i - 1]) / ad[i]);
      const float dtdya = (float)(0.5 * ymapr[i] * (tx[i + nx] - tx[i - nx]) / ad[i]);
      tfp[i] = -(dadx * dtdxa + dady * dtdya);
    } else {
      tfp[i] = undef;
      nundef += 1;
    }
  }
  free(ad);
  *fDefined = check_defined(nundef, (size_t)(n - 2 * nx));
  fill_edges(nx, ny, tfp);
  return 1;
}

static int momentum_coordinate(int xdir, int nx, int ny, const float* w, const float* mapr, const float* fcoriolis, float fcoriolisMin, float* out,
                               int* fDefined, float undef)
{ /* FC.cc:2351-2386 (x), 2388-2422 (y) */
  if (nx < 3 || ny < 3)
    return 0;
  const int n = nx * ny;
  const float fcormin = fabsf(fcoriolisMin), fcormax = -fcormin;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (int i = 0; i < n; ++i) {
    if (all || is_def(w[i], undef)) {
      float fcor = fcoriolis[i];
      if (fcor >= 0. && fcor < fcormin)
        fcor = fcormin;
      else if (fcor <= 0. && fcor > fcormax)
        fcor = fcormax;
      if (xdir)
        out[i] = (float)(i % nx) + w[i] * mapr[i] / fcor;
      else
        out[i] = (float)(i / nx) - w[i] * mapr[i] / fcor;
    } else {
      out[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, (size_t)n);
  return 1;
}

int fco_momentumXcoordinate(int nx, int ny, const float* v, const float* xmapr, const float* fcoriolis, float fcoriolisMin, float* mxy,
                            int* fDefined, float undef)
{
  return momentum_coordinate(1, nx, ny, v, xmapr, fcoriolis, fcoriolisMin, mxy, fDefined, undef);
}

int fco_momentumYcoordinate(int nx, int ny, const float* u, const float* ymapr, const float* fcoriolis, float fcoriolisMin, float* nxy,
                            int* fDefined, float undef)
{
  return momentum_coordinate(0, nx, ny, u, ymapr, fcoriolis, fcoriolisMin, nxy, fDefined, undef);
}

int fco_jacobian(int nx, int ny, const float* f1, const float* f2, const float* xmapr, const float* ymapr, float* fjacobian, int* fDefined,
                 float undef)
{ /* FC.cc:2424-2460: four derivatives rounded to float, then a*b - c*d in float without FMA */
  if (nx < 3 || ny < 3)
    return 0;
  const int n = nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (int i = nx; i < n - nx; ++i) {
    if (DEF4(f1[i - nx], f1[i - 1], f1[i + 1], f1[i + nx]) && DEF4(f2[i - nx], f2[i - 1], f2[i + 1], f2[i + nx])) {
      const float df1dx = (float)(0.5 * xmapr[i] * (f1[i + 1] - f1[i - 1]));
      const float df1dy = (float)(0.5 * ymapr[i] * (f1[i + nx] - f1[i - nx]));
      const float df2dx = (float)(0.5 * xmapr[i] * (f2[i + 1] - f2[i - 1]));
      const float df2dy = (float)(0.5 * ymapr[i] * (f2[i + nx] - f2[i - nx]));
      fjacobian[i] = df1dx * df2dy - df1dy * df2dx;
    } else {
      fjacobian[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, (size_t)(n - 2 * nx));
  fill_edges(nx, ny, fjacobian);
  return 1;
}

/* ================================================================== vessel icing */

static inline int def6(int all, float a, float b, float c, float d, float e, float f, float undef)
{
  return all || (is_def(a, undef) && is_def(b, undef) && is_def(c, undef) && is_def(d, undef) && is_def(e, undef) && is_def(f, undef));
}

/* freezing point of sea water, Stallabrass (1980): double, but the square of sal is a float product (VI.cc:95, 127) */
static inline double freezing_point(float sal)
{
  return (-0.002 - 0.0524 * sal) - 6.0E-5 * (sal * sal);
}

int fco_vesselIcingOverland(int nx, int ny, const float* airtemp, const float* seatemp, const float* u, const float* v, const float* sal,
                            const float* aice, float* icing, int* fDefined, float undef)
{ /* VI.cc:77-112 */
  const size_t n = (size_t)nx * ny;
  const double A = 2.73e-2, B = 2.91e-4, C = 1.84e-6;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (size_t i = 0; i < n; ++i) {
    if (def6(all, airtemp[i], seatemp[i], u[i], v[i], sal[i], aice[i], undef) && aice[i] < 0.4) {
      const double Tf = freezing_point(sal[i]);
      if (seatemp[i] < Tf) {
        icing[i] = undef;
        nundef += 1;
      } else {
        const double ff = sqrtf(u[i] * u[i] + v[i] * v[i]);
        const double ppr = ff * (Tf - airtemp[i]) / (1 + 0.3 * (seatemp[i] - Tf));
        icing[i] = (float)(A * ppr + B * (ppr * ppr) + C * ppr * ppr * ppr);
      }
    } else {
      icing[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, n);
  return 1;
}

int fco_vesselIcingMertins(int nx, int ny, const float* airtemp, const float* seatemp, const float* u, const float* v, const float* sal,
                           const float* aice, float* icing, int* fDefined, float undef)
{ /* VI.cc:114-180: wind class x air/sea temperature decision table */
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (size_t i = 0; i < n; ++i) {
    if (!(def6(all, airtemp[i], seatemp[i], u[i], v[i], sal[i], aice[i], undef) && aice[i] < 0.4)) {
      icing[i] = undef;
      nundef += 1;
      continue;
    }
    const double Tf = freezing_point(sal[i]);
    if (seatemp[i] < Tf) {
      icing[i] = undef;
      nundef += 1;
      continue;
    }
    const double ff = sqrtf(u[i] * u[i] + v[i] * v[i]);
    const double ta = airtemp[i], sst = seatemp[i];
    if (!(ff >= 10.8)) {
      icing[i] = 0;
      continue;
    }
    double t1, t2, t3;
    if (ff < 17.2) {
      t1 = -1.15 * sst - 4.3;
      t2 = -1.5 * sst - 10;
      t3 = -10000;
    } else if (ff < 20.8) {
      t1 = -0.6 * sst - 3.2;
      t2 = -1.05 * sst - 5.6;
      t3 = -1.75 * sst - 12.5;
    } else if (ff < 28.5) {
      t1 = -0.3 * sst - 2.6;
      t2 = -0.66 * sst - 3.32;
      t3 = -1.325 * sst - 7.651;
    } else {
      t1 = -0.14 * sst - 2.28;
      t2 = -0.3 * sst - 2.6;
      t3 = -1.16 * sst - 5.22;
    }
    if (ta > -2)
      icing[i] = 0;
    else if (ta > t1)
      icing[i] = (float)0.8333;
    else if (ta > t2)
      icing[i] = (float)2.0833;
    else if (ta <= t3 || ff < 17.2)
      icing[i] = (float)4.375;
    else
      icing[i] = (float)6.25;
  }
  *fDefined = check_defined(nundef, n);
  return 1;
}

static inline double icing_f1_d(double t)
{ /* VI.cc:53-57 with T = double */
  return 0.6112 * exp(17.67 * t / (t + 243.5));
}

static inline float icing_f1_f(float t)
{ /* VI.cc:53-57 with T = float */
  return (float)0.6112 * expf((float)17.67 * t / (t + (float)243.5));
}

static inline float kT4_f(float tc)
{ /* VI.cc:65-70 with T = float */
  const float sigma = (float)5.67e-8;
  const float a = tc + K_T0;
  const float a2 = a * a;
  return sigma * (a2 * a2);
}

static inline int def10(int all, float a, float b, float c, float d, float e, float f, float g, float h, float i, float j, float undef)
{
  return all || (is_def(a, undef) && is_def(b, undef) && is_def(c, undef) && is_def(d, undef) && is_def(e, undef) && is_def(f, undef) &&
                 is_def(g, undef) && is_def(h, undef) && is_def(i, undef) && is_def(j, undef));
}

int fco_vesselIcingModStall(int nx, int ny, const float* sal, const float* wave, const float* x_wind, const float* y_wind, const float* airtemp,
                            const float* rh, const float* sst, const float* p, const float* Pw, const float* aice, const float* depth, float vs,
                            float alpha, float zmin, float zmax, float* icing, int* fDefined, float undef)
{ /* VI.cc:182-337: everything in double; Pw is not part of the definedness test */
  const size_t n = (size_t)nx * ny;
  const double num = zmax - zmin;
  const int number = (int)(num * 2 + 1);
  if (zmax < zmin || fmod(num, 1) != 0)
    return 0;
  if (vs < 0 || alpha < 0 || zmin < 0 || zmax < 0)
    return 0;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (size_t i = 0; i < n; ++i) {
    if (!(def10(all, sal[i], wave[i], x_wind[i], y_wind[i], airtemp[i], rh[i], sst[i], p[i], aice[i], depth[i], undef) && aice[i] < 0.4)) {
      icing[i] = undef;
      nundef += 1;
      continue;
    }
    /* deep-water wave speed, then the shallow-water fixed point (:218-237) */
    double c = (9.81 / (2 * M_PI)) * Pw[i];
    if (depth[i] <= c * Pw[i] && c != 0) {
      c = 1.0;
      double err = 1.0;
      int j = 0;
      while (err > 1e-5) {
        const double c_new = (9.81 * Pw[i] / (2 * M_PI)) * tanh(2 * M_PI * depth[i] / (Pw[i] * c));
        err = fabs(c_new - c);
        c = c_new;
        j = j + 1;
        if (j > 10000) {
          c = 0.0;
          break;
        }
      }
    }
    const double Vr = c - vs * cos((double)alpha);
    const double v = sqrtf(x_wind[i] * x_wind[i] + y_wind[i] * y_wind[i]);
    const double Tf = freezing_point(sal[i]);
    const double ha_ = 5.17, ha = ha_ * pow(v, 0.8);
    const double ratio = 89.5 / ha_;
    const double tau = 11.25 - v / 4.0;
    double td = sst[i];
    if (tau > 0.0) { /* droplet cooling, RK4 with 50 steps (:262-281) */
      const double K = 311000.0 / ((p[i] / 10.0) * 1005.0);
      const double M = 0.2 * airtemp[i] + K * rh[i] * icing_f1_f(airtemp[i]); /* float argument -> icing_f1<float> */
      const double h = tau / 50.0;
      double y = sst[i];
      for (int s = 0; s < 50; ++s) {
        const double k1 = (M - 0.2 * y) - K * icing_f1_d(y);
        const double y2 = y + 0.5 * h * k1;
        const double k2 = (M - 0.2 * y2) - K * icing_f1_d(y2);
        const double y3 = y + 0.5 * h * k2;
        const double k3 = (M - 0.2 * y3) - K * icing_f1_d(y3);
        const double y4 = y + h * k3;
        y += h * ((1.0 / 6.0) * (((k1 + 2.0 * k2) + 2.0 * k3) + ((M - 0.2 * y4) - K * icing_f1_d(y4))));
      }
      td = y;
    }
    double ice = 0;
    for (int k = 0; k < number; ++k) { /* freezing fraction per height (:288-326) */
      const double rw = 6.46E-5 * wave[i] * (Vr * Vr) * exp(-0.55 * (zmin + 0.5 * k)) * v;
      double N = 0.0, err = 1.0;
      int j = 0;
      while (err >= 1.0E-5 && N >= 0 && N <= 1) {
        const double Ts = (1.0 + N) * Tf;
        const double ri =
            (0.012012012 * rw * (Ts - td) + (ha / 333000.0) * ((Ts - airtemp[i]) + ratio * (icing_f1_d(Ts) - rh[i] * icing_f1_f(airtemp[i]))));
        const double N1 = ri / rw;
        err = fabs(N1 - N);
        N = N1;
        j = j + 1;
        if (j > 1000) {
          N = 0.0;
          break;
        }
      }
      if (N < 0.0)
        N = 0.0;
      else if (N > 1.0)
        N = 1.0;
      ice += N * (rw / 890.0) * 3600.0 * 100.0;
    }
    icing[i] = (float)fabs(ice / number);
  }
  *fDefined = check_defined(nundef, n);
  return 1;
}

/* ---- MINCOG, instantiated with V = float in the reference (VI.cc:339-675) ---- */

typedef struct
{
  float Sw, Ta, ha, he, ea, RH, rw, Tsp, Lwdown, Swdown;
} ffz_t;

static float freeze_frac_zero(const ffz_t* z, float N)
{ /* VI.cc:345-361 */
  const float cw = 4000;
  const float lfs = (float)(3.33e5 * 0.7);
  const float Sb = (float)(z->Sw / (1 - N * (1 - 0.3)));
  const float Ts = (float)-54.1126 * (Sb / (1000 - Sb));
  const float es = 10 * icing_f1_f(Ts);
  const float Qc = z->ha * (Ts - z->Ta);
  const float Qe = z->he * (es - z->RH * z->ea);
  const float Qd = z->rw * cw * (Ts - z->Tsp);
  const float Lwup = kT4_f(Ts);
  const float Qr = (float)(Lwup - z->Lwdown - 0.44 * z->Swdown);
  const float ri = (1 / lfs) * (Qc + Qe + Qd + Qr);
  const float N1 = ri / z->rw;
  return N1 - N;
}

static float mincog_bisection(const ffz_t* z, float a, float b, float epsilon)
{ /* VI.cc:381-415 */
  float ffa = freeze_frac_zero(z, a);
  const float ffb = freeze_frac_zero(z, b);
  if ((ffa > 0) == (ffb > 0))
    return 0;
  int iterations = (int)log2f((b - a) / epsilon);
  if (iterations > 100)
    iterations = 100;
  float c = 0;
  int j = 0;
  for (; j < iterations; ++j) {
    c = (a + b) / 2;
    const float ffc = freeze_frac_zero(z, c);
    if (ffc == 0)
      return c;
    if ((ffc > 0) != (ffa > 0)) {
      b = c;
    } else {
      a = c;
      ffa = ffc;
    }
  }
  if (j >= 100)
    c = 0;
  return c;
}

static inline float f10mk(float t, float M, float K)
{ /* VI.cc:59-63 */
  return (M - (float)0.2 * t) - K * 10 * icing_f1_f(t);
}

static float mincog_point(float sal, float wave, float x_wind, float y_wind, float airtemp, float rh, float sst, float p, float Pw, float depth,
                          float vs, float alpha, float zmin, float zmax, int alt)
{ /* VI.cc:465-675 */
  const float v = sqrtf(x_wind * x_wind + y_wind * y_wind);
  if (v < 1 || wave < 0.1)
    return 0;

  const float c_0 = (float)(9.81 / (2 * M_PI) * Pw);
  float c = c_0;
  if (depth <= c * Pw && c_0 != 0) {
    c = 1;
    int j = 0;
    const float a = (float)(2 * M_PI * depth / Pw);
    for (; j < 1000; ++j) {
      const float c_new = (float)(c_0 * tanh((double)(a / c)));
      const float err = fabsf(c_new - c);
      c = c_new;
      if (err <= 1e-5)
        break;
    }
    if (j >= 1000)
      c = 0;
  }

  const float cos_alpha = (float)cos((double)alpha);
  const float Vr = c - vs * cos_alpha;
  const float tper = fabsf(c * Pw / Vr);
  if (tper <= 0)
    return 0;

  const float beta = alpha, sin_beta = (float)sin((double)beta);
  const float Wrx = (float)fabs(v * cos((double)beta) - vs);
  const float Wry = fabsf(v * sin_beta);
  const float Wr_inv = 1 / sqrtf(Wrx * Wrx + Wry * Wry);

  const float hax = (float)(6.0617 * pow((double)Wrx, 1.82));
  const float hay = (float)(4.8496 * pow((double)Wry, 1.8));
  const float ha = (hax + hay) / (Wrx + Wry);

  const float tdur = (float)(0.1230 + 0.7008 * fabsf(Vr * wave) / ((v < 5.f) ? 5.f : v)); /* std::max<V>(v, 5) */
  const float Nf = 1 / (4 * tper);

  const float beta_r = (float)(M_PI - asinf(v * sin_beta * Wr_inv));
  float br;
  if (beta_r <= (M_PI / 2))
    br = (float)(91 * M_PI / 180);
  else if (beta_r > (M_PI))
    br = (float)M_PI;
  else
    br = beta_r;
  const float sin_br = sinf(br);
  const float sin_beta_r_2 = sin_br * sin_br;
  const float cos_beta_r = cosf(br);
  const float cos_2_beta_r = cosf(2 * br);

  const float r0 = (float)13.18, a0 = (float)32.88, b0 = (float)6.605;
  const float a0_2 = a0 * a0, b0_2 = b0 * b0, r0_2 = r0 * r0;
  const float c0 = (float)(sqrt(2.0) * a0 * b0 * sqrtf((b0_2 - a0_2) * cos_2_beta_r + a0_2 + b0_2 - 2 * r0_2 * sin_beta_r_2));
  const float r = (r0 * 2 * b0_2 * cos_beta_r + c0) / ((b0_2 - a0_2) * cos_2_beta_r + a0_2 + b0_2);

  const float tau_const = r * Wr_inv;
  const float beta_deg = (float)(beta * (180 / M_PI));
  const float drag = (float)(-0.0046 * beta_deg + 2.1912);
  const float tau = tau_const * drag;

  const float ea = 10 * icing_f1_f(airtemp);
  const float K = (float)(0.2 * 0.622 * 2.5E6 / (p * 1005.0));
  const float M = (float)(0.2 * airtemp + K * rh * ea);

  /* droplet temperature: RK4 in the reference's 3/8-free form (:450-463) */
  float y = sst;
  {
    const float h = tau / 50, h2 = h / 2;
    for (int s = 0; s < 50; ++s) {
      const float k1 = h2 * f10mk(y, M, K);
      const float k2 = h * f10mk(y + k1, M, K);
      const float k3 = h * f10mk(y + k2 / 2, M, K);
      const float k4 = h2 * f10mk(y + k3, M, K);
      y += (k1 + k2 + k3 + k4) / 3;
    }
  }
  const float Td = y;
  const float Tsp = (float)(0.5 * (Td + sst));

  const float Vdz = (float)6.67;
  const float Vdcomp = (float)(Wrx * 0.9962 + Vdz * 0.0872);

  float lwc0;
  if (alt == 1) {
    lwc0 = (float)(6.36E-5 * wave * (Vr * Vr));
  } else {
    const float lambda = c * Pw, dl = (float)(4 * M_PI * depth / lambda);
    const float cg = (c / 2) * (1 + dl / sinhf(dl));
    const float Vgr = cg - vs * cos_alpha;
    lwc0 = (float)(9.5205E-4 * (wave * wave) * sqrtf(wave / lambda) * Vgr);
  }
  lwc0 = fabsf(lwc0);

  const float he = (float)(ha * 1738.6 / p);
  const float Ta = airtemp;
  const float Vf = (float)((1 + cos(85 * M_PI / 180)) / 2);
  const float Swdown_model = 0;
  const float eps_atm = (float)0.7;
  const float Lwdown = eps_atm * kT4_f(airtemp);
  const float Swdown = Swdown_model * Vf;

  float icing = 0;
  const float num = zmax - zmin;
  const int number = (int)(num * 2 + 1);
  for (int k = 0; k < number; ++k) {
    const float lwc = (float)(lwc0 * exp(-0.55 * (zmin + 0.5 * k)));
    const float rw = lwc * Vdcomp * Nf * tdur;
    const ffz_t z = {sal, Ta, ha, he, ea, rh, rw, Tsp, Lwdown, Swdown};
    float N = mincog_bisection(&z, (float)-0.5, (float)1.3, (float)1e-5);
    if (N < 0)
      N = 0;
    else if (1 < N)
      N = 1;
    icing += rw * N;
  }
  return fabsf(icing / number) * (float)(3600.0 * 100.0 / 890.0);
}

int fco_vesselIcingMincog(int nx, int ny, const float* sal, const float* wave, const float* x_wind, const float* y_wind, const float* airtemp,
                          const float* rh, const float* sst, const float* p, const float* Pw, const float* aice, const float* depth, float vs,
                          float alpha, float zmin, float zmax, int alt, float* icing, int* fDefined, float undef)
{ /* VI.cc:677-705 */
  if (vs < 0 || alpha < 0 || zmin < 0 || zmax < 0 || zmax < zmin || fmod(zmax - zmin, 1) != 0)
    return 0;
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (size_t i = 0; i < n; ++i) {
    if (def10(all, sal[i], wave[i], x_wind[i], y_wind[i], airtemp[i], rh[i], sst[i], p[i], aice[i], depth[i], undef) && aice[i] < 0.4 &&
        sst[i] > (-54.1126 * sal[i] / (1000 - sal[i]))) {
      icing[i] = mincog_point(sal[i], wave[i], x_wind[i], y_wind[i], airtemp[i], rh[i], sst[i], p[i], Pw[i], depth[i], vs, alpha, zmin, zmax, alt);
    } else {
      icing[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, n);
  return 1;
}

/* ================================================================== field arithmetic */

int fco_fieldOPERfield(int compute, int nx, int ny, const float* field1, const float* field2, float* fres, int* fDefined, float undef)
{ /* FC.cc:2611-2625; + - * leave the flag alone (:126-140), / recounts with b == 0 -> undef (:84-92, :161-179) */
  if (compute < 1 || compute > 4)
    return 0;
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (size_t i = 0; i < n; ++i) {
    const float a = field1[i], b = field2[i];
    if (!(all || (is_def(a, undef) && is_def(b, undef)))) {
      fres[i] = undef;
      nundef += 1;
    } else if (compute == 1)
      fres[i] = a + b;
    else if (compute == 2)
      fres[i] = a - b;
    else if (compute == 3)
      fres[i] = a * b;
    else if (b != 0)
      fres[i] = a / b;
    else {
      fres[i] = undef;
      nundef += 1;
    }
  }
  if (compute == 4)
    *fDefined = check_defined(nundef, n);
  return 1;
}

/* ================================================================== ensemble reductions
 * Sequential float accumulation in member order, exactly as the reference's inner loops. */

int fco_meanValue(int nx, int ny, const float* const* fields, int nfields, const int* fDefinedIn, float* fres, int* fDefinedOut, float undef)
{ /* FC.cc:2696-2724 */
  const size_t n = (size_t)nx * ny;
  size_t nundef = 0;
  for (size_t i = 0; i < n; ++i) {
    float sum = 0;
    int cnt = 0;
    for (int j = 0; j < nfields; ++j) {
      const float x = fields[j][i];
      if (fDefinedIn[j] == ALL_DEFINED || is_def(x, undef)) {
        cnt++;
        sum += x;
      }
    }
    if (cnt > 0)
      fres[i] = sum / cnt;
    else {
      fres[i] = undef;
      nundef += 1;
    }
  }
  *fDefinedOut = check_defined(nundef, n);
  return 1;
}

int fco_stddevValue(int nx, int ny, const float* const* fields, int nfields, const int* fDefinedIn, float* fres, int* fDefinedOut, float undef)
{ /* FC.cc:2726-2757: Welford in float, population sigma */
  const size_t n = (size_t)nx * ny;
  size_t nundef = 0;
  for (size_t i = 0; i < n; ++i) {
    int cnt = 0;
    float m = 0, m2 = 0;
    for (int j = 0; j < nfields; ++j) {
      const float x = fields[j][i];
      if (fDefinedIn[j] == ALL_DEFINED || is_def(x, undef)) {
        const float delta = x - m;
        cnt += 1;
        m += delta / cnt;
        m2 += delta * (x - m);
      }
    }
    if (cnt > 0)
      fres[i] = (float)sqrt((double)(m2 / cnt));
    else {
      fres[i] = undef;
      nundef += 1;
    }
  }
  *fDefinedOut = check_defined(nundef, n);
  return 1;
}

int fco_extremeValue(int compute, int nx, int ny, const float* const* fields, int nfields, float* fres, int* fDefined, float undef)
{ /* FC.cc:2759-2805: the running value starts at undef and takes member j whenever it still equals undef --
   * even if that member is itself undefined or NaN; an invalid compute only touches the flag */
  if (nfields <= 0)
    return 0;
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  if (compute >= 1 && compute <= 4) {
    const int want_max = (compute == 1 || compute == 3);
    const int want_index = (compute >= 3);
    for (size_t i = 0; i < n; ++i) {
      float cur = undef, idx = undef;
      for (int j = 0; j < nfields; ++j) {
        const float x = fields[j][i];
        if (cur == undef || ((all || is_def(x, undef)) && (want_max ? (cur < x) : (cur > x)))) {
          cur = x;
          idx = (float)j;
        }
      }
      fres[i] = want_index ? idx : cur;
      if (fres[i] == undef)
        nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, n);
  return 1;
}

int fco_probability(int compute, int nx, int ny, const float* const* fields, int nfields, const int* fDefinedIn, const float* limits,
                    int nlimits, float* fres, int* fDefinedOut, float undef)
{ /* FC.cc:2807-2860: members are counted by their FIELD flag (!= NONE_DEFINED), values are tested with != undef only */
  const size_t n = (size_t)nx * ny;
  const int between = (nlimits >= 2) && (compute == 3 || compute == 6);
  const int above = (nlimits >= 1) && (compute == 1 || compute == 4 || between);
  const int below = (nlimits >= 1) && (compute == 2 || compute == 5 || between);
  if (!(above || below)) {
    for (size_t i = 0; i < n; ++i)
      fres[i] = undef;
    *fDefinedOut = NONE_DEFINED;
    return 0;
  }
  const float v_above = limits[0];
  const float v_below = between ? limits[1] : limits[0];
  size_t nundef = 0;
  for (size_t i = 0; i < n; ++i) {
    float cnt = 0;
    int members = 0;
    for (int j = 0; j < nfields; ++j) {
      if (fDefinedIn[j] != NONE_DEFINED) {
        members += 1;
        const float x = fields[j][i];
        if ((x != undef) && (!above || x > v_above) && (!below || x < v_below))
          cnt += 1;
      }
    }
    if (members == 0) {
      fres[i] = undef;
      nundef += 1;
    } else if (compute < 4)
      fres[i] = (float)(cnt / (members / 100.0));
    else
      fres[i] = cnt;
  }
  *fDefinedOut = check_defined(nundef, n);
  return 1;
}

/* ================================================================================================
 * The rest of the reference's Python subset (python/py_mi_fieldcalc.cc:189-207)
 * ================================================================================================ */

/* MC.h:42, :53 */
#define K_RCP (K_R / K_CP)
#define K_CPLR (K_XLH / K_RCP)
#define K_EXL (K_EPS * K_XLH)
static const double K_MS2KNOTS = 3600.0 / 1852.0;

static inline float ms2knots(float ff)
{ /* MC.h:132-135: float * double -> float */
  return (float)(ff * K_MS2KNOTS);
}

int fco_kIndex(int nx, int ny, const float* t500, const float* t700, const float* rh700, const float* t850, const float* rh850, float p500, float p700,
               float p850, int compute, float* kfield, int* fDefined, float undef)
{ /* FC.cc:745-814 */
  if (p500 <= 0.0 || p500 >= p700 || p700 >= p850)
    return 0;
  float cvt500, cvt700, cvt850;
  if (compute == 1) {
    cvt500 = cvt700 = cvt850 = 1.f;
  } else if (compute == 2) {
    cvt500 = pidcp_from_p(p500);
    cvt700 = pidcp_from_p(p700);
    cvt850 = pidcp_from_p(p850);
  } else
    return 0;
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nu = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(t500[i], undef) && is_def(t700[i], undef) && is_def(rh700[i], undef) && is_def(t850[i], undef) && is_def(rh850[i], undef))) {
      const float rh8 = clamp_rh((float)(0.01 * rh850[i]));
      const float tc850 = cvt850 * t850[i] - K_T0;
      const float tc700 = cvt700 * t700[i] - K_T0;
      const ewt_t e850 = ewt_make(tc850), e700 = ewt_make(tc700);
      if (!(ewt_defined(e850) && ewt_defined(e700))) {
        kfield[i] = undef;
        nu += 1;
      } else {
        const float etd850 = ewt_value(e850) * rh8;
        const float tdc850 = ewt_inverse(e850, etd850);
        const float rh7 = clamp_rh((float)(0.01 * rh700[i]));
        const float etd700 = ewt_value(e700) * rh7;
        const float tdc700 = ewt_inverse(e700, etd700);
        const float tc500 = cvt500 * t500[i] - K_T0;
        kfield[i] = (tc850 + tdc850) - (tc700 - tdc700) - tc500;
      }
    } else {
      kfield[i] = undef;
      nu += 1;
    }
  }
  *fDefined = check_defined(nu, n);
  return 1;
}

int fco_ductingIndex(int nx, int ny, const float* t850, const float* rh850, float p850, int compute, float* duct, int* fDefined, float undef)
{ /* FC.cc:816-870 */
  const float bduct = (float)3.8e+5;
  if (p850 <= 0.0)
    return 0;
  float tconvert;
  if (compute == 1)
    tconvert = 1.f;
  else if (compute == 2)
    tconvert = pidcp_from_p(p850);
  else
    return 0;
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nu = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(t850[i], undef) && is_def(rh850[i], undef))) {
      const float rh = clamp_rh((float)(0.01 * rh850[i]));
      const float tk = t850[i] * tconvert;
      const ewt_t e = ewt_make(tk - K_T0);
      if (!ewt_defined(e)) {
        duct[i] = undef;
        nu += 1;
      } else {
        const float et = ewt_value(e);
        const float etd = et * rh;
        const float tdk = ewt_inverse(e, etd) + K_T0;
        duct[i] = bduct * (et / (tk * tk) - etd / (tdk * tdk));
      }
    } else {
      duct[i] = undef;
      nu += 1;
    }
  }
  *fDefined = check_defined(nu, n);
  return 1;
}

int fco_showalterIndex(int nx, int ny, const float* t500, const float* t850, const float* rh850, float p500, float p850, int compute, float* sfield,
                       int* fDefined, float undef)
{ /* FC.cc:872-971: moist adiabat by 7 adjustment iterations; an undefined input point is counted but its
   * output point is NOT written (:966-968) */
  if (p500 <= 0.0 || p500 >= p850)
    return 0;
  const float pi500 = pi_from_p(p500);
  const float pi850 = pi_from_p(p850);
  float cvt500, cvt850, dryadiabat;
  if (compute == 1) {
    cvt500 = 1.f;
    cvt850 = 1.f;
    dryadiabat = K_CP * (K_CP / pi850) * (pi500 / K_CP);
  } else if (compute == 2) {
    cvt500 = pi500 / K_CP;
    cvt850 = pi850 / K_CP;
    dryadiabat = K_CP * (pi500 / K_CP);
  } else
    return 0;
  const float cplr = K_CPLR, exl = K_EXL;
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nu = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(t500[i], undef) && is_def(t850[i], undef) && is_def(rh850[i], undef))) {
      const float tk500 = cvt500 * t500[i];
      const float tk850 = cvt850 * t850[i];
      const float rh = clamp_rh((float)(0.01 * rh850[i]));
      const ewt_t e = ewt_make(tk850 - K_T0);
      if (!ewt_defined(e)) {
        sfield[i] = undef;
        nu += 1;
      } else {
        const float etd = ewt_value(e) * rh;
        float tcl = dryadiabat * t850[i];
        float qcl = K_EPS * etd / p850;
        for (int it = 0; it < 7; ++it) {
          const ewt_t e2 = ewt_make(tcl / K_CP - K_T0);
          if (!ewt_defined(e2))
            break;
          const float esat = ewt_value(e2);
          const float qsat = K_EPS * esat / p500;
          float dq = qcl - qsat;
          const float a1 = cplr * qcl / tcl;
          const float a2 = exl / tcl;
          dq = (float)(dq / (1. + a1 * a2));
          qcl = qcl - dq;
          tcl = tcl + dq * K_XLH;
        }
        const float tx500 = tcl / K_CP;
        sfield[i] = tk500 - tx500;
      }
    } else {
      nu += 1;
    }
  }
  *fDefined = check_defined(nu, n);
  return 1;
}

int fco_boydenIndex(int nx, int ny, const float* t700, const float* z700, const float* z1000, float p700, float p1000, int compute, float* bfield,
                    int* fDefined, float undef)
{ /* FC.cc:973-1014 */
  if (compute <= 0 || compute >= 3)
    return 0;
  if (p700 <= 0.0 || p700 >= p1000)
    return 0;
  const float pi700 = K_CP * powf(p700 / K_P0, K_R / K_CP);
  const float tconv = (compute == 2) ? pi700 / K_CP : 1.f;
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nu = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(t700[i], undef) && is_def(z700[i], undef) && is_def(z1000[i], undef))) {
      const float tc700 = t700[i] * tconv - K_T0;
      bfield[i] = (float)((z700[i] - z1000[i]) / 10. - tc700 - 200.);
    } else {
      bfield[i] = undef;
      nu += 1;
    }
  }
  *fDefined = check_defined(nu, n);
  return 1;
}

int fco_sweatIndex(int nx, int ny, const float* t850, const float* t500, const float* td850, const float* td500, const float* u850, const float* v850,
                   const float* u500, const float* v500, float* sindex, int* fDefined, float undef)
{ /* FC.cc:1016-1040: the float terms are summed left to right in float, the last term is double */
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nu = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(t850[i], undef) && is_def(t500[i], undef) && is_def(td850[i], undef) && is_def(td500[i], undef) && is_def(u850[i], undef) &&
                is_def(v850[i], undef) && is_def(u500[i], undef) && is_def(v500[i], undef))) {
      const float ff850 = sqrtf(u850[i] * u850[i] + v850[i] * v850[i]);
      const float ff500 = sqrtf(u500[i] * u500[i] + v500[i] * v500[i]);
      const float sind = (u500[i] * v850[i] - v500[i] * u850[i]) / (ff850 * ff500);
      const float lhs = 32 * td850[i] + 20 * t850[i] - 40 * t500[i] - 20 * 49 + 2 * ms2knots(ff850) + ms2knots(ff500);
      sindex[i] = (float)(lhs + 125 * (sind + 0.2));
    } else {
      sindex[i] = undef;
      nu += 1;
    }
  }
  *fDefined = check_defined(nu, n);
  return 1;
}

int fco_seaSoundSpeed(int nx, int ny, const float* t, const float* s, float z_, int compute, float* soundspeed, int* fDefined, float undef)
{ /* FC.cc:1555-1602 */
  if (compute != 1 && compute != 2)
    return 0;
  const float tconv = (compute == 1) ? 0.f : K_T0;
  const double Z = fabsf(z_);
  const double Cz = 0.01635 * Z + 0.000000175 * Z * Z;
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nu = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(t[i], undef) && is_def(s[i], undef))) {
      const float T = t[i] - tconv;
      const float S = s[i];
      const double Ct = 4.565 * T - 0.0517 * T * T + 0.000221 * T * T * T;
      const double Cs = (1.338 - 0.013 * T + 0.0001 * T * T) * (S - 35.0);
      const double speed = 1449.1 + Ct + Cs + Cz;
      soundspeed[i] = (float)speed;
    } else {
      soundspeed[i] = undef;
      nu += 1;
    }
  }
  *fDefined = check_defined(nu, n);
  return 1;
}

int fco_cvtemp(int nx, int ny, const float* tinp, int compute, float* tout, int* fDefined, float undef)
{ /* FC.cc:1608-1674 (serial build: the average is a float accumulation in index order) */
  float tconvert;
  switch (compute) {
  case 1:
  case 3:
    tconvert = -K_T0;
    break;
  case 2:
  case 4:
    tconvert = +K_T0;
    break;
  default:
    return 0;
  }
  const size_t n = (size_t)((long long)nx * ny > 0 ? (long long)nx * ny : 0);
  const int all = *fDefined == ALL_DEFINED;
  if (compute == 3 || compute == 4) {
    float tavg = 0.f;
    int navg = 0;
    for (size_t i = 0; i < n; ++i)
      if (all || is_def(tinp[i], undef)) {
        tavg += tinp[i];
        navg += 1;
      }
    if (navg > 0)
      tavg /= (float)navg;
    if ((compute == 3 && tavg < K_T0 / 2.) || (compute == 4 && tavg > K_T0 / 2.)) {
      if (tout != tinp)
        for (size_t i = 0; i < n; ++i)
          tout[i] = tinp[i];
      return 1;
    }
  }
  size_t nu = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || is_def(tinp[i], undef))
      tout[i] = tinp[i] + tconvert;
    else {
      tout[i] = undef;
      nu += 1;
    }
  }
  *fDefined = check_defined(nu, n);
  return 1;
}

int fco_cvhum(int nx, int ny, const float* t, const float* huminp, const char* unit, int compute, float* humout, int* fDefined, float undef)
{ /* FC.cc:1738-1817 */
  float unit_scale = 100.f;
  if (compute == 1 && strcmp(unit, "celsius") == 0)
    compute = 2;
  if ((compute == 4 || compute == 5) && strcmp(unit, "1") == 0)
    unit_scale = 1.f;
  const size_t n = (size_t)nx * ny;
  const float tconv = (compute == 1 || compute == 2 || compute == 4) ? K_T0 : 0.f;
  const float tdconv = (compute == 1) ? K_T0 : 0.f;
  const int all = *fDefined == ALL_DEFINED;
  size_t nu = 0;
  if (compute >= 1 && compute <= 3) {
    for (size_t i = 0; i < n; ++i) {
      if (all || (is_def(t[i], undef) && is_def(huminp[i], undef))) {
        const ewt_t e = ewt_make(t[i] - tconv);
        if (!ewt_defined(e)) {
          humout[i] = undef;
          nu += 1;
        } else {
          const float et = ewt_value(e);
          const float rh = clamp_rh((float)(0.01 * huminp[i]));
          const float etd = rh * et;
          humout[i] = ewt_inverse(e, etd) + tdconv;
        }
      } else {
        humout[i] = undef;
        nu += 1;
      }
    }
  } else if (compute == 4 || compute == 5) {
    for (size_t i = 0; i < n; ++i) {
      if (all || (is_def(t[i], undef) && is_def(huminp[i], undef))) {
        const ewt_t e = ewt_make(t[i] - tconv), e2 = ewt_make(huminp[i] - tconv);
        if (!(ewt_defined(e) && ewt_defined(e2))) {
          humout[i] = undef;
          nu += 1;
        } else {
          const float rh = ewt_value(e2) / ewt_value(e);
          humout[i] = rh * unit_scale;
        }
      } else {
        humout[i] = undef;
        nu += 1;
      }
    }
  } else
    return 0;
  *fDefined = check_defined(nu, n);
  return 1;
}

int fco_abshum(int nx, int ny, const float* t, const float* rhum, float* abshumout, int* fDefined, float undef)
{ /* FC.cc:1676-1736.  <cmath> without `using namespace std`: the unqualified sqrt(v) and exp(x) with float
   * arguments bind to the C library's double functions */
  const float C = (float)2.16679, C1 = (float)-7.85951783, C2 = (float)1.84408259, C3 = (float)-11.7866497, C4 = (float)22.6807411,
              C5 = (float)-15.9618719, C6 = (float)1.80122502, Tc = (float)647.096, Pc = 220640.f;
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nu = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(t[i], undef) && is_def(rhum[i], undef))) {
      const float v = 1 - t[i] / Tc, tii = 1 / t[i];
      const float v2 = v * v, v3 = v * v2, v4 = v2 * v2, v1_5 = (float)(v * sqrt((double)v)), v3_5 = v2 * v1_5, v7_5 = v4 * v3_5;
      const float Pws = (float)(Pc * exp((double)(Tc * tii * (C1 * v + C2 * v1_5 + C3 * v3 + C4 * v3_5 + C5 * v4 + C6 * v7_5))));
      const float Pw = Pws * rhum[i];
      abshumout[i] = C * Pw * 100 * tii;
    } else {
      abshumout[i] = undef;
      nu += 1;
    }
  }
  *fDefined = check_defined(nu, n);
  return 1;
}

int fco_underCooledRain(int nx, int ny, const float* precip, const float* snow, const float* tk, float precipMin, float snowRateMax, float tcMax,
                        float* undercooled, int* fDefined, float undef)
{ /* FC.cc:2231-2264 */
  const size_t n = (size_t)nx * ny;
  const float tkMax = tcMax + K_T0;
  const int all = *fDefined == ALL_DEFINED;
  size_t nu = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(precip[i], undef) && is_def(snow[i], undef) && is_def(tk[i], undef))) {
      undercooled[i] = (precip[i] >= precipMin && tk[i] <= tkMax && snow[i] <= precip[i] * snowRateMax) ? 1.f : 0.f;
    } else {
      undercooled[i] = undef;
      nu += 1;
    }
  }
  *fDefined = check_defined(nu, n);
  return 1;
}

/* ================================================================================================
 * The rest of SURVEY.md 8f rank 1: field arithmetic, element functions, p-level siblings
 * ================================================================================================ */
static int fill_undef(int nx, int ny, float* fres, int* fDefined, float undef)
{ /* FC.cc:76-82 */
  const size_t n = (size_t)nx * ny;
  *fDefined = NONE_DEFINED;
  for (size_t i = 0; i < n; ++i)
    fres[i] = undef;
  return 1;
}

static inline float tk_rh_the(float tk, float rh, float thconv, float undef, size_t* nu)
{ /* FC.cc:269-278 */
  EWT_OR_UNDEF(e, tk - K_T0)
  return tk * thconv + ewt_value(e) * rh;
}

int fco_plevelthe(int nx, int ny, const float* t, const float* rh, float p, int compute, float* the, int* fDefined, float undef)
{ /* FC.cc:369-398 */
  if (compute != 1 && compute != 2)
    return 0;
  if (p <= 0.0)
    return 0;
  const float pidcp = pidcp_from_p(p), pi = pidcp * K_CP;
  const float cvrh = (float)(0.01 * (K_XLH / pi) * K_EPS / p);
  const float tconv = (compute == 2) ? pidcp : 1.f;
  const float thconv = 1 / pidcp;
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nu = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(t[i], undef) && is_def(rh[i], undef)))
      the[i] = tk_rh_the(t[i] * tconv, rh[i] * cvrh, thconv, undef, &nu);
    else {
      the[i] = undef;
      nu += 1;
    }
  }
  *fDefined = check_defined(nu, n);
  return 1;
}

int fco_pleveldz2tmean(int nx, int ny, const float* z1, const float* z2, float p1, float p2, int compute, float* tmean, int* fDefined, float undef)
{ /* FC.cc:466-503; binaryFunctionFieldField: flag untouched */
  if (p1 <= 0 || p2 <= 0 || p1 == p2)
    return 0;
  const float g = (float)9.8;
  const float pi1 = pi_from_p(p1), pi2 = pi_from_p(p2);
  float convert, tconvert;
  switch (compute) {
  case 1:
    convert = (float)(g * 0.5 * (pi1 + pi2) / ((pi2 - pi1) * K_CP));
    tconvert = -K_T0;
    break;
  case 2:
    convert = (float)(g * 0.5 * (pi1 + pi2) / ((pi2 - pi1) * K_CP));
    tconvert = 0.f;
    break;
  case 3:
    convert = g / (pi2 - pi1);
    tconvert = 0.f;
    break;
  default:
    return 0;
  }
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(z1[i], undef) && is_def(z2[i], undef)))
      tmean[i] = (z1[i] - z2[i]) * convert + tconvert;
    else
      tmean[i] = undef;
  }
  return 1;
}

int fco_plevelducting(int nx, int ny, const float* t, const float* h, float p, int compute, float* duct, int* fDefined, float undef)
{ /* FC.cc:597-636 */
  if (p <= 0)
    return 0;
  const float tconv = (compute % 2 == 0) ? pidcp_from_p(p) : 1.f;
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  if (compute == 1 || compute == 2) {
    for (size_t i = 0; i < n; ++i) {
      if (all || (is_def(t[i], undef) && is_def(h[i], undef)))
        duct[i] = tk_q_duct(t[i] * tconv, h[i], p);
      else
        duct[i] = undef;
    }
    return 1;
  }
  if (compute == 3 || compute == 4) {
    size_t nu = 0;
    for (size_t i = 0; i < n; ++i) {
      if (all || (is_def(t[i], undef) && is_def(h[i], undef)))
        duct[i] = tk_rh_duct(t[i] * tconv, h[i], p, undef, &nu);
      else {
        duct[i] = undef;
        nu += 1;
      }
    }
    *fDefined = check_defined(nu, n);
    return 1;
  }
  return 0;
}

int fco_vectorabs(int nx, int ny, const float* u, const float* v, float* ff, int* fDefined, float undef)
{ /* FC.cc:1819-1841 */
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nu = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(u[i], undef) && is_def(v[i], undef)))
      ff[i] = sqrtf(u[i] * u[i] + v[i] * v[i]);
    else {
      ff[i] = undef;
      nu += 1;
    }
  }
  *fDefined = check_defined(nu, n);
  return 1;
}

int fco_pressure2FlightLevel(int nx, int ny, const float* pressure, float* flightlevel, int* fDefined, float undef)
{ /* FC.cc:2311-2349, tables MC.h:87-89 */
  static const float pT[16] = {1000, 925, 850, 800, 700, 500, 400, 300, 250, 200, 150, 100, 70, 50, 30, 10};
  static const float fT[16] = {5, 25, 50, 65, 100, 185, 235, 300, 340, 385, 445, 530, 605, 675, 780, 1020};
  const int nTab = 15;
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nu = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || is_def(pressure[i], undef)) {
      float p = pressure[i];
      if (p > pT[0])
        p = pT[0];
      if (p < pT[nTab])
        p = pT[nTab];
      int k = 1;
      while (k < nTab && pT[k] > p)
        k++;
      const float ratio = (p - pT[k - 1]) / (pT[k] - pT[k - 1]);
      flightlevel[i] = fT[k - 1] + (fT[k] - fT[k - 1]) * ratio;
    } else {
      flightlevel[i] = undef;
      nu += 1;
    }
  }
  *fDefined = check_defined(nu, n);
  return 1;
}

int fco_values2classes(int nx, int ny, const float* fvalue, float* fclass, const float* values, int nvalues_in, int* fDefined, float undef)
{ /* FC.cc:2462-2499 */
  if (nvalues_in < 2)
    return 0;
  const int nvalues = nvalues_in - 2;
  const float fmin = values[0], fmax = values[nvalues + 1];
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nu = 0;
  for (size_t i = 0; i < n; ++i) {
    if ((all || is_def(fvalue[i], undef)) && fvalue[i] >= fmin && fvalue[i] < fmax) {
      int j = 1;
      while (j < nvalues && values[j] < fvalue[i])
        j++;
      fclass[i] = (float)(j - 1);
    } else {
      fclass[i] = undef;
      nu += 1;
    }
  }
  *fDefined = check_defined(nu, n);
  return 1;
}

/* unaryFunctionField / binaryFunctionFieldField, FC.cc:94-140: undefined -> undef, flag untouched */
#define UNARY_LOOP(EXPR)                                                                                                                             \
  do {                                                                                                                                               \
    const size_t n = (size_t)nx * ny;                                                                                                                \
    const int all = *fDefined == ALL_DEFINED;                                                                                                        \
    for (size_t i = 0; i < n; ++i) {                                                                                                                 \
      const float a = field[i];                                                                                                                      \
      fres[i] = (all || is_def(a, undef)) ? (EXPR) : undef;                                                                                          \
    }                                                                                                                                                \
  } while (0)

static inline float std_min(float a, float b)
{ /* std::min(a, b) */
  return (b < a) ? b : a;
}
static inline float std_max(float a, float b)
{ /* std::max(a, b) */
  return (a < b) ? b : a;
}

int fco_minvalueFields(int nx, int ny, const float* field1, const float* field2, float* fres, int* fDefined, float undef)
{ /* FC.cc:2501-2505 */
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  for (size_t i = 0; i < n; ++i)
    fres[i] = (all || (is_def(field1[i], undef) && is_def(field2[i], undef))) ? std_min(field1[i], field2[i]) : undef;
  return 1;
}

int fco_maxvalueFields(int nx, int ny, const float* field1, const float* field2, float* fres, int* fDefined, float undef)
{ /* FC.cc:2516-2520 */
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  for (size_t i = 0; i < n; ++i)
    fres[i] = (all || (is_def(field1[i], undef) && is_def(field2[i], undef))) ? std_max(field1[i], field2[i]) : undef;
  return 1;
}

int fco_minvalueFieldConst(int nx, int ny, const float* field, float value, float* fres, int* fDefined, float undef)
{ /* FC.cc:2507-2514 */
  if (value == undef)
    return fill_undef(nx, ny, fres, fDefined, undef);
  UNARY_LOOP(std_min(a, value));
  return 1;
}

int fco_maxvalueFieldConst(int nx, int ny, const float* field, float value, float* fres, int* fDefined, float undef)
{ /* FC.cc:2522-2529 */
  if (value == undef)
    return fill_undef(nx, ny, fres, fDefined, undef);
  UNARY_LOOP(std_max(a, value));
  return 1;
}

int fco_absvalueField(int nx, int ny, const float* field, float* fres, int* fDefined, float undef)
{ /* FC.cc:2531-2534 */
  UNARY_LOOP(fabsf(a));
  return 1;
}

int fco_log10Field(int nx, int ny, const float* field, float* fres, int* fDefined, float undef)
{ /* FC.cc:2536-2539 */
  UNARY_LOOP(log10f(a));
  return 1;
}

int fco_pow10Field(int nx, int ny, const float* field, float* fres, int* fDefined, float undef)
{ /* FC.cc:2541-2544; math_util.h:121-125: std::pow(10, float) is the double pow */
  UNARY_LOOP((float)pow(10.0, (double)a));
  return 1;
}

int fco_logField(int nx, int ny, const float* field, float* fres, int* fDefined, float undef)
{ /* FC.cc:2546-2549 */
  UNARY_LOOP(logf(a));
  return 1;
}

int fco_expField(int nx, int ny, const float* field, float* fres, int* fDefined, float undef)
{ /* FC.cc:2551-2554 */
  UNARY_LOOP(expf(a));
  return 1;
}

int fco_powerField(int nx, int ny, const float* field, float value, float* fres, int* fDefined, float undef)
{ /* FC.cc:2556-2563 */
  if (value == undef)
    return fill_undef(nx, ny, fres, fDefined, undef);
  UNARY_LOOP(powf(a, value));
  return 1;
}

int fco_replaceUndefined(int nx, int ny, const float* field, float value, float* fres, int* fDefined, float undef)
{ /* FC.cc:2565-2587: tests `== undef` only (a NaN is kept) */
  const size_t n = (size_t)nx * ny;
  if (value == undef || *fDefined == ALL_DEFINED) {
    if (fres != field)
      memcpy(fres, field, sizeof(float) * n);
    return 1;
  }
  if (*fDefined == NONE_DEFINED) {
    for (size_t i = 0; i < n; ++i)
      fres[i] = value;
  } else {
    for (size_t i = 0; i < n; ++i)
      fres[i] = (field[i] == undef) ? value : field[i];
  }
  *fDefined = ALL_DEFINED;
  return 1;
}

int fco_replaceDefined(int nx, int ny, const float* field, float value, float* fres, int* fDefined, float undef)
{ /* FC.cc:2589-2609 */
  const size_t n = (size_t)nx * ny;
  if (value == undef || *fDefined == NONE_DEFINED) {
    for (size_t i = 0; i < n; ++i)
      fres[i] = undef;
    *fDefined = NONE_DEFINED;
    return 1;
  }
  if (*fDefined == ALL_DEFINED) {
    for (size_t i = 0; i < n; ++i)
      fres[i] = value;
  } else {
    for (size_t i = 0; i < n; ++i)
      fres[i] = (field[i] != undef) ? value : field[i];
  }
  *fDefined = ALL_DEFINED;
  return 1;
}

int fco_fieldOPERconstant(int compute, int nx, int ny, const float* field, float value, float* fres, int* fDefined, float undef)
{ /* FC.cc:2627-2645 */
  if ((value == undef) || (compute == 4 && value == 0))
    return fill_undef(nx, ny, fres, fDefined, undef);
  switch (compute) {
  case 1:
    UNARY_LOOP(a + value);
    return 1;
  case 2:
    UNARY_LOOP(a - value);
    return 1;
  case 3:
    UNARY_LOOP(a * value);
    return 1;
  case 4:
    UNARY_LOOP(a / value);
    return 1;
  default:
    return 0;
  }
}

int fco_constantOPERfield(int compute, int nx, int ny, float value, const float* field, float* fres, int* fDefined, float undef)
{ /* FC.cc:2647-2669 */
  if (value == undef)
    return fill_undef(nx, ny, fres, fDefined, undef);
  switch (compute) {
  case 1:
    UNARY_LOOP(value + a);
    return 1;
  case 2:
    UNARY_LOOP(value - a);
    return 1;
  case 3:
    UNARY_LOOP(value * a);
    return 1;
  case 4: {
    const size_t n = (size_t)nx * ny;
    const int all = *fDefined == ALL_DEFINED;
    size_t nu = 0;
    for (size_t i = 0; i < n; ++i) {
      if ((all || is_def(field[i], undef)) && field[i] != 0)
        fres[i] = value / field[i];
      else {
        fres[i] = undef;
        nu += 1;
      }
    }
    *fDefined = check_defined(nu, n);
    return 1;
  }
  default:
    return 0;
  }
}

int fco_sumFields(int nx, int ny, const float* const* fields, int nfields, float* fres, int* fDefined, float undef)
{ /* FC.cc:2671-2694 */
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nu = 0;
  for (size_t i = 0; i < n; ++i) {
    fres[i] = 0;
    for (int j = 0; j < nfields; ++j) {
      if (all || is_def(fields[j][i], undef))
        fres[i] += fields[j][i];
      else {
        fres[i] = undef;
        nu += 1;
        break;
      }
    }
  }
  *fDefined = check_defined(nu, n);
  return 1;
}

int fco_snow_in_cm(int nx, int ny, const float* snow_water, const float* tk2m, const float* td2m, float* snow_cm, int* fDefined, float undef)
{ /* FC.cc:3063-3118: exp() on a double argument, every line a double expression rounded to float */
  const size_t n = (size_t)nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nu = 0;
  for (size_t i = 0; i < n; ++i) {
    if (all || (is_def(snow_water[i], undef) && is_def(tk2m[i], undef) && is_def(td2m[i], undef))) {
      if (snow_water[i] <= 0.) {
        snow_cm[i] = 0.f;
        continue;
      }
      const float t = (float)((tk2m[i] + td2m[i]) / 2.);
      const float logit_t = (float)((1 - exp((t - 274.3) * 3.5)) / (1 + exp((t - 274.3) * 3.5)));
      const float mm2cm_t = (float)(0.13 / (0.02 + 0.1 * ((t - 252.0) / 20.0) * ((t - 252.0) / 20.0)));
      const float fac = logit_t * mm2cm_t;
      if (fac <= 1.)
        snow_cm[i] = snow_water[i];
      else
        snow_cm[i] = snow_water[i] * fac;
    } else {
      snow_cm[i] = undef;
      nu += 1;
    }
  }
  *fDefined = check_defined(nu, n);
  return 1;
}

/* ================================================================================================
 * Geostrophic stencil siblings
 * ================================================================================================ */
int fco_plevelgwind_xcomp(int nx, int ny, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis, float* ug, int* fDefined,
                          float undef)
{ /* FC.cc:638-672: n_undefined += 1 for EVERY point (:664) -> the flag is always NONE_DEFINED */
  (void)xmapr;
  if (nx < 3 || ny < 3)
    return 0;
  const float g = (float)9.8;
  const int n = nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (int i = nx; i < n - nx; ++i) {
    if (DEF4(z[i - nx], z[i - 1], z[i + 1], z[i + nx]))
      ug[i] = (float)(-0.5 * ymapr[i] * (z[i + nx] - z[i - nx]) * g / fcoriolis[i]);
    else
      ug[i] = undef;
    nundef += 1;
  }
  *fDefined = check_defined(nundef, (size_t)(n - 2 * nx));
  fill_edges(nx, ny, ug);
  return 1;
}

int fco_plevelgwind_ycomp(int nx, int ny, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis, float* vg, int* fDefined,
                          float undef)
{ /* FC.cc:674-706.  The reference has no nx, ny >= 3 test here and then reads and writes out of bounds in
   * fillEdges (undefined behaviour); this restatement and the product reject such grids like xcomp does. */
  (void)ymapr;
  if (nx < 3 || ny < 3)
    return 0;
  const float g = (float)9.8;
  const int n = nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (int i = nx; i < n - nx; ++i) {
    if (DEF4(z[i - nx], z[i - 1], z[i + 1], z[i + nx]))
      vg[i] = (float)(0.5 * xmapr[i] * (z[i + 1] - z[i - 1]) * g / fcoriolis[i]);
    else {
      vg[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, (size_t)(n - 2 * nx));
  fill_edges(nx, ny, vg);
  return 1;
}

int fco_plevelgvort(int nx, int ny, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis, float* gvort, int* fDefined,
                    float undef)
{ /* FC.cc:708-743 */
  const float g = (float)9.8;
  const float g4 = (float)(g * 4.);
  if (nx < 3 || ny < 3)
    return 0;
  const int n = nx * ny;
  const int all = *fDefined == ALL_DEFINED;
  size_t nundef = 0;
  for (int i = nx; i < n - nx; ++i) {
    if (all || (is_def(z[i - nx], undef) && is_def(z[i - 1], undef) && is_def(z[i], undef) && is_def(z[i + 1], undef) && is_def(z[i + nx], undef)))
      gvort[i] = (float)((0.25 * xmapr[i] * xmapr[i] * (z[i - 1] - 2. * z[i] + z[i + 1]) + 0.25 * ymapr[i] * ymapr[i] * (z[i - nx] - 2. * z[i] + z[i + nx])) *
                         g4 / fcoriolis[i]);
    else {
      gvort[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, (size_t)(n - 2 * nx));
  fill_edges(nx, ny, gvort);
  return 1;
}

int fco_plevelqvector(int nx, int ny, const float* z, const float* t, const float* xmapr, const float* ymapr, const float* fcoriolis, float p, int compute,
                      float* qcomp, int* fDefined, float undef)
{ /* FC.cc:505-595 */
  if (p <= 0.0)
    return 0;
  if (nx < 3 || ny < 3)
    return 0;
  float tscale;
  switch (compute) {
  case 1:
  case 3:
    tscale = 1.0f;
    break;
  case 2:
  case 4: {
    const float pi = K_CP * powf(p / K_P0, K_R / K_CP);
    tscale = pi / K_CP;
    break;
  }
  default:
    return 0;
  }
  const int n = nx * ny;
  float* ug = (float*)malloc(sizeof(float) * (size_t)n);
  float* vg = (float*)malloc(sizeof(float) * (size_t)n);
  if (!fco_plevelgwind_xcomp(nx, ny, z, xmapr, ymapr, fcoriolis, ug, fDefined, undef) ||
      !fco_plevelgwind_ycomp(nx, ny, z, xmapr, ymapr, fcoriolis, vg, fDefined, undef)) {
    free(ug);
    free(vg);
    return 0;
  }
  const float c = (float)(-K_R / (p * 100.));
  size_t nundef = 0;
  for (int i = nx; i < n - nx; ++i) {
    if (ug[i - nx] != undef && ug[i - 1] != undef && ug[i + 1] != undef && ug[i + nx] != undef && vg[i - nx] != undef && vg[i - 1] != undef &&
        vg[i + 1] != undef && vg[i + nx] != undef && t[i - nx] != undef && t[i - 1] != undef && t[i + 1] != undef && t[i + nx] != undef) {
      const float dtdx = (float)(0.5 * xmapr[i] * tscale * (t[i + 1] - t[i - 1]));
      const float dtdy = (float)(0.5 * ymapr[i] * tscale * (t[i + nx] - t[i - nx]));
      if (compute < 3) {
        const float dugdx = (float)(0.5 * xmapr[i] * (ug[i + 1] - ug[i - 1]));
        const float dvgdx = (float)(0.5 * xmapr[i] * (vg[i + 1] - vg[i - 1]));
        qcomp[i] = c * (dugdx * dtdx + dvgdx * dtdy);
      } else {
        const float dugdy = (float)(0.5 * ymapr[i] * (ug[i + nx] - ug[i - nx]));
        const float dvgdy = (float)(0.5 * ymapr[i] * (vg[i + nx] - vg[i - nx]));
        qcomp[i] = c * (dugdy * dtdx + dvgdy * dtdy);
      }
    } else {
      qcomp[i] = undef;
      nundef += 1;
    }
  }
  *fDefined = check_defined(nundef, (size_t)(n - 2 * nx));
  fill_edges(nx, ny, qcomp);
  free(ug);
  free(vg);
  return 1;
}

/* ================================================================================================
 * Neighbourhood functions
 * ================================================================================================ */
static void neighbour_border_undef(int nx, int ny, int range, float* fres, float undef)
{ /* FC.cc:2929-2949 / 2989-3008 */
  for (int j = 0; j < range; j++)
    for (int i = 0; i < nx; i++)
      fres[i + j * nx] = undef;
  for (int j = range; j < ny - range; j++) {
    for (int i = 0; i < range; i++)
      fres[i + j * nx] = undef;
    for (int i = nx - range; i < nx; i++)
      fres[i + j * nx] = undef;
  }
  for (int j = ny - range; j < ny; j++)
    for (int i = 0; i < nx; i++)
      fres[i + j * nx] = undef;
}

int fco_neighbourProbFunctions(int nx, int ny, const float* field, const float* constants, int nconstants, int compute, float* fres, int* fDefined,
                               float undef)
{ /* FC.cc:2862-2953.  Outside the reference's defined behaviour (it indexes out of bounds): range < 0 or
   * range > min(nx, ny) -- rejected here and in the product. */
  if (*fDefined != ALL_DEFINED)
    return 0;
  if (nconstants < 2)
    return 0;
  const int fsize = nx * ny;
  const int limit = (int)constants[0];
  const int range = (int)constants[1];
  if (range < 0 || range > nx || range > ny)
    return 0;
  if (compute == 5) {
    for (int i = 0; i < fsize; i++)
      fres[i] = field[i] > limit ? 1 : 0;
  } else if (compute == 6) {
    for (int i = 0; i < fsize; i++)
      fres[i] = field[i] < limit ? 1 : 0;
  }
  if (range == 0)
    return 1;
  float* tmp = (float*)malloc(sizeof(float) * (size_t)fsize);
  for (int i = 0; i < nx; i++) {
    tmp[i] = fres[i];
    for (int j = 1; j < ny; j++)
      tmp[i + j * nx] = fres[i + j * nx] + tmp[i + (j - 1) * nx];
  }
  for (int j = 0; j < ny; j++)
    for (int i = 1; i < nx; i++)
      tmp[i + j * nx] += tmp[(i - 1) + j * nx];
  const int N = (2 * range + 1) * (2 * range + 1);
  for (int i = range; i < nx - range; i++) {
    const int imax = i + range;
    for (int j = range; j < ny - range; j++) {
      const int jmax = j + range;
      fres[i + j * nx] = tmp[imax + jmax * nx];
      if (i > range) {
        fres[i + j * nx] -= tmp[i - range - 1 + jmax * nx];
        if (j > range)
          fres[i + j * nx] += tmp[i - range - 1 + (j - range - 1) * nx] - tmp[imax + (j - range - 1) * nx];
      } else if (j > range) {
        fres[i + j * nx] -= tmp[imax + (j - range - 1) * nx];
      }
      fres[i + j * nx] /= N;
    }
  }
  free(tmp);
  *fDefined = SOME_DEFINED;
  neighbour_border_undef(nx, ny, range, fres, undef);
  return 1;
}

static int cmp_float(const void* a, const void* b)
{
  const float x = *(const float*)a, y = *(const float*)b;
  return (x > y) - (x < y);
}

int fco_neighbourFunctions(int nx, int ny, const float* field, const float* constants, int nconstants, int compute, float* fres, int* fDefined,
                           float undef)
{ /* FC.cc:2955-3061.  Outside the reference's defined behaviour: step / 2 > range (it paints outside the rows)
   * and a percentile index beyond the window (it reads past its vector) -- rejected here and in the product. */
  if (*fDefined != ALL_DEFINED)
    return 0;
  if (nconstants < 1 || (nconstants < 2 && compute > 3))
    return 0;
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
  if (range > nx || range > ny || range < 1)
    return 0;
  if (step < 1)
    return 0;
  const float ngridp = (float)((2 * range + 1) * (2 * range + 1));
  const int ii = (int)(ngridp * limit / 100);
  if (step / 2 > range)
    return 0;
  if (compute == 4 && (ii < 0 || ii >= (2 * range + 1) * (2 * range + 1)))
    return 0;
  *fDefined = SOME_DEFINED;
  neighbour_border_undef(nx, ny, range, fres, undef);
  const int nwin = (2 * range + 1) * (2 * range + 1);
  float* values = (float*)malloc(sizeof(float) * (size_t)nwin);
  for (int j = range; j < ny - range; j += step) {
    for (int i = range; i < nx - range; i += step) {
      int nv = 0;
      float value = 0.0f;
      if (compute == 2 || compute == 3)
        value = field[(i - range) + (j - range) * nx];
      for (int k = j - range; k < j + range + 1; k++) {
        for (int jj = i - range; jj < i + range + 1; jj++) {
          const float thisvalue = field[jj + k * nx];
          if (compute == 1)
            value += thisvalue;
          if ((compute == 2 && thisvalue > value) || (compute == 3 && thisvalue < value))
            value = thisvalue;
          else if (compute == 4)
            values[nv++] = thisvalue;
          if ((compute == 5 && thisvalue > limit) || (compute == 6 && thisvalue < limit))
            value++;
        }
      }
      if (compute == 4) {
        qsort(values, (size_t)nv, sizeof(float), cmp_float);
        value = values[ii];
      }
      if (compute == 1 || compute > 4)
        value /= ngridp;
      for (int l = j - (step - 1) / 2; l < j + step / 2 + 1; l++)
        for (int k = i - (step - 1) / 2; k < i + step / 2 + 1; k++)
          fres[k + l * nx] = value;
    }
  }
  free(values);
  return 1;
}
