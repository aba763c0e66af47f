// test_fieldcalc_api.cc -- the reference's own known-answer tests for the hot path, restated
// against the drop-in C++ API (include/mi_fieldcalc/FieldCalculations.h).
//
// Source of the vectors: /root/reference/test/FieldCalculationsTest.cc -- absHum :56-68, XLevelHum :70-143,
// ALevelTempPerformance :145-170, XOperX :172-223, Probability :225-283, Probability12 :285-305, Neighbour :307-451,
// ReplaceDefined :453-482, ReplaceUndefined :484-513 -- and test/MetConstantsTest.cc :37-113 (ICAO standard atmosphere).
// gtest is not installed, so a twenty-line checker stands in for it.
//
// The same source links against EITHER implementation of the API:
//   - libmi-fieldcalc.so.0 from this repository (CUDA; pytest -m gpu), or
//   - oracle/_ref/libfcref.so, the unmodified reference (CPU; validates this test itself).
#include "mi_fieldcalc/FieldCalculations.h"
#include "mi_fieldcalc/MetConstants.h"

#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

namespace fc = miutil::fieldcalc;
using miutil::ALL_DEFINED;
using miutil::NONE_DEFINED;
using miutil::SOME_DEFINED;
using miutil::ValuesDefined;

static int g_checks = 0, g_failed = 0;

#define CHECK(cond, ...)                                                                                                                             \
  do {                                                                                                                                               \
    ++g_checks;                                                                                                                                      \
    if (!(cond)) {                                                                                                                                   \
      ++g_failed;                                                                                                                                    \
      std::printf("FAIL %s:%d  %s  ", __FILE__, __LINE__, #cond);                                                                                    \
      std::printf(__VA_ARGS__);                                                                                                                      \
      std::printf("\n");                                                                                                                             \
    }                                                                                                                                                \
  } while (0)

static bool near(float expect, float got, float tol)
{
  return std::fabs(expect - got) <= tol;
}

// gtest's EXPECT_FLOAT_EQ: within 4 units in the last place
static bool float_eq_4ulp(float a, float b)
{
  if (a == b)
    return true;
  if (std::isnan(a) || std::isnan(b))
    return false;
  union
  {
    float f;
    int i;
  } ua = {a}, ub = {b};
  if ((ua.i < 0) != (ub.i < 0))
    return false;
  const int d = ua.i - ub.i;
  return d >= -4 && d <= 4;
}

// ---- XLevelHum: eight known-answer rows through alevelhum / hlevelhum / plevelhum -------------------
static void test_xlevelhum()
{
  const float UNDEF = 12356789, T0 = 273.15f;
  struct Row
  {
    int c_ah, c_p; // compute for a/h-level and for p-level (numbers >= 5 are permuted between the two)
    float t, hum, p, expect, tol;
  };
  const Row rows[] = {
      {1, 1, 30.68f + T0, .025f, 1013, 91.9f, 0.1f},   {2, 2, 302.71f, .025f, 1013, 91.9f, 0.1f},
      {3, 3, 30.68f + T0, 55, 1013, 0.014963f, 1e-6f}, {4, 4, 302.71f, 55, 1013, 0.014963f, 1e-6f},
      {5, 7, 30.68f + T0, .015f, 1013, 20.6f, 0.1f},   {6, 8, 302.71f, .015f, 1013, 20.6f, 0.1f},
      {7, 5, 30.68f + T0, 55, 1013, 20.6f, 0.1f},      {8, 6, 302.71f, 55, 1013, 20.6f, 0.1f},
  };
  const float alevel = 0, blevel = 1; // p = ps
  for (const Row& r : rows) {
    for (ValuesDefined in_flag : {ALL_DEFINED, SOME_DEFINED}) {
      float out = 2 * UNDEF;
      ValuesDefined f = in_flag;
      CHECK(fc::alevelhum(1, 1, &r.t, &r.hum, &r.p, "celsius", r.c_ah, &out, f, UNDEF), "alevelhum c=%d", r.c_ah);
      CHECK(near(r.expect, out, r.tol), "alevelhum c=%d got %g want %g", r.c_ah, out, r.expect);
      CHECK(f == ALL_DEFINED, "alevelhum c=%d flag %d", r.c_ah, (int)f);

      out = 2 * UNDEF;
      f = in_flag;
      CHECK(fc::hlevelhum(1, 1, &r.t, &r.hum, &r.p, alevel, blevel, "celsius", r.c_ah, &out, f, UNDEF), "hlevelhum c=%d", r.c_ah);
      CHECK(near(r.expect, out, r.tol), "hlevelhum c=%d got %g want %g", r.c_ah, out, r.expect);
      CHECK(f == ALL_DEFINED, "hlevelhum c=%d flag %d", r.c_ah, (int)f);

      out = 2 * UNDEF;
      f = in_flag;
      CHECK(fc::plevelhum(1, 1, &r.t, &r.hum, r.p, "celsius", r.c_p, &out, f, UNDEF), "plevelhum c=%d", r.c_p);
      CHECK(near(r.expect, out, r.tol), "plevelhum c=%d got %g want %g", r.c_p, out, r.expect);
      CHECK(f == ALL_DEFINED, "plevelhum c=%d flag %d", r.c_p, (int)f);
    }
    if (r.c_ah < 5)
      continue;
    // dew point in Kelvin when the unit says so
    float out = 2 * UNDEF;
    ValuesDefined f = ALL_DEFINED;
    CHECK(fc::alevelhum(1, 1, &r.t, &r.hum, &r.p, "kelvin", r.c_ah, &out, f, UNDEF), "alevelhum K c=%d", r.c_ah);
    CHECK(near(r.expect + T0, out, r.tol), "alevelhum K c=%d got %g", r.c_ah, out);
    f = ALL_DEFINED;
    CHECK(fc::hlevelhum(1, 1, &r.t, &r.hum, &r.p, alevel, blevel, "kelvin", r.c_ah, &out, f, UNDEF), "hlevelhum K c=%d", r.c_ah);
    CHECK(near(r.expect + T0, out, r.tol), "hlevelhum K c=%d got %g", r.c_ah, out);
    f = ALL_DEFINED;
    CHECK(fc::plevelhum(1, 1, &r.t, &r.hum, r.p, "kelvin", r.c_p, &out, f, UNDEF), "plevelhum K c=%d", r.c_p);
    CHECK(near(r.expect + T0, out, r.tol), "plevelhum K c=%d got %g", r.c_p, out);
    CHECK(f == ALL_DEFINED, "plevelhum K flag");
  }
}

// ---- ALevelTempPerformance: T -> theta on 719*929 points, 4 ulp ---------------------------------------
static void test_aleveltemp_large()
{
  const float UNDEF = 1e30f, T0 = 273.15f;
  const int N = 719 * 929;
  const float F = 0.00001f;
  const float p0inv = (float)(1. / 1000.f), kappa = 287.f / 1004.f;
  std::vector<float> tk(N), p(N), th(N, 2 * UNDEF);
  for (int i = 0; i < N; ++i) {
    tk[i] = 20 + (i * F) + T0;
    p[i] = 1005 + (i * F);
  }
  ValuesDefined f = ALL_DEFINED;
  CHECK(fc::aleveltemp(1, N, tk.data(), p.data(), "kelvin", 3, th.data(), f, UNDEF), "aleveltemp c=3");
  int bad = 0;
  for (int i = 0; i < N; ++i)
    if (!float_eq_4ulp(tk[i] / powf(p[i] * p0inv, kappa), th[i]))
      ++bad;
  CHECK(bad == 0, "%d of %d points differ by more than 4 ulp", bad, N);
  CHECK(f == ALL_DEFINED, "flag %d", (int)f);
}

// ---- XOperX: field (+ - * /) field incl. division by zero ---------------------------------------------
static void test_field_oper_field()
{
  const float UNDEF = 12356789;
  struct Row
  {
    int c;
    float a, b;
    ValuesDefined in;
    float expect;
  };
  const Row rows[] = {{1, 1, 3, ALL_DEFINED, 4},     {1, 1, 3, SOME_DEFINED, 4},     {2, 1, 3, ALL_DEFINED, -2},   {2, 1, 3, SOME_DEFINED, -2},
                      {3, 1.5f, 3, ALL_DEFINED, 4.5f}, {3, 1.5f, 3, SOME_DEFINED, 4.5f}, {4, 3, 1.5f, ALL_DEFINED, 2}, {4, 3, 1.5f, SOME_DEFINED, 2},
                      {4, 3, 0, ALL_DEFINED, UNDEF}, {4, 3, 0, SOME_DEFINED, UNDEF}};
  for (const Row& r : rows) {
    float out = 2 * UNDEF;
    ValuesDefined f = r.in;
    CHECK(fc::fieldOPERfield(r.c, 1, 1, &r.a, &r.b, &out, f, UNDEF), "fieldOPERfield c=%d", r.c);
    CHECK((r.expect == UNDEF) == (f == NONE_DEFINED), "fieldOPERfield c=%d a=%g b=%g flag %d", r.c, r.a, r.b, (int)f);
    CHECK(near(r.expect, out, 1e-6f), "fieldOPERfield c=%d a=%g b=%g got %g want %g", r.c, r.a, r.b, out, r.expect);

    out = 2 * UNDEF;
    f = r.in;
    CHECK(fc::fieldOPERconstant(r.c, 1, 1, &r.a, r.b, &out, f, UNDEF), "fieldOPERconstant c=%d", r.c);
    CHECK((r.expect == UNDEF) == (f == NONE_DEFINED), "fieldOPERconstant c=%d a=%g b=%g flag %d", r.c, r.a, r.b, (int)f);
    CHECK(near(r.expect, out, 1e-6f), "fieldOPERconstant c=%d a=%g b=%g got %g want %g", r.c, r.a, r.b, out, r.expect);

    out = 2 * UNDEF;
    f = r.in;
    CHECK(fc::constantOPERfield(r.c, 1, 1, r.a, &r.b, &out, f, UNDEF), "constantOPERfield c=%d", r.c);
    CHECK((r.expect == UNDEF) == (f == NONE_DEFINED), "constantOPERfield c=%d a=%g b=%g flag %d", r.c, r.a, r.b, (int)f);
    CHECK(near(r.expect, out, 1e-6f), "constantOPERfield c=%d a=%g b=%g got %g want %g", r.c, r.a, r.b, out, r.expect);
  }
}

// ---- absHum ---------------------------------------------------------------------------------------------
static void test_abshum()
{
  const float UNDEF = 12356789;
  float out = 2 * UNDEF, t = 293.16f, rh = 0.8f;
  ValuesDefined f = ALL_DEFINED;
  CHECK(fc::abshum(1, 1, &t, &rh, &out, f, UNDEF), "abshum");
  CHECK(near(13.82f, out, 0.1f), "abshum got %g", out);
  CHECK(f == ALL_DEFINED, "abshum flag %d", (int)f);
}

// ---- Neighbour ------------------------------------------------------------------------------------------
static void test_neighbour()
{
  const float UNDEF = 123456;
  const int NX = 10, NY = 10;
  float in[NX * NY], out[NX * NY], out2[NX * NY];
  std::vector<float> c;
  ValuesDefined f = ALL_DEFINED, f2 = ALL_DEFINED;
  for (int i = 0; i < NX * NY; ++i)
    in[i] = 0;

  c = {float(NX + 1)};
  CHECK(!fc::neighbourFunctions(NX, NY, in, c, 2, out, f, UNDEF), "range > nx must be rejected");
  c = {float(NX), 0};
  CHECK(!fc::neighbourFunctions(NX, NY, in, c, 2, out, f, UNDEF), "step < 1 must be rejected");

  in[16] = 6;
  for (int i = 0; i < NX * NY; ++i)
    out[i] = UNDEF / 2;
  c = {3, 3};
  CHECK(fc::neighbourFunctions(NX, NY, in, c, 2, out, f, UNDEF), "max r3 s3");
  CHECK(f == SOME_DEFINED, "flag %d", (int)f);
  for (int i = 0; i < NX; i++)
    for (int j = 0; j < NY; j++) {
      const float e = (i < 2 || i >= NX - 2 || j < 2 || j >= NY - 2) ? UNDEF : (j < 5 ? 6.f : 0.f);
      CHECK(e == out[i + j * NX], "max i=%d j=%d got %g want %g", i, j, out[i + j * NX], e);
    }

  auto four = [&]() {
    for (int i = 0; i < NX * NY; ++i) {
      in[i] = 0;
      out[i] = out2[i] = UNDEF / 2;
    }
    in[25] = in[26] = in[35] = in[36] = 6;
    f = f2 = ALL_DEFINED;
  };
  four();
  c = {90, 2, 1};
  CHECK(fc::neighbourFunctions(NX, NY, in, c, 4, out, f, UNDEF), "percentile");
  CHECK(f == SOME_DEFINED, "flag %d", (int)f);
  for (int i = 0; i < NX; i++)
    for (int j = 0; j < NY; j++) {
      const float e = (i < 2 || i >= NX - 2 || j < 2 || j >= NY - 2) ? UNDEF : ((i > 3 && i < 8 && j > 1 && j < 5) ? 6.f : 0.f);
      CHECK(e == out[i + j * NX], "pct i=%d j=%d got %g want %g", i, j, out[i + j * NX], e);
    }

  four();
  c = {5, 2, 1};
  CHECK(fc::neighbourFunctions(NX, NY, in, c, 5, out, f, UNDEF), "prob above");
  CHECK(fc::neighbourProbFunctions(NX, NY, in, c, 5, out2, f2, UNDEF), "prob above (SAT)");
  CHECK(f == SOME_DEFINED && f2 == SOME_DEFINED, "flags %d %d", (int)f, (int)f2);
  for (int i = 0; i < NX; i++)
    for (int j = 0; j < NY; j++) {
      CHECK(float_eq_4ulp(out[i + j * NX], out2[i + j * NX]), "above i=%d j=%d %g vs %g", i, j, out[i + j * NX], out2[i + j * NX]);
      float e;
      if (i < 2 || i >= NX - 2 || j < 2 || j >= NY - 2)
        e = UNDEF;
      else if (i > 3 && j < 5)
        e = 0.16;
      else if (i == 3 && j == 5)
        e = 0.04;
      else if (i > 2 && j < 6)
        e = 0.08;
      else
        e = 0;
      CHECK(e == out[i + j * NX], "above i=%d j=%d got %g want %g", i, j, out[i + j * NX], e);
    }

  four();
  c = {5, 3, 1};
  CHECK(fc::neighbourFunctions(NX, NY, in, c, 6, out, f, UNDEF), "prob below");
  CHECK(fc::neighbourProbFunctions(NX, NY, in, c, 6, out2, f2, UNDEF), "prob below (SAT)");
  CHECK(f == SOME_DEFINED && f2 == SOME_DEFINED, "flags %d %d", (int)f, (int)f2);
  for (int i = 0; i < NX; i++)
    for (int j = 0; j < NY; j++) {
      CHECK(float_eq_4ulp(out[i + j * NX], out2[i + j * NX]), "below i=%d j=%d %g vs %g", i, j, out[i + j * NX], out2[i + j * NX]);
      const float e = (i < 3 || i >= NX - 3 || j < 3 || j >= NY - 3) ? UNDEF : (j < 6 ? float(45. / 49.) : float(47. / 49.));
      CHECK(float_eq_4ulp(e, out[i + j * NX]), "below i=%d j=%d got %g want %g", i, j, out[i + j * NX], e);
    }
}

// ---- ReplaceDefined / ReplaceUndefined ----------------------------------------------------------------------
static void test_replace()
{
  const int n = 2;
  const float in1[n] = {0, 1};
  float out1[n] = {-1, -1};
  ValuesDefined d = SOME_DEFINED;
  fc::replaceDefined(n, 1, in1, 5, out1, d, 0);
  CHECK(out1[0] == 0 && out1[1] == 5 && d == ALL_DEFINED, "replaceDefined SOME: %g %g %d", out1[0], out1[1], (int)d);
  d = ALL_DEFINED;
  fc::replaceDefined(n, 1, in1, 7, out1, d, -1);
  CHECK(out1[0] == 7 && out1[1] == 7 && d == ALL_DEFINED, "replaceDefined ALL: %g %g %d", out1[0], out1[1], (int)d);
  d = NONE_DEFINED;
  fc::replaceDefined(n, 1, in1, 7, out1, d, -1);
  CHECK(out1[0] == -1 && out1[1] == -1 && d == NONE_DEFINED, "replaceDefined NONE: %g %g %d", out1[0], out1[1], (int)d);
  d = SOME_DEFINED;
  fc::replaceDefined(n, 1, in1, 1, out1, d, 1); // value == undef
  CHECK(out1[0] == 1 && out1[1] == 1 && d == NONE_DEFINED, "replaceDefined value==undef: %g %g %d", out1[0], out1[1], (int)d);

  out1[0] = out1[1] = -1;
  d = SOME_DEFINED;
  fc::replaceUndefined(n, 1, in1, 5, out1, d, 0);
  CHECK(out1[0] == 5 && out1[1] == 1 && d == ALL_DEFINED, "replaceUndefined SOME: %g %g %d", out1[0], out1[1], (int)d);
  d = ALL_DEFINED;
  fc::replaceUndefined(n, 1, in1, 7, out1, d, -1);
  CHECK(out1[0] == 0 && out1[1] == 1 && d == ALL_DEFINED, "replaceUndefined ALL: %g %g %d", out1[0], out1[1], (int)d);
  d = NONE_DEFINED;
  fc::replaceUndefined(n, 1, in1, 7, out1, d, -1);
  CHECK(out1[0] == 7 && out1[1] == 7 && d == ALL_DEFINED, "replaceUndefined NONE: %g %g %d", out1[0], out1[1], (int)d);
  d = SOME_DEFINED;
  fc::replaceUndefined(n, 1, in1, 1, out1, d, 1); // value == undef
  CHECK(out1[0] == 0 && out1[1] == 1 && d == SOME_DEFINED, "replaceUndefined value==undef: %g %g %d", out1[0], out1[1], (int)d);
}

// ---- MetConstantsTest: ICAO standard atmosphere (doc 7488) --------------------------------------------------
static void test_icao()
{
  using namespace miutil::constants;
  const double doc7488[][2] = {{8.7, 31985},  {10.0, 31055}, {11.1, 30360}, {19.4, 26680}, {97.3, 16353}, {139.5, 14069},
                               {244.1, 10517}, {354.2, 8035}, {459.7, 6189}, {590.8, 4324}, {739.7, 2576}, {840.7, 1547},
                               {936.8, 657},   {1010.0, 27},  {1020.0, -56}, {1050.0, -302}, {1130.0, -929}};
  for (const auto& row : doc7488) {
    CHECK(std::fabs(row[1] - ICAO_geo_altitude_from_pressure(row[0])) <= 1.55, "altitude of %g hPa", row[0]);
    CHECK(std::fabs(row[0] - ICAO_pressure_from_geo_altitude(row[1])) <= 0.01 * row[0], "pressure at %g m", row[1]);
  }
  const int examples[][2] = {{600, 140}, {500, 185}, {400, 235}, {300, 300}, {250, 340}, {200, 385}, {150, 445}};
  for (const auto& row : examples)
    CHECK(row[1] == FL_from_geo_altitude(ICAO_geo_altitude_from_pressure(row[0])), "FL of %d hPa", row[0]);
  for (int i = 0; i < nLevelTable; ++i)
    CHECK((int)fLevelTable[i] == FL_from_geo_altitude(ICAO_geo_altitude_from_pressure(pLevelTable[i])), "FL of level %d", i);
}

// ---- Probability / Probability12 ------------------------------------------------------------------------
static void test_probability()
{
  const float UNDEF = 123456;
  const int M = 10;
  float v[M];
  std::vector<float*> members;
  for (int i = 0; i < M; ++i) {
    v[i] = UNDEF;
    members.push_back(&v[i]);
  }
  v[2] = 940;
  v[4] = 3500;
  std::vector<ValuesDefined> flags(M, SOME_DEFINED);
  flags[0] = NONE_DEFINED;
  flags[8] = NONE_DEFINED; // 8 members count
  std::vector<float> limits(2, 3000);

  auto run = [&](int compute, float expect) {
    float out = UNDEF;
    ValuesDefined f = NONE_DEFINED;
    CHECK(fc::probability(compute, 1, 1, members, flags, limits, &out, f, UNDEF), "probability c=%d", compute);
    CHECK(near(expect, out, 1e-6f), "probability c=%d got %g want %g", compute, out, expect);
    CHECK(f == ALL_DEFINED, "probability c=%d flag %d", compute, (int)f);
  };
  run(2, 100.0f * 1 / 8); // below 3000: 940
  run(1, 100.0f * 1 / 8); // above 3000: 3500
  limits[0] = 4000;
  run(2, 100.0f * 2 / 8);
  limits[0] = 500;
  limits[1] = 4000;
  run(3, 100.0f * 2 / 8); // between

  // Probability12: 8 of 10 members hold 12, two are undefined
  for (int i = 0; i < M; ++i)
    v[i] = 12;
  v[3] = v[5] = UNDEF;
  flags.assign(M, SOME_DEFINED);
  limits.assign(1, 3000);
  run(2, 80);
  run(1, 0);
}

// ---- a stencil through the C++ API: relvort of a solid-body rotation is constant.  The operator evaluates
// 0.5 * mapr * (f[i+1] - f[i-1]) (reference FieldCalculations.cc:1862), so mapr = 1/h gives d/dx.
static void test_relvort_solid_body()
{
  const int nx = 37, ny = 29;
  const float UNDEF = 1e35f, omega = 1e-4f, h = 2500.f;
  std::vector<float> u(nx * ny), v(nx * ny), xm(nx * ny, 1.f / h), ym(nx * ny, 1.f / h), out(nx * ny, -1.f);
  for (int y = 0; y < ny; ++y)
    for (int x = 0; x < nx; ++x) {
      u[y * nx + x] = -omega * (y * h);
      v[y * nx + x] = omega * (x * h);
    }
  ValuesDefined f = SOME_DEFINED;
  CHECK(fc::relvort(nx, ny, u.data(), v.data(), xm.data(), ym.data(), out.data(), f, UNDEF), "relvort");
  CHECK(f == ALL_DEFINED, "relvort flag %d", (int)f);
  int bad = 0;
  for (int i = 0; i < nx * ny; ++i)
    if (!near(2 * omega, out[i], 2e-9f))
      ++bad;
  CHECK(bad == 0, "%d points differ from 2*omega", bad);
  CHECK(!fc::relvort(2, 5, u.data(), v.data(), xm.data(), ym.data(), out.data(), f, UNDEF), "relvort must reject nx < 3");
}

int main()
{
  test_xlevelhum();
  test_aleveltemp_large();
  test_field_oper_field();
  test_abshum();
  test_probability();
  test_neighbour();
  test_replace();
  test_icao();
  test_relvort_solid_body();
  std::printf("%d checks, %d failed\n", g_checks, g_failed);
  return g_failed ? 1 : 0;
}
