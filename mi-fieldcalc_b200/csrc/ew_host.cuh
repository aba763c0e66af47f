// ew_host.cuh -- host-side helpers shared by the translation units that define elementwise operators
// (ops_elementwise.cu, ops_diagnostics.cu): the batch descriptor, the adapter from an entry point's
// argument list to an EwJob, and the host copies of the meteorological constants.
#pragma once

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ew_driver.cuh"

namespace fcb200 {
namespace {

// host copies of MC.h:39-49 (same float values as the device constexprs)
static const float H_CP = 1004.f, H_P0INV = (float)(1. / 1000.f), H_KAPPA = 287.f / 1004.f, H_T0 = (float)273.15;

inline float host_pidcp(float p)
{ // FC.cc:308-311, evaluated once per field on the host with glibc powf -- exactly what the reference
  // does for every plevel* operator (FC.cc:347, 434), so these operators carry no device-powf ulps
  return powf(p * H_P0INV, H_KAPPA);
}

inline bool unit_is(const char* unit, const char* what)
{
  return unit && strcmp(unit, what) == 0;
}

struct Batch
{
  int nx, ny, nfields;
  long long n;
  bool valid() const { return nx > 0 && ny > 0 && nfields > 0 && (long long)nx * ny < 0x7fffffffLL; }
};

inline Batch make_batch(int nx, int ny, int nfields)
{
  Batch b;
  b.nx = nx;
  b.ny = ny;
  b.nfields = nfields;
  b.n = (long long)nx * ny;
  return b;
}

enum FlagRule { FLAG_FROM_COUNT, FLAG_UNCHANGED };

// Runs one elementwise operator over a batch.  `stride[k]` = 1 for per-field arrays, 0 for arrays shared
// by the batch.  `fill_meta(k, meta)` sets the per-field scalars.  The output may alias an input.
// Adapter from the argument lists of the entry points to an EwJob (single output, one counter).
template <class Op, class FillMeta>
int run_elementwise(const Batch& b, const Op& op, const float* const* host_in, const int* per_field, float* host_out, int* fDefined, float undef,
                    FlagRule rule, FillMeta fill_meta)
{
  EwJob<Op> job;
  job.nx = b.nx;
  job.ny = b.ny;
  job.nfields = b.nfields;
  for (int k = 0; k < Op::NIN; ++k) {
    job.in[k] = host_in[k];
    job.per_field[k] = per_field[k] != 0;
  }
  job.out[0] = host_out;
  job.flags_in = fDefined;
  job.flags_out[0] = (rule == FLAG_FROM_COUNT && Op::NCOUNT) ? fDefined : nullptr;
  job.undef = undef;
  job.fill_meta = fill_meta;
  return run_ew_job(op, job);
}

struct NoMeta
{
  void operator()(int, FieldMeta&) const {}
};

} // namespace
} // namespace fcb200
