// ew_driver.cuh -- host driver of the elementwise engine: one description of a batched call (EwJob),
// two ways to run it.
//
//   * device-resident fields (or a small call): one launch on the thread's main stream;
//   * large batches in HOST memory: the batch is cut into chunks of a few fields and software-
//     pipelined over PIPE_SLOTS streams -- copy-in of chunk c+1, kernel of chunk c and copy-out of
//     chunk c-1 overlap, so the call runs at PCIe speed in both directions at once instead of
//     H2D + kernel + D2H back to back.  Results are identical (the kernels see the same fields).
#pragma once

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "elementwise.cuh"

#include <functional>

namespace fcb200 {

template <class Op>
struct EwJob
{
  static constexpr int NC = Op::NCOUNT > 0 ? Op::NCOUNT : 1;
  int nx = 0, ny = 0, nfields = 0;
  const float* in[Op::NIN] = {};
  bool per_field[Op::NIN] = {}; // false: one [ny][nx] field shared by the whole batch
  float* out[Op::NOUT] = {};
  const int* flags_in = nullptr; // [nfields] ValuesDefined given to the reference call(s)
  int* flags_out[NC] = {};       // per counter: [nfields] flags to derive from it, or nullptr (flag untouched)
  float undef = 0.f;
  std::function<void(int, FieldMeta&)> fill_meta; // per-field scalars (optional)
};

namespace detail {

template <class Op>
int run_ew_range(Call& call, const Op& op, const EwJob<Op>& job, int f0, int f1)
{
  const long long n = (long long)job.nx * job.ny;
  const int nf = f1 - f0;
  const float* in[Op::NIN];
  long long stride[Op::NIN];
  for (int k = 0; k < Op::NIN; ++k) {
    stride[k] = job.per_field[k] ? n : 0;
    in[k] = job.per_field[k] ? call.in(job.in[k] + (size_t)f0 * n, (size_t)(n * nf)) : call.in(job.in[k], (size_t)n);
  }
  float* out[Op::NOUT];
  for (int k = 0; k < Op::NOUT; ++k)
    out[k] = call.out(job.out[k] + (size_t)f0 * n, (size_t)(n * nf));
  FieldMeta* meta = call.meta_host(nf);
  if (!call.ok())
    return -1;
  for (int k = 0; k < nf; ++k) {
    meta[k].all = (job.flags_in[f0 + k] == ALL_DEFINED) ? 1 : 0;
    meta[k].a = meta[k].b = meta[k].c = 0.f;
    if (job.fill_meta)
      job.fill_meta(f0 + k, meta[k]);
  }
  const FieldMeta* dmeta = call.upload_meta();
  unsigned long long* counters = Op::NCOUNT ? call.counters(nf * Op::NCOUNT) : nullptr;
  if (!call.ok())
    return -1;
  if (!launch_elementwise(call, op, in, stride, out, n, nf, job.nx, job.undef, dmeta, counters))
    return -1;
  bool any = false;
  int* flags_out[EwJob<Op>::NC];
  for (int c = 0; c < EwJob<Op>::NC; ++c) {
    flags_out[c] = (Op::NCOUNT > 0) ? job.flags_out[c] : nullptr;
    any = any || flags_out[c];
  }
  if (!any)
    return call.finish(Finalizer());
  const unsigned long long un = (unsigned long long)n;
  int* fo[EwJob<Op>::NC];
  for (int c = 0; c < EwJob<Op>::NC; ++c)
    fo[c] = flags_out[c];
  // capture by value: plain ints and pointers into caller memory
  struct Cap
  {
    int* fo[EwJob<Op>::NC];
  } cap;
  for (int c = 0; c < EwJob<Op>::NC; ++c)
    cap.fo[c] = fo[c];
  return call.finish([=](const unsigned long long* cnt) {
    for (int k = 0; k < nf; ++k)
      for (int c = 0; c < EwJob<Op>::NC; ++c)
        if (cap.fo[c])
          cap.fo[c][f0 + k] = check_defined(cnt[(size_t)k * EwJob<Op>::NC + c], un);
  });
}

} // namespace detail

template <class Op>
int run_ew_job(const Op& op, const EwJob<Op>& job)
{
  if (job.nx <= 0 || job.ny <= 0 || job.nfields <= 0 || (long long)job.nx * job.ny >= 0x7fffffffLL) {
    set_error("fcb200: invalid grid or batch size (nx=%d ny=%d nfields=%d)", job.nx, job.ny, job.nfields);
    return -1;
  }
  const long long n = (long long)job.nx * job.ny;

  // pipeline only if a per-field array lives in host memory and the batch is big enough to matter
  bool host_fields = false;
  {
    Call probe;
    if (!probe.ok())
      return -1;
    for (int k = 0; k < Op::NIN; ++k)
      if (job.per_field[k] && !probe.is_device(job.in[k]))
        host_fields = true;
    for (int k = 0; k < Op::NOUT; ++k)
      if (!probe.is_device(job.out[k]))
        host_fields = true;
    if (!probe.ok())
      return -1;
    // in-place host calls (output aliases an input) stay on the simple path
    for (int k = 0; k < Op::NIN; ++k)
      for (int o = 0; o < Op::NOUT; ++o)
        if (static_cast<const void*>(job.in[k]) == static_cast<const void*>(job.out[o]))
          host_fields = false;
  }
  const size_t bytes_per_field = sizeof(float) * (size_t)n * (Op::NIN + Op::NOUT);
  // Chunk sizes (measured on the 65-level chain, pinned buffers, B200 + PCIe 5: profiles/r02_e2e_chunking.txt): a chunk costs
  // ~40 us on top of its bytes, so big chunks -- but the first copy-in and the last copy-out overlap nothing, so SMALL chunks at
  // both ends: 1, 2, 4, ... fields up to ~192 MB per chunk and down again.  (48 MB chunks throughout: 26.8 ms per step; 200 MB
  // throughout: 24.5 ms; the copies alone: 20.3 ms.)
  static const char* chunk_env = getenv("FCB200_CHUNK_MB"); // (development switch)
  const size_t chunk_target = size_t(chunk_env ? atoi(chunk_env) : 192) << 20;
  int max_fields = (int)(chunk_target / bytes_per_field);
  if (max_fields < 1)
    max_fields = 1;
  if (!host_fields || job.nfields < 2 || bytes_per_field * (size_t)job.nfields < (size_t(96) << 20)) {
    Call call;
    return detail::run_ew_range(call, op, job, 0, job.nfields);
  }
  std::vector<int> head, tail;
  for (int left = job.nfields, h = 1; left > 0; h = (h < max_fields) ? 2 * h : h) {
    const int a = std::min(std::min(h, max_fields), left);
    head.push_back(a);
    left -= a;
    if (left <= 0)
      break;
    const int b = std::min(std::min(h, max_fields), left);
    tail.push_back(b);
    left -= b;
  }
  head.insert(head.end(), tail.rbegin(), tail.rend());

  if (!pipeline_fork())
    return -1;
  int rc = 1, f0 = 0;
  for (size_t c = 0; c < head.size() && rc == 1; ++c) {
    const int f1 = f0 + head[c];
    Call call(1 + (int)(c % PIPE_SLOTS));
    rc = detail::run_ew_range(call, op, job, f0, f1);
    f0 = f1;
  }
  const int jr = pipeline_join();
  return (rc == 1) ? jr : rc;
}

} // namespace fcb200
