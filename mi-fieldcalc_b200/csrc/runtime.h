// runtime.h -- host-side plumbing shared by every fcb200_* entry point.
//
// The reference (mi-fieldcalc) is a stateless host library: the caller owns every buffer and each
// function returns when its output and its ValuesDefined flag are final.  This layer
// keeps that contract on a GPU:
//   * every field pointer may be DEVICE memory (used in place -- the roofline path) or HOST memory
//     (pinned or pageable; staged through a per-thread device arena -- the drop-in path);
//   * per-field metadata (the allDefined bit and up to three scalars such as p / alevel / blevel)
//     goes to the device in one small copy; per-field undefined counters come back in one small copy;
//   * one stream per host thread (or the stream the caller installs with fcb200_set_stream), one
//     synchronisation per call, no global mutable state without a lock -> callable concurrently;
//   * between fcb200_begin_deferred() and fcb200_end_deferred() calls only enqueue work: outputs,
//     counters and flags become final at fcb200_end_deferred() (one synchronisation for many calls).
// There is no CPU compute path anywhere: if CUDA is unavailable every entry point returns < 0.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <functional>
#include <vector>

namespace fcb200 {

enum { ALL_DEFINED = 0, NONE_DEFINED = 1, SOME_DEFINED = 2 }; // FieldDefined.h:41 of the reference

// per-field record read by the kernels (16 bytes, one per field of a batch)
struct FieldMeta
{
  int all;       // 1 = the field's input flag is ALL_DEFINED -> skip every undefined test
  float a, b, c; // operator specific scalars (p, pidcp, alevel, blevel ...)
};

// flag = f(n_undefined, N), FieldDefined.cc:62-70 of the reference
inline int check_defined(unsigned long long n_undefined, unsigned long long n)
{
  if (n_undefined == 0)
    return ALL_DEFINED;
  if (n_undefined == n)
    return NONE_DEFINED;
  return SOME_DEFINED;
}

void set_error(const char* fmt, ...);
// slab.cu installs the cross-rank reduction (an in-place ncclAllReduce(MIN) of `n` ints on `stream`; false = error text set)
typedef bool (*SlabReduceFn)(int* words, size_t n, cudaStream_t stream);
void set_slab_reduce(SlabReduceFn fn);
// enqueue the cross-rank combination of the flags of the calls queued so far in deferred mode (see Call::finish_counted)
int reduce_queued_flags();
bool cuda_ok(cudaError_t e, const char* what);
void count_launch(unsigned n = 1);
int sm_count();

// A finaliser receives the host copy of the call's counters once the stream has drained; it writes
// the ValuesDefined flags (and nothing else) into caller memory.
typedef std::function<void(const unsigned long long* counters)> Finalizer;

// Chunked host-pointer calls are software-pipelined over PIPE_SLOTS extra streams: while chunk c
// computes, chunk c+1 is copied in and chunk c-1 is copied out (PCIe is full duplex and the copy
// engines run beside the SMs).  Each slot owns a stream and a device arena.
constexpr int PIPE_SLOTS = 4;

// fork the calling thread's pipeline streams off its main stream / join them back and drain.
bool pipeline_fork();
int pipeline_join();

// One Call object lives for the duration of one fcb200_* entry point (slot 0, the thread's main
// stream) or of one chunk of a pipelined entry point (slot 1..PIPE_SLOTS).
class Call
{
public:
  explicit Call(int slot = 0);
  ~Call();

  bool ok() const { return ok_; }
  cudaStream_t stream() const { return stream_; }

  // resolve a read-only field of `count` floats: device pointers are returned unchanged, host pointers
  // are copied to the arena (asynchronously on the call's stream)
  const float* in(const float* p, size_t count);
  // resolve a written field: device pointers unchanged; host pointers get an arena buffer that finish()
  // copies back.  If `p` equals a host pointer already resolved with in()/inout() the same arena buffer
  // is returned (the reference allows output == input for elementwise operators).
  float* out(float* p, size_t count);
  // field that is read and written in place
  float* inout(float* p, size_t count);
  // scratch bytes in the device arena (never copied)
  void* scratch(size_t bytes);
  // true if `p` is device memory (or the call already failed)
  bool is_device(const void* p);

  // per-field metadata: fill the returned host array, then call upload_meta() to get the device copy
  FieldMeta* meta_host(int nfields);
  const FieldMeta* upload_meta();
  // `count` zero-initialised 64-bit device counters
  unsigned long long* counters(int count);
  // small read-only table (pointers, limits ...) copied to the device
  const void* upload_small(const void* host, size_t bytes);

  // copy host outputs back, fetch the counters, synchronise and run `fin` -- or, in deferred mode,
  // queue all of that for fcb200_end_deferred().  Returns 1 on success, -1 on a runtime error.
  int finish(const Finalizer& fin);
  // finish() for the common rule "fDefined[k] = check_defined(counters[counter_offset + k], denom)".  Calls that finish this way
  // can have their flags combined across the ranks of a row-slab run ON THE STREAM (fcb200_slab_reduce_flags, slab.cu): the
  // runtime knows the rule, so {count == 0, count == denom} can be formed and all-reduced on the device.
  int finish_counted(int* fDefined, int nfields, unsigned long long denom, int counter_offset = 0);
  // bytes of host memory this call staged so far (0 = everything was device resident)
  size_t staged_bytes() const { return staged_; }

private:
  struct Pending
  {
    const void* host;
    void* dev;
    size_t bytes;
    bool copy_back;
  };
  void* arena_alloc(size_t bytes);
  void* pinned_alloc(size_t bytes);
  bool classify(const void* p, bool* is_dev);

  struct ThreadState* ts_ = nullptr;
  int slot_ = 0;
  size_t staged_ = 0;
  cudaStream_t stream_ = nullptr;
  bool ok_ = true;
  std::vector<Pending> pending_;
  FieldMeta* meta_host_ = nullptr;
  int meta_n_ = 0;
  unsigned long long* counters_dev_ = nullptr;
  int counters_n_ = 0;
  bool counters_pooled_ = false;
  bool counters_graph_ = false; // from the counter block of the graph being captured
  bool finished_ = false;
};

} // namespace fcb200
