/* fcb200.h -- C-ABI of the B200-native FieldCalculations hot path (libfcb200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  A maintainer
 * of the reference binds exactly these symbols (INTEGRATION.md shows the binding): the C++ shim
 * mi-fieldcalc_b200/csrc/shim.cc re-exports them under the reference's own mangled names
 * miutil::fieldcalc::<op> (src/mi_fieldcalc/FieldCalculations.h:113-303 of the reference).
 *
 *   fcb200_<op>(...)           one field per call; signature = the reference function of the same
 *                              name (listed with file:line in fcb200_api.inc)
 *   fcb200_<op>_batched(...)   many fields per launch (fcb200_batched.inc)
 *
 * Memory: every field pointer may be DEVICE memory (used in place) or HOST memory (staged through
 * a per-thread device arena; pinned host memory copies at full PCIe speed).  The caller owns all
 * buffers; the library keeps no pointer after a call returns (after fcb200_end_deferred() in
 * deferred mode).  Scalars, `fDefined` flag arrays, per-field scalar arrays, pointer tables and
 * limits are always HOST memory.
 *
 * Return value: 1 = the reference would return true, 0 = the reference would return false
 * (arguments rejected; outputs untouched unless the reference touches them), < 0 = runtime
 * failure (no CUDA device, CUDA error); fcb200_last_error() describes it.  There is no CPU path.
 *
 * Threads: entry points are re-entrant; each host thread has its own stream and staging arena.
 */
#ifndef FCB200_H
#define FCB200_H

#ifdef __cplusplus
extern "C" {
#endif

enum { FCB200_ALL_DEFINED = 0, FCB200_NONE_DEFINED = 1, FCB200_SOME_DEFINED = 2 };

/* ---- runtime ------------------------------------------------------------------------------ */
const char* fcb200_version(void);
/* text of the calling thread's last runtime error ("" if none) */
const char* fcb200_last_error(void);
/* number of visible CUDA devices, < 0 if the CUDA runtime cannot be initialised */
int fcb200_device_count(void);
/* make `device` current for the calling thread (cudaSetDevice) */
int fcb200_set_device(int device);
/* run the calling thread's work on `cuda_stream` (a cudaStream_t) if use_it != 0, else on the
 * library's own per-thread stream (default) */
int fcb200_set_stream(void* cuda_stream, int use_it);
/* deferred mode: calls between begin and end only enqueue work; outputs in host memory, counters
 * and fDefined flags become final when fcb200_end_deferred() returns.  Input flags are read at
 * call time, so a deferred call must not depend on the flag of an earlier deferred call. */
int fcb200_begin_deferred(void);
int fcb200_end_deferred(void);
/* 1 if the calling thread is between fcb200_begin_deferred() and fcb200_end_deferred() */
int fcb200_in_deferred(void);
/* wait for the calling thread's stream */
int fcb200_synchronize(void);
/* graphs: a chain of calls captured ONCE into a CUDA graph and replayed with ONE launch (the reference has no counterpart:
 * its callers chain single-field calls, e.g. pleveltemp -> relvort -> divergence, FieldCalculations.cc:328/1843/1910, and pay
 * one kernel launch + one synchronisation per call on a GPU).  Between fcb200_graph_begin() and fcb200_graph_end() every
 * fcb200_* call on DEVICE-resident fields is recorded instead of run (host-memory fields and operators that need a
 * host-side decision, cvtemp compute 3, make the capture fail; fcb200_graph_end() then returns -1 and
 * fcb200_last_error() names the call).  fcb200_graph_launch() runs the whole chain and returns when every output and
 * every fDefined flag is final (in deferred mode: at fcb200_end_deferred()).  The field pointers and the fDefined pointers of
 * the captured calls belong to the caller and must stay valid while the graph is in use; input flags and scalar arguments
 * are frozen at capture time; new DATA in the same buffers is what a launch sees.  A graph belongs to the device it was
 * captured on and is used by one host thread at a time. */
int fcb200_graph_begin(void);
int fcb200_graph_end(void** graph);
int fcb200_graph_launch(void* graph);
/* kernel nodes in the graph (what one launch adds to fcb200_launch_count()), < 0 if `graph` is NULL */
int fcb200_graph_kernels(void* graph);
int fcb200_graph_destroy(void* graph);
/* kernels launched by this library since it was loaded (all threads) */
unsigned long long fcb200_launch_count(void);

/* ---- row slabs: ONE large grid over several GPUs (one process or host thread per GPU) -------------
 * The reference has no counterpart (it is a single-host library): BASELINE.json configs[2] asks for
 * "row-slab sharded across 8 GPUs with NVLink halo exchange".  A rank owns rows [r0, r1) of every field and holds the
 * extended slab [lo, hi) = its rows + `halo` rows of each neighbour (1 for the five-point stencils, 2 for
 * thermalFrontParameter and shapiro2_filter); it calls the ordinary fcb200_<op>_batched(nx, hi - lo, ...) on the extended
 * slab and keeps its own rows.  NCCL (bound at run time, libnccl.so.2) moves the halo rows. */
/* 128-byte NCCL unique id: call on ONE rank, hand the bytes to every rank (any transport: MPI, torch.distributed, a file) */
int fcb200_slab_unique_id(char* id128);
/* join the communicator with the calling thread's current device; nranks == 1 is valid (every slab call is then a no-op) */
int fcb200_slab_init(int rank, int nranks, const char* id128);
int fcb200_slab_finalize(void);
int fcb200_slab_rank(void);
int fcb200_slab_nranks(void);
/* rows [r0, r1) owned by `rank` and rows [lo, hi) of its extended slab; 0 if ny is too small for nranks slabs of this halo */
int fcb200_slab_partition(int ny, int halo, int rank, int nranks, int* r0, int* r1, int* lo, int* hi);
/* halo exchange of `nfields` extended slabs (DEVICE memory, field k at ext + k*ext_rows*nx, owned rows already in place):
 * the first / last `halo` owned rows go to rank-1 / rank+1, theirs arrive in this slab's halo rows.  Enqueued on the calling
 * thread's stream (ordered against the operators before and after it); every rank of the communicator must call it. */
int fcb200_slab_exchange(float* ext, int nx, int ext_rows, int nfields, int halo);
/* global flags from the per-rank flags of a sharded operator (HOST array, in/out): ALL iff every rank says ALL, NONE iff every
 * rank says NONE, else SOME.  Synchronises the calling thread first (local flags must be final); collective. */
int fcb200_slab_combine_flags(int* fDefined, int nfields);
/* the same combination ON THE STREAM, for the calls queued so far in deferred mode: between fcb200_begin_deferred() and
 * fcb200_end_deferred(), after the sharded operators, enqueue one tiny kernel + one ncclAllReduce over {no undefined point,
 * all points undefined} per field; fcb200_end_deferred() then writes the GLOBAL flags into the fDefined arrays of those calls --
 * one synchronisation per step instead of two and no host round trip before the collective.  Covers every operator whose flag
 * is checkDefined(count, n) of its own counter (the stencil family); other calls keep their local flags (combine those with
 * fcb200_slab_combine_flags).  Collective: every rank queues the same calls.  No-op for one rank. */
int fcb200_slab_reduce_flags(void);
/* payload bytes this process has sent through fcb200_slab_exchange */
unsigned long long fcb200_slab_bytes_sent(void);

/* ---- operators ----------------------------------------------------------------------------- */
#define FC_FN(name, args) int fcb200_##name args;
#include "fcb200_api.inc"
#undef FC_FN

#define FCB_FN(name, args) int fcb200_##name args;
#include "fcb200_batched.inc"
#undef FCB_FN

#ifdef __cplusplus
}
#endif

#endif /* FCB200_H */
