/*
 * gsfield.h -- C ABI of the B200-native randomization-method field summation.
 *
 * Drop-in boundary for GSTools-Core's `field::summator*` hot path.  Each entry point is what the
 * reference's own Rust function body (and through it the pyo3 module) binds after dispatching to
 * the CUDA path; see INTEGRATION.md for the Rust `extern "C"` block and the ctypes binding.
 *
 *   gsf_summate          replaces field::summator          /root/reference/src/field.rs:37-65
 *   gsf_summate_incompr  replaces field::summator_incompr  /root/reference/src/field.rs:97-182
 *   gsf_summate_fourier  replaces field::summator_fourier  /root/reference/src/field.rs:219-249
 *
 * Conventions
 *   - Every array is f64 and addressed as base[i*stride0 + j*stride1] with ELEMENT strides, so any
 *     ndarray / numpy view the reference accepts (src/lib.rs:43-46) can be passed without a copy.
 *   - The caller owns every buffer and pre-allocates the output.  The library owns device buffers,
 *     streams and pinned staging inside a lazily created process-wide context (gsf_shutdown frees it).
 *   - `pos` and `out` may live in pageable host memory, pinned host memory or device memory; the
 *     library detects which (cudaPointerGetAttributes).  Host data is streamed through the GPU in
 *     chunks (H2D / kernel / D2H overlapped); device-resident data is processed in place.
 *   - Device-resident arguments of the BLOCKING entry points (gsf_summate*, gsf_summate_ex,
 *     gsf_krige): the library works on its own non-blocking streams, so it first waits for all
 *     work queued on the owning device (cudaDeviceSynchronize) -- arrays still being produced on
 *     one of the caller's streams are safe to pass.  gsf_summate_on_stream instead orders its work
 *     on the stream the caller names and never synchronises: the arrays must be produced on that
 *     stream (or be complete) -- the usual CUDA stream contract.
 *   - `num_threads` is accepted for signature compatibility with the reference (src/field.rs:42);
 *     it never affects results and only caps the host threads used for staging copies of
 *     pageable memory.  <= 0 means "default".
 *   - Return value: GSF_OK or an error code; gsf_last_error() gives a thread-local message.  The
 *     library never aborts and never throws across the ABI (the reference panics => abort,
 *     Cargo.toml:22; a Rust shim turns a non-zero status back into panic!).
 *   - There is NO CPU fallback: without a usable CUDA device every compute entry point returns
 *     GSF_ERR_NO_DEVICE.
 *   - Calls are serialised per process by an internal mutex; callable from any thread.
 *
 * Environment (read once per process; tuning and A/B switches, none changes a result beyond the
 * documented 1e-9 sigma contract)
 *   GSF_DEVICES="0,1,.."      default device list of the host-memory entry points (gsf_set_devices overrides)
 *   GSF_GRID_DETECT=0         no automatic structured-grid detection (gsf_set_grid_detection)
 *   GSF_MODE_GROUPS=0         structured-grid path: never sum runs of tensor-structured modes before the GEMM
 *   GSF_POLY_DEGREE=6         always the high-degree cosine polynomial (gsf_set_poly_degree)
 *   GSF_STAGING_THREADS=n     host threads staging pageable memory per device (default: from the demand)
 *   GSF_NUMA_BIND=0|1         one process per GPU: bind the rank's threads to the GPU's NUMA node
 *                             (default: only when a launcher exports LOCAL_WORLD_SIZE > 1)
 *   GSF_ZERO_COPY=0           pinned pos/out through the copy pipeline instead of mapped access
 *   GSF_SMALL_KB=n            positions up to n KB take the one-launch paths (default 800)
 *   GSF_SMALL_FUSED=0, GSF_SMALL_PROMOTE=0   small calls: no gsf_small_kernel / no repeat-mode promotion
 *   GSF_CHUNK_FIRST / _CAP / _CAP_DIV / _GROWTH / _DOWN, GSF_TAIL_WAVES, GSF_GRID_KSPLIT, GSF_GRID_CHUNK_MB
 *                             pipeline chunk schedule, short-tile tail, grid-path split / chunk size
 *   GSF_PINNED_CACHE_MB, GSF_PINNED_LIVE_MB   caching pinned allocator behind gsf_host_alloc
 *   GSF_POOL_SPIN_US          how long idle pool workers spin before sleeping (0 under a multi-rank launcher)
 *   GSF_TRACE=1               per-phase host timeline of every call on stderr
 */
#ifndef GSFIELD_H
#define GSFIELD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSF_ABI_VERSION 2

enum gsf_status {
    GSF_OK = 0,
    GSF_ERR_DIM = 1,        /* dim < 1, or incompr with dim not in {2,3} (field.rs:180)                   */
    GSF_ERR_SHAPE = 2,      /* negative sizes / inconsistent arguments (field.rs:44-46 asserts)          */
    GSF_ERR_EMPTY_MODES = 3,/* summator_incompr with 0 modes (reduce_with(..).unwrap(), field.rs:163)    */
    GSF_ERR_NO_DEVICE = 4,  /* no CUDA device / driver; there is no CPU path                             */
    GSF_ERR_CUDA = 5,       /* a CUDA runtime call failed (message has the details)                      */
    GSF_ERR_ARG = 6,        /* NULL pointer, bad device id, mixed device/host placement not supported    */
    GSF_ERR_ALLOC = 7       /* host or device allocation failed                                          */
};

/* Largest dim served by the tuned kernel templates.  Scalar / Fourier calls with a larger dim are
 * accepted too (the reference takes any dim) and run a plain one-point-per-thread kernel. */
#define GSF_MAX_DIM 8

/* Per-call statistics of the most recent compute call on this thread's context. Times in ms. */
typedef struct gsf_stats {
    double total_ms;        /* host wall clock of the whole call                                        */
    double kernel_ms;       /* summation kernel(s): CUDA events on the launching stream(s), summed over
                               chunks, max over devices                                                  */
    double prep_ms;         /* mode pre-processing kernel                                                */
    int64_t point_modes;    /* n_points * n_modes                                                        */
    int64_t h2d_bytes;      /* bytes copied host->device                                                 */
    int64_t d2h_bytes;      /* bytes copied device->host                                                 */
    int32_t kernel_launches;/* number of kernels launched by this call (prep + summation)                */
    int32_t n_devices;      /* devices that took part                                                    */
    int32_t n_chunks;       /* pipeline chunks (over all devices)                                        */
    int32_t points_per_thread; /* P of the kernel variant used                                          */
    int32_t lanes_per_point;   /* L of the kernel variant used                                          */
    int32_t pos_memory;     /* 0 pageable host, 1 pinned host, 2 device                                  */
    int32_t out_memory;
    int32_t grid_path;      /* 0 general kernel, 1 structured grid auto-detected, 2 structured grid requested */
    int32_t poly_degree;    /* degree of the cosine polynomial the summation kernel used (0: grid path)  */
    int32_t fp64_slots;     /* FP64-pipe instructions per point*mode of that kernel: dim + 5 + degree + nc */
    int32_t staging_threads;/* host threads that staged pageable memory through the pinned rings         */
    int32_t mode_group;     /* grid path: consecutive modes summed before the GEMM (tensor-structured modes), 1 = none; 0 otherwise */
} gsf_stats;

/* ---- the three reference functions ------------------------------------------------------- */

/* out[j] = sum_i z1[i]*cos(<k_i,x_j>) + z2[i]*sin(<k_i,x_j>);  out: n_points contiguous f64. */
int gsf_summate(int dim, int64_t n_modes, int64_t n_points,
                const double *cov_samples, int64_t cov_s0, int64_t cov_s1, /* (dim, n_modes)  */
                const double *z1, int64_t z1_s,                            /* (n_modes)       */
                const double *z2, int64_t z2_s,                            /* (n_modes)       */
                const double *pos, int64_t pos_s0, int64_t pos_s1,         /* (dim, n_points) */
                double *out, int num_threads);

/* out[a*out_s0 + j*out_s1] = sum_i p_a(k_i) * (z1 cos + z2 sin);  dim in {2,3}.  The reference
 * returns shape (dim, n_points) in Fortran order (field.rs:166-174): out_s0 = 1, out_s1 = dim. */
int gsf_summate_incompr(int dim, int64_t n_modes, int64_t n_points,
                        const double *cov_samples, int64_t cov_s0, int64_t cov_s1,
                        const double *z1, int64_t z1_s,
                        const double *z2, int64_t z2_s,
                        const double *pos, int64_t pos_s0, int64_t pos_s1,
                        double *out, int64_t out_s0, int64_t out_s1, int num_threads);

/* out[j] = sum_i sf[i] * (z1[i]*cos(<k_i,x_j>) + z2[i]*sin(<k_i,x_j>)). */
int gsf_summate_fourier(int dim, int64_t n_modes, int64_t n_points,
                        const double *spectrum_factor, int64_t sf_s,       /* (n_modes)       */
                        const double *modes, int64_t modes_s0, int64_t modes_s1,
                        const double *z1, int64_t z1_s,
                        const double *z2, int64_t z2_s,
                        const double *pos, int64_t pos_s0, int64_t pos_s1,
                        double *out, int num_threads);

/* ---- extended request: fused post-scale / offset and structured grids ---------------------- */

/* One struct for everything beyond the three reference signatures:
 *   out[a, j] = scale * sum_i(...) + offset[a]      (GSTools applies sqrt(var/N) and the mean right
 *                                                    after the call; here it costs nothing)
 *   n_axes > 0: the points are the rectilinear grid axis[0] x axis[1] (x axis[2]) flattened in C
 *   order (GSTools mesh_type="structured"); pos is ignored, n_points is prod(axis_n).  The sum then
 *   factorises per axis and runs as an FP64 tensor-core GEMM (2 FMA per point*mode instead of 15).
 * The plain entry points detect such grids in host-resident `pos` automatically (exact, bitwise
 * check; gsf_set_grid_detection(0) or GSF_GRID_DETECT=0 turns that off).
 * On this path tensor-structured MODES are exploited too: when consecutive runs of g host-resident modes
 * share every wave-vector component but the last (the Fourier method's lattice, modes = meshgrid(kx, ky
 * [, kz]) flattened in C order; detected exactly), the runs are summed before the GEMM and the
 * contraction shrinks by g (gsf_stats::mode_group; GSF_MODE_GROUPS=0 turns it off).
 * Set struct_size = sizeof(gsf_request); zero-initialise the rest you do not use (scale = 1). */
typedef struct gsf_request {
    int32_t struct_size;
    int32_t kind;           /* 0 summate, 1 summate_incompr, 2 summate_fourier */
    int32_t dim;
    int32_t num_threads;
    int64_t n_modes, n_points;
    const double *spectrum_factor; int64_t sf_s;
    const double *modes; int64_t modes_s0, modes_s1;
    const double *z1; int64_t z1_s;
    const double *z2; int64_t z2_s;
    const double *pos; int64_t pos_s0, pos_s1;
    double *out; int64_t out_s0, out_s1;
    double scale;
    double offset[3];
    int32_t n_axes;         /* 0, or == dim (2 or 3) for a structured grid */
    int32_t reserved;
    const double *axis[3]; int64_t axis_n[3]; int64_t axis_s[3];   /* host-resident axis vectors */
} gsf_request;

int gsf_summate_ex(const gsf_request *request);
/* 1 / 0: enable / disable automatic structured-grid detection; -1: follow GSF_GRID_DETECT (default on). */
int gsf_set_grid_detection(int enabled);

/* ---- kriging (reference: src/krige.rs:24-118) ----------------------- */

/* field[p] = sum_i cond[i] * <krig_mat[:, i], krig_vecs[:, p]>, and, if `error` is not NULL,
 * error[p] = sum_i krig_vecs[i, p] * <krig_mat[:, i], krig_vecs[:, p]>  (the kriging variance
 * term).  krig_mat is (n_cond, n_cond), krig_vecs (n_cond, n_points), element strides as usual;
 * arrays may be host or device resident.  Replaces calculator_field_krige (error == NULL,
 * src/krige.rs:93-118) and calculator_field_krige_and_variance (src/krige.rs:24-73). */
int gsf_krige(int64_t n_cond, int64_t n_points,
              const double *krig_mat, int64_t mat_s0, int64_t mat_s1,
              const double *krig_vecs, int64_t vecs_s0, int64_t vecs_s1,
              const double *cond, int64_t cond_s,
              double *field, double *error, int num_threads);

/* ---- empirical variograms (reference: src/variogram.rs; bindings src/lib.rs:119-216) -------- */

/* estimator_type: 'c' = Cressie, anything else = Matheron (src/variogram.rs:22-39).  Host-resident
 * arrays with element strides (mask: byte strides == element strides of a bool array).  Counts are
 * exact; for Euclidean distances every pair lands in the bin(s) the reference puts it in (see
 * gsf_variogram_kernels.cuh); the sums run in a fixed, run-to-run deterministic order that differs
 * from the reference's sequential one (relative difference ~1e-15).
 *
 * variogram_structured / variogram_ma_structured (src/variogram.rs:136-178, :190-240): f is
 * (n0, n1), mask NULL or (n0, n1) with non-zero = masked; variogram gets n0 entries, [0] = 0. */
int gsf_variogram_structured(int64_t n0, int64_t n1, const double *f, int64_t f_s0, int64_t f_s1,
                             const uint8_t *mask, int64_t mask_s0, int64_t mask_s1, char estimator_type,
                             double *variogram, int num_threads);
/* variogram_unstructured (src/variogram.rs:465-545): f (n_fields, n_points), bin_edges (n_bins + 1),
 * pos (dim, n_points), dim 1..3; distance_type 'e' = Euclid, anything else = Haversine (dim 2,
 * degrees).  variogram and counts get n_bins entries.  NaN field differences are skipped. */
int gsf_variogram_unstructured(int dim, int64_t n_fields, int64_t n_points, int64_t n_bins, const double *f,
                               int64_t f_s0, int64_t f_s1, const double *bin_edges, int64_t edges_s,
                               const double *pos, int64_t pos_s0, int64_t pos_s1, char estimator_type,
                               char distance_type, double *variogram, uint64_t *counts, int num_threads);
/* variogram_directional (src/variogram.rs:315-447): direction is (n_dirs, dim), expected normed;
 * variogram and counts are (n_dirs, n_bins) row-major.  angles_tol must be > 0; bandwidth <= 0
 * disables the band test; separate_dirs stops at the first matching direction. */
int gsf_variogram_directional(int dim, int64_t n_fields, int64_t n_points, int64_t n_bins, int64_t n_dirs,
                              const double *f, int64_t f_s0, int64_t f_s1, const double *bin_edges,
                              int64_t edges_s, const double *pos, int64_t pos_s0, int64_t pos_s1,
                              const double *direction, int64_t dir_s0, int64_t dir_s1, double angles_tol,
                              double bandwidth, int separate_dirs, char estimator_type, double *variogram,
                              uint64_t *counts, int num_threads);
/* Host logic of the bit-faithful binning (no device needed): min{x : sqrt(x) >= edge} and
 * max{a in [0,1) : acos(a) >= angles_tol} (-1 if none).  Either output may be NULL. */
int gsf_debug_variogram_thresholds(double edge, double angles_tol, double *sqrt_thr, double *acos_thr);

/* Host-logic introspection for tests (no device needed): the pipeline chunk sizes for n_points
 * host-resident points (returns the count, or -count if max_sizes is too small), and the exact
 * structured-grid detector (returns 1 and the axis lengths, or 0). */
int gsf_debug_chunk_schedule(int64_t n_points, int64_t forced_chunk, int64_t *sizes, int max_sizes);
/* The sysfs cpulist parser behind the rank-to-NUMA-node binding ("0-15,64-79"): writes up to max_cpus
 * cpu ids in increasing order, returns how many the list names. */
int gsf_debug_parse_cpulist(const char *list, int *cpus, int max_cpus);
/* The exact detector of tensor-structured (lattice) modes used by the structured-grid path: returns the
 * common run length g of consecutive host-resident modes sharing all but the last wave-vector component,
 * or 1 when there is no such structure. */
int64_t gsf_debug_mode_group(int dim, int64_t n_modes, const double *modes, int64_t modes_s0, int64_t modes_s1);
int gsf_debug_detect_grid(int dim, int64_t n_points, const double *pos, int64_t pos_s0, int64_t pos_s1,
                          int64_t *axis_n);

/* ---- stream-ordered variant for device-resident data -------------------------------------- */

/* kind: 0 summate, 1 summate_incompr, 2 summate_fourier.  pos and out must be device pointers on
 * the same device; mode arrays may be host or device.  Work is enqueued on `cuda_stream`
 * (a cudaStream_t; NULL = legacy default stream) and the call returns without synchronising.
 * spectrum_factor is ignored unless kind == 2; out strides are ignored unless kind == 1
 * (scalar outputs are contiguous). */
int gsf_summate_on_stream(int kind, int dim, int64_t n_modes, int64_t n_points,
                          const double *spectrum_factor, int64_t sf_s,
                          const double *cov_samples, int64_t cov_s0, int64_t cov_s1,
                          const double *z1, int64_t z1_s,
                          const double *z2, int64_t z2_s,
                          const double *pos, int64_t pos_s0, int64_t pos_s1,
                          double *out, int64_t out_s0, int64_t out_s1,
                          void *cuda_stream);

/* ---- context / diagnostics ---------------------------------------------------------------- */

int gsf_abi_version(void);
/* Number of CUDA devices visible (0 if none / no driver). */
int gsf_device_count(void);
/* Devices the host-memory entry points shard points over (contiguous ranges, no collective).
 * n == 0 restores the default: env GSF_DEVICES="0,1,.." if set, else device 0 only. */
int gsf_set_devices(const int *device_ids, int n);
/* Pinned (page-locked) host memory from a caching pool, for result arrays: a D2H into such a
 * buffer needs no staging copy.  gsf_host_free returns the block to the pool.  Page-locked memory
 * cannot be swapped, so the pool refuses (GSF_ERR_ALLOC) once the blocks handed out exceed
 * GSF_PINNED_LIVE_MB (default min(4 GiB, RAM/8)); freed blocks are cached up to GSF_PINNED_CACHE_MB
 * (default 2048: two results of the largest size the Python module pins, so that the usual
 * `field = summate(...)` rebinding in a loop never reaches cudaMallocHost). */
int gsf_host_alloc(int64_t bytes, void **ptr);
int gsf_host_free(void *ptr);
/* Page-lock / unlock a caller-owned host range (cudaHostRegister): positions that are evaluated
 * again and again (ensembles over one mesh) can then be read by the GPU in place, without the
 * staging copy pageable memory needs.  Registering costs ~0.3 ms per MB once; the caller keeps the
 * range allocated until gsf_host_unregister. */
int gsf_host_register(void *ptr, int64_t bytes);
int gsf_host_unregister(void *ptr);
/* The contiguous point range [begin, end) that shard `shard` of `n_shards` owns -- the partition
 * the library uses across devices and bench.py uses across ranks (one process per GPU). */
int gsf_shard_bounds(int64_t n_points, int n_shards, int shard, int64_t *begin, int64_t *end);
/* Points per pipeline chunk for host-resident data (0 = default). */
int gsf_set_chunk_points(int64_t chunk_points);
/* Force the kernel variant (0,0 = heuristic): P points per thread, L lanes sharing one point's
 * modes.  Built for dim <= 3: (P, 1) with P in {1,2,3,4}, and (1 or 2, L) with L in
 * {2,4,8,16,32}; for dim 4..8: (1,1), (1,4), (1,32).  Anything else returns GSF_ERR_ARG (checked
 * against the dim-3 table here; a forced pair a later call's dim lacks falls back to the heuristic). */
int gsf_set_variant(int points_per_thread, int lanes_per_point);
/* Degree of the cosine polynomial in the point x mode kernels: 0 = automatic (6 for small problems,
 * 5 -- one DFMA cheaper, |error| <= 2.2e-13 per term, measured <= 1e-12 sigma -- from 2^27
 * point*modes), or 5 / 6 to force one.  One degree per call; never changes with chunking or sharding.
 * Environment: GSF_POLY_DEGREE. */
int gsf_set_poly_degree(int degree);
/* CUDA-event timing of the summation kernels (fills kernel_ms / prep_ms; the event sync happens
 * only in gsf_get_last_stats).  0 off; 1 per call; 2 accumulate: kernel_ms of gsf_get_last_stats is
 * then the sum over every call since this setting was made (bench.py: kernel time and step time
 * from one and the same loop). */
int gsf_set_profiling(int enabled);
int gsf_get_last_stats(gsf_stats *out);
const char *gsf_last_error(void);
/* Frees all device / pinned memory and streams.  The context is re-created on the next call. */
int gsf_shutdown(void);

/* Micro-benchmark: sustained FP64 DFMA issue rate of `device` in DFMA (thread-level) per second,
 * measured with CUDA events over ~`min_ms` milliseconds of dependent-chain DFMAs at full
 * occupancy.  This is the measured roofline denominator. */
int gsf_dfma_peak(int device, double min_ms, double *dfma_per_s, double *elapsed_ms);
/* Same for the FP64 tensor path (DMMA.8x8x4 streams), in thread-level FMA per second: the
 * roofline denominator of the structured-grid GEMM kernel. */
int gsf_dmma_peak(int device, double min_ms, double *fma_per_s, double *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* GSFIELD_H */
