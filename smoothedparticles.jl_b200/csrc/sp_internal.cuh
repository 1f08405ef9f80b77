// sp_internal.cuh — host-side system state and shared device helpers of libsp_b200.so.
//
// HBM layout (all Float64 unless stated; capacity `cap` is a multiple of 128 so every plane is
// 1 KiB aligned):
//   field f with ncomp components: planes  f.d[c*cap + s], s = device slot (true SoA)
//   ref[s]   int32  reference index (0-based position in the reference's sys.particles) of slot s
//   key[s]   int32  1-based linear cell key of slot s at the last create_cell_list
//   cell_start[k] int32, k = 1..key_max+1: cell k owns slots [cell_start[k], cell_start[k+1]);
//            k = key_max+1 is the "trash" cell of particles outside the domain (dropped by the build)
// After sp_create_cell_list the slots are sorted by (key ascending, ref descending): the slot order
// inside a cell IS the reference's descending-index order of Cell.entries (src/core.jl:26-41).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/sp_b200.h"

#define SP_MAX_FIELDS 64
// device counters (ints) of sp_system::counters / the pinned mirror h_counters
#define SP_CNT_TRASH 0    /* particles in the trash cell of the build in flight (dead tail + newly culled) */
#define SP_CNT_ALIVE 1    /* alive slots: [0, alive) */
#define SP_CNT_REMOVED 2  /* particles culled by all builds so far (sp_num_removed) */
#define SP_CNT_CULLED 3   /* particles culled by the last build */
#define SP_CNT_LOST 4     /* slab systems: particles that left the local cell window although they are inside the global
                             box, i.e. moved more than the ghost width between two rebuilds (cumulative) */
#define SP_CNT_CGFLAG 32
#define SP_CNT_NBRMAX 40
#define SP_FLAG_INTERNAL_PR_READY (1 << 30) /* library-internal: _pr = P/rho^2 is already up to date */

struct SpField {
    std::string name;
    int ncomp = 0;
    double* d = nullptr;    // current planes
    double* alt = nullptr;  // permutation target (swapped with d by the cell-list build)
    bool transient = false;  // solver scratch: contents need not survive a cell-list rebuild
    long long version = 1;   // bumped by every call that may write the field (host upload, operators, halos)
    bool known_zero = false; // every slot holds +0.0 (set by the operators that reset a field): the cell-list
                             // build has nothing to permute for such a field
};

// Parameters every device kernel needs about the cell grid (passed by value).
struct SpGrid {
    double h;       // neighbour radius
    double T2;      // largest double t with sqrt_rn(t) <= h:  (r > h) <=> (d2 > T2)  for r = sqrt_rn(d2)
    double lo[3], hi[3];
    long long phase[3];  // key_phase
    long long lim[3];    // key_lim
    long long key_max;
    int dim;  // 2 iff key_lim[2] == 1 (src/structs.jl:70)
    // slab decomposition (sp_slab.cu): phase/lim above describe the LOCAL cell window (owned layers + one ghost
    // layer per side along slab_axis); lo/hi stay the GLOBAL domain box.
    int slab_axis;      // -1 = not a slab system
    int slab_periodic;  // the slab axis wraps: no domain test along it
};

struct SlabState;  // sp_slab.cu

struct sp_system {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;    // per-call timing
    cudaEvent_t tev0 = nullptr, tev1 = nullptr;  // sp_timer_start/stop
    cudaEvent_t ev_count = nullptr;              // the removed-count read-back of the cell-list build
    SpGrid g{};
    int n_key_diff = 0;
    long long key_diff[27]{};

    // Particle count.  `n` is what the host launches with: slots [0, n) exist.  The number of ALIVE slots lives on the
    // device (counters[SP_CNT_ALIVE]): a cell-list build sorts the particles it culled (outside the domain, NaN) behind
    // the alive ones and lowers that count without telling the host, so the step loop has no device->host read-back
    // (and can be captured in a CUDA graph).  Slots [alive, n) are the dead tail: every kernel leaves at `slot >= alive`.
    // `n_exact` says the host knows alive == n; sp_settle() makes it so (one stream synchronisation) and is called by the
    // entry points that hand particle counts or particle data to the host.
    long long n = 0;
    bool n_exact = true;
    bool count_pending = false;  // a build's alive count is on its way to h_counters (adopted lazily, without waiting)
    long long cap = 0;  // plane stride
    std::vector<SpField> fields;
    int *ref = nullptr, *ref_alt = nullptr;
    int *key = nullptr, *key_alt = nullptr;
    int* cell_start = nullptr;  // key_max + 3 ints
    int* cell_fill = nullptr;   // key_max + 3 ints (scatter cursors)
    int* perm = nullptr;        // cap ints
    int* tmp_slot = nullptr;    // cap ints
    int* flags = nullptr;       // cap ints (cull flags / scratch)
    int* scan_tmp = nullptr;    // block sums for the scan
    long long scan_tmp_len = 0;
    int* counters = nullptr;    // small device counters
    int* h_counters = nullptr;  // pinned mirror
    double* stage = nullptr;    // upload/download staging + reduction scratch
    long long stage_len = 0;    // in doubles
    float* ucoord = nullptr;    // 3 planes of FP32 cell-unit coordinates (sweep pre-filter)
    long long ucoord_cap = 0;
    long long x_version = 1;       // bumped whenever positions or the slot order may have changed
    long long ucoord_version = 0;  // x_version the ucoord planes were computed for
    int* nbr_ids = nullptr;        // cached neighbour lists (sp_sweep.cu): cap/32 warp tiles x CAPK x 32 slots
    int* nbr_cnt = nullptr;        // neighbours per target slot
    long long nbr_cap = 0;
    long long nbr_version = 0;     // x_version the lists were built for
    long long nbr_n = 0;
    int nbr_capk = 64;             // list entries per target (multiple of 32; grows when a build reports more)
    bool nbr_max_pending = false;  // a build's longest-list report is on its way to h_counters[40]
    cudaEvent_t ev_nbr = nullptr;
    // what the last balance_of_mass sweep left in the scratch fields _kx/_kv (sp_ops.cuh: OpBalanceOfMassAux)
    struct {
        bool valid = false;
        long long x_version = 0, v_version = 0, n = 0;
        int v_fid = -1, kernel = -1, f_kx = -1, f_kv = -1;
        double m = 0, h = 0;
    } pair_aux;
    double* ell_val = nullptr;  // ISPH: Poisson-operator coefficients in the neighbour-list layout (sp_isph.cu)
    long long ell_cap = 0;
    int ell_capk = 0;
    double* dscal = nullptr;    // CG scalars + dot partials (3*1024 + 16 doubles)
    double* h_scal = nullptr;   // pinned mirror of a few scalars

    // CUDA graph of a unit of time steps of a step program (sp_program.cu)
    struct StepGraph {
        cudaGraphExec_t exec = nullptr;
        unsigned long long sig = 0;  // signature of everything the captured launches depend on (sp_program.cu)
        int program = 0, unit = 0;
        long long launches = 0;      // kernel launches per replay
        double params[16] = {0};
        int32_t fields[8] = {0};
    } graph;
    // graphs recorded by the host (sp_graph_begin / sp_graph_end)
    std::vector<StepGraph> user_graphs;
    unsigned long long rec_sig = 0;  // signature at sp_graph_begin
    long long rec_launches = 0;
    bool recording = false;

    bool have_cells = false;
    bool capturing = false;   // the stream is being captured into a CUDA graph (sp_program.cu): no host read-backs
    bool in_program = false;  // inside sp_run_program: nested entry points do not touch the per-call timing events
    bool identity_order = true;  // slot s holds reference particle s
    long long n_removed = 0;
    long long last_culled = 0;  // particles dropped by the last cell-list build
    long long launches = 0;
    float last_ms = 0.f;
    std::string err;
    SlabState* slab = nullptr;
};

extern thread_local std::string g_sp_create_error;

int sp_fail(sp_system* s, int code, const std::string& msg);
int sp_fail_cuda(sp_system* s, cudaError_t e, const char* what, const char* file, int line);

#define SP_CUDA(sys, call)                                                             \
    do {                                                                               \
        cudaError_t _e = (call);                                                       \
        if (_e != cudaSuccess) return sp_fail_cuda((sys), _e, #call, __FILE__, __LINE__); \
    } while (0)

// entry points that wait for the device cannot be recorded into a step graph
#define SP_NOT_WHILE_RECORDING(sys)                                                                              \
    do {                                                                                                         \
        if ((sys)->capturing)                                                                                    \
            return sp_fail((sys), SP_ERR_STATE, "this call waits for the device and cannot be recorded into a step graph"); \
    } while (0)

// launch + count + check
#define SP_LAUNCH(sys, kernel, grid, block, smem, ...)                                       \
    do {                                                                                     \
        kernel<<<(grid), (block), (smem), (sys)->stream>>>(__VA_ARGS__);                     \
        (sys)->launches++;                                                                   \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess) return sp_fail_cuda((sys), _e, #kernel, __FILE__, __LINE__);  \
    } while (0)

static inline unsigned sp_blocks(long long n, int block) { return (unsigned)((n + block - 1) / block); }

// Device memory comes from the device's stream-ordered pool with the release threshold lifted, i.e. a caching
// allocator: sp_destroy hands the blocks back to the pool (not to the driver) and the next sp_create in the same
// process reuses them.  A 10 M-particle system is ~6 GB in ~60 blocks; cudaFree of those costs ~0.2 s, the pool ~1 ms.
cudaError_t sp_dmalloc_impl(void** p, size_t bytes);
template <class T>
static inline cudaError_t sp_dmalloc(T** p, size_t bytes) {
    return sp_dmalloc_impl(reinterpret_cast<void**>(p), bytes);
}
// the caller guarantees that no work using p is still in flight (sp_dfree(sys, p) waits for the system's stream)
cudaError_t sp_dfree_impl(void* p);
static inline cudaError_t sp_dfree(sp_system* s, void* p) {
    if (!p) return cudaSuccess;
    if (s && s->stream) {
        cudaError_t e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) return e;
    }
    return sp_dfree_impl(p);
}

// grow-only device scratch
int sp_ensure_stage(sp_system* s, long long doubles);
int sp_ensure_capacity(sp_system* s, long long n);
// exclusive scan of `len` ints in place (device), sp_cells.cu
int sp_exclusive_scan_i32(sp_system* s, int* data, long long len);
// validate a field binding
int sp_check_fields(sp_system* s, const int32_t* fields, int nfields, const int* ncomps, int nexpected);

// RAII-less timing helpers
int sp_time_begin(sp_system* s);
int sp_time_end(sp_system* s);
// bookkeeping for the per-field caches: call BEFORE launching anything that writes field `fid`
static inline void sp_wrote(sp_system* s, int fid) {
    SpField& f = s->fields[fid];
    f.version++;
    f.known_zero = false;
    if (fid == 0) s->x_version++;
}
static inline void sp_zeroed(sp_system* s, int fid) {
    sp_wrote(s, fid);
    s->fields[fid].known_zero = true;
}
void sp_slab_free(sp_system* s);  // sp_slab.cu
void sp_program_free(sp_system* s);  // sp_program.cu: drops the cached step graph
int sp_build_cells(sp_system* s);  // sp_cells.cu
// make the host's particle count exact (waits for the stream if a build's count has not been fetched yet)
int sp_settle(sp_system* s);
// the host changed the particle count itself (resize, slab arrivals): publish it to the device counter
int sp_publish_count(sp_system* s);
static inline const int* sp_alive(const sp_system* s) { return s->counters + SP_CNT_ALIVE; }
// |u| bound of the FP32 pre-filter coordinates (sp_sweep.cu: sp_ensure_prefilter), also used by the cell-list permute
float sp_prefilter_range(const sp_system* s);
int sp_self_visits(const sp_system* s);  // sp_sweep.cu: visits of the own cell by the stencil (diagonal multiplicity)
void sp_slab_host_touched(sp_system* s);  // positions / particle set changed by the host: full selection next time
const double* sp_slab_ghost_mask(sp_system* s);  // nullptr unless a slab system: 0 = owned, 1/2 = ghost
const double* sp_slab_gid(sp_system* s);         // nullptr unless a slab system: the global id plane (in-cell order)

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__

// find_key, src/structs.jl:97-106: true IEEE division and floor; 1-based linear key.
// Returns -1 when a coordinate is NaN/Inf (Int64(floor(.)) would throw).
__device__ __forceinline__ long long sp_find_key(const SpGrid& g, double x, double y, double z) {
    double q0 = floor(__ddiv_rn(x, g.h)), q1 = floor(__ddiv_rn(y, g.h)), q2 = floor(__ddiv_rn(z, g.h));
    if (!(fabs(q0) < 9.0e18) || !(fabs(q1) < 9.0e18) || !(fabs(q2) < 9.0e18)) return -1;
    long long i = 1 + (long long)q0 - g.phase[0];
    long long j = 1 + (long long)q1 - g.phase[1];
    long long k = 1 + (long long)q2 - g.phase[2];
    return i + g.lim[0] * (j - 1) + g.lim[0] * g.lim[1] * (k - 1);
}

// is_inside(x, Box), src/geometry.jl:24-30 (closed; NaN -> false)
__device__ __forceinline__ bool sp_inside(const SpGrid& g, double x, double y, double z) {
    if (g.slab_axis < 0)
        return g.lo[0] <= x && x <= g.hi[0] && g.lo[1] <= y && y <= g.hi[1] && g.lo[2] <= z && z <= g.hi[2];
    // slab system: global box on the other axes (and on the slab axis unless it is periodic), plus the local
    // cell window along the slab axis (anything else has been migrated away before the build)
    const double p[3] = {x, y, z};
    bool in = true;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (a == g.slab_axis) {
            if (!g.slab_periodic) in = in && g.lo[a] <= p[a] && p[a] <= g.hi[a];
            const double q = floor(__ddiv_rn(p[a], g.h));
            in = in && (q >= (double)g.phase[a]) && (q < (double)(g.phase[a] + g.lim[a]));
        } else
            in = in && g.lo[a] <= p[a] && p[a] <= g.hi[a];
    }
    return in;
}

// squared distance exactly as dist() rounds it before the sqrt: (dx*dx + dy*dy) + dz*dz, no FMA
// (src/core.jl:8-10, src/algebra.jl:49-60).
__device__ __forceinline__ double sp_d2(double dx, double dy, double dz) {
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

#endif
