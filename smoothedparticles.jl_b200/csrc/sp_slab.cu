// sp_slab.cu — slab decomposition of one ParticleSystem over the GPUs of a node (one process per GPU).
//
// The reference is single-process shared memory; this is the multi-GPU extension SURVEY §8(e) specifies.
// The global cell grid of sp_create is cut along the SLOWEST key axis (z in 3-D, y in 2-D) into `nranks`
// slabs of whole cell layers.  Because that axis is the slowest in the linear key, after the local sort a
// rank's two boundary cell layers and its two ghost layers are contiguous slot ranges.
//
// A rank's LOCAL cell window is its owned layers [c0, c1) plus SLAB_W = 2 ghost layers per side.  Two layers, because
// then the inner ghost layer sees all of its own neighbours: balance_of_mass! integrates its density exactly as the owner
// does, and the WCSPH step needs ONE exchange per time step (no refresh of rho / P after find_pressure!).
//
// sp_slab_create_cell_list (replaces create_cell_list! on a slab system) — ONE exchange round, no host synchronisation
// in the steady state:
//   1. last step's ghosts are dropped (they come back fresh);
//   2. every owned particle whose CURRENT position lies in the rank's first / last two owned layers or beyond them goes
//      into the message to the lower / upper neighbour — ghost copies and migrants alike, every field, in slot order
//      (deterministic: block counts, one scan, ordered pack).  Ownership is a function of the position alone: the
//      receiver owns what falls into its owned layers and keeps the rest of its window as ghosts; the sender keeps
//      its copy of a particle that migrated away as a ghost (it is bit-identical to what the new owner holds);
//   3. both messages travel at a CAPACITY known to both ends without talking: twice the larger of the counts that went
//      over the same link two and three rebuilds ago (both ends have them); the actual count rides in the message header
//      and stays on the device.  On a link both ends mapped through CUDA IPC the pack kernel stores straight into the
//      neighbour's receive buffer over NVLink and flag words say "message k is there" / "message k is consumed"; any
//      other link goes through one ncclGroup of send/recv at the full capacity (unused entries become dead slots);
//   4. arrivals are appended behind the alive slots, the ordinary cell-list build (sp_cells.cu, no read-back) sorts
//      everything, "_ghost" is set from the cell layer.
// The host runs at most two rebuilds ahead of the device: rebuild b waits for the counts of rebuild b-2 (an event that
// has long completed when the device is busy), takes the message capacities and the slot bound from them, and never
// waits for anything younger.  The first two rebuilds after sp_slab_init or after the host changed particles exchange
// their counts explicitly (one host synchronisation each).
// In-cell order on a slab system is by descending "_gid" (a global id given at insertion), which is the same on every
// rank: a rank's two boundary layers and its neighbour's two ghost layers are then IDENTICAL slot sequences, and
// sp_slab_halo_refresh (needed by the ISPH CG, whose search vector changes every iteration) is a plain copy of slot
// ranges: pack, one ncclGroup, unpack — no index fields.
// Reductions and CG dot products skip ghosts and are summed with ncclAllReduce.
//
// NCCL is loaded lazily with dlopen, so the single-GPU library has no NCCL dependency.
#include <dlfcn.h>
#if defined(__has_include)
#if __has_include(<nccl.h>)
#include <nccl.h>
#define SP_HAVE_NCCL_H 1
#endif
#endif
#ifndef SP_HAVE_NCCL_H
// Building without the NCCL headers: the few declarations this file uses (NCCL is only ever reached through dlopen, so
// a machine without NCCL builds and runs the single-GPU library; sp_slab_* then fails with SP_ERR_NCCL at run time).
extern "C" {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclInt32 = 2, ncclInt64 = 4, ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
}
#endif

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sp_internal.cuh"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};
NcclApi g_nccl;

bool nccl_load() {
    if (g_nccl.handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) {
        g_nccl.err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
        return false;
    }
#define LOAD(sym)                                                                \
    *(void**)(&g_nccl.sym) = dlsym(g_nccl.handle, "nccl" #sym);                  \
    if (!g_nccl.sym) {                                                           \
        g_nccl.err = "libnccl lacks nccl" #sym;                                  \
        g_nccl.handle = nullptr;                                                 \
        return false;                                                            \
    }
    LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(Send) LOAD(Recv) LOAD(GroupStart) LOAD(GroupEnd)
    LOAD(AllReduce) LOAD(GetErrorString)
#undef LOAD
    return true;
}

}  // namespace

#define SLAB_W 2        /* ghost layers per side */
#define SLAB_NB 512     /* blocks of the selection passes */
#define SLAB_HDR 8      /* doubles in front of a message: [0] = particle count */
#define SLAB_RING 4

struct SlabState {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, periodic = 0, axis = 2;
    long long gphase = 0, glim = 1;  // global key_phase / key_lim along the axis
    long long c0 = 0, c1 = 1;        // owned global cell layers [c0, c1), 0-based from gphase
    double period = 0.0;
    int f_ghost = -1, f_gid = -1;
    double* sendbuf[2] = {nullptr, nullptr};
    double* recvbuf[2] = {nullptr, nullptr};
    long long buf_len = 0;  // doubles per buffer
    // device ints: [0],[1] particles sent down / up, [2],[3] received from below / above, [4] overflow flag,
    // [5] refresh mismatch flag, [8] owned count, [16 ..] block counts / offsets of the selection passes
    int* d_cnt = nullptr;
    int* h_cnt = nullptr;                 // pinned: SLAB_RING x 16 ints (counts of the last rebuilds) + scratch
    cudaEvent_t ring_ev[SLAB_RING] = {nullptr, nullptr, nullptr, nullptr};
    long long build_no = 0;               // rebuilds so far
    int history = 0;                      // consecutive rebuilds whose counts are in the ring (steady state from 2 on)
    long long cap_send[2] = {0, 0};       // message capacities of the last rebuild: down / up
    long long cap_recv[2] = {0, 0};       // ... from below / from above
    long long gid_base = 0;
    bool fresh_particles = true;          // the host added particles: they have no "_gid" yet
    // ---- peer-to-peer exchange over NVLink (default when CUDA IPC works; SP_SLAB_P2P=0 keeps every link on NCCL).
    // One cudaMalloc'ed block per rank: [64 flag words][receive buffer "from below"][receive buffer "from above"]; the
    // neighbours map it through CUDA IPC and WRITE their messages straight into it from the pack kernel (no staging
    // buffer, no padding: the exact count is in the header), then raise a flag word; acknowledgements come back the same way.
    unsigned long long* p2p_block = nullptr;       // my block (device memory)
    unsigned long long* p2p_peer[2] = {nullptr, nullptr};  // the block of the rank below / above as mapped here
    bool p2p_mapped[2] = {false, false};           // p2p_peer[d] came from cudaIpcOpenMemHandle (close it when done)
    long long p2p_cap = 0, p2p_planes = 0;         // my receive buffers: entries per plane, planes
    long long p2p_peer_cap[2] = {0, 0}, p2p_peer_planes[2] = {0, 0};
    long long p2p_send_seq[2] = {0, 0}, p2p_recv_seq[2] = {0, 0};
    int p2p_state = 0;                             // 0 = not tried, 1 = ready, -1 = unavailable (NCCL only)
    double trace_s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long trace_calls = 0;
};

// how long a flag wait may spin, in SM clocks.  Generous on purpose: a neighbour that is merely late (module loading, a
// host thread descheduled) must not be mistaken for one that stopped.
static long long slab_wait_limit(const sp_system* s) {
    static long long clocks = 0;
    if (!clocks) {
        double seconds = 30.0;
        if (const char* e = getenv("SP_SLAB_TIMEOUT_S"))
            if (atof(e) > 0.0) seconds = atof(e);
        int khz = 0;
        if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, s->device) != cudaSuccess || khz <= 0) khz = 2000000;
        clocks = (long long)(seconds * 1000.0 * (double)khz);
    }
    return clocks;
}
static bool slab_trace_on() {
    static const bool on = getenv("SP_SLAB_TRACE") && atoi(getenv("SP_SLAB_TRACE"));
    return on;
}
#define SP_NCCL(s, call)                                                                                  \
    do {                                                                                                  \
        ncclResult_t _r = (call);                                                                         \
        if (_r != ncclSuccess)                                                                            \
            return sp_fail((s), SP_ERR_NCCL, std::string("NCCL: ") + g_nccl.GetErrorString(_r) + " in " #call); \
    } while (0)

void sp_slab_free(sp_system* s) {
    SlabState* sl = s->slab;
    if (!sl) return;
    if (sl->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(sl->comm);
    for (int d = 0; d < 2; d++) {
        sp_dfree(s, sl->sendbuf[d]);
        sp_dfree(s, sl->recvbuf[d]);
    }
    sp_dfree(s, sl->d_cnt);
    for (int d = 0; d < 2; d++)
        if (sl->p2p_mapped[d] && sl->p2p_peer[d] && !(d == 1 && sl->p2p_peer[1] == sl->p2p_peer[0])) cudaIpcCloseMemHandle(sl->p2p_peer[d]);
    if (sl->p2p_block) {
        if (s->stream) cudaStreamSynchronize(s->stream);
        cudaFree(sl->p2p_block);
    }
    cudaGetLastError();
    if (sl->h_cnt) cudaFreeHost(sl->h_cnt);
    for (int r = 0; r < SLAB_RING; r++)
        if (sl->ring_ev[r]) cudaEventDestroy(sl->ring_ev[r]);
    delete sl;
    s->slab = nullptr;
}

void sp_slab_host_touched(sp_system* s) {
    if (!s->slab) return;
    s->slab->history = 0;  // the next two rebuilds exchange their counts explicitly and select over all slots
    s->slab->fresh_particles = true;
}

const double* sp_slab_ghost_mask(sp_system* s) {
    if (!s->slab) return nullptr;
    return s->fields[s->slab->f_ghost].d;
}
const double* sp_slab_gid(sp_system* s) {
    if (!s->slab) return nullptr;
    return s->fields[s->slab->f_gid].d;
}

int sp_slab_allreduce_device(sp_system* s, double* d_inout, int count, int is_max) {
    if (!s->slab || s->slab->nranks == 1) return SP_OK;
    SP_NCCL(s, g_nccl.AllReduce(d_inout, d_inout, (size_t)count, ncclFloat64, is_max ? ncclMax : ncclSum, s->slab->comm,
                                s->stream));
    return SP_OK;
}

// ------------------------------------------------------------------ kernels
#define SLAB_PLANES 40
struct SlabPlanes {
    double* p[SLAB_PLANES];
    int count;
    int axis_plane;  // index of the plane holding the slab-axis coordinate, or -1
    int x_plane;     // index of the plane holding x[0] (a NaN there kills a slot), or -1
};

// Where things are in the slot order of the LAST build (slots are sorted by cell, the slab axis is the slowest key axis,
// so a cell layer is one slot range).  `full` = the host touched the particles: no valid cell list, look at every slot.
struct SlabWin {
    const int* cell_start;
    const int* counters;   // sp_system::counters
    long long L;           // cells per layer
    int nl;                // layers of the local window (owned + 2 * SLAB_W)
    int full;
    __device__ long long layer_begin(int l) const { return cell_start[(long long)l * L + 1]; }
    __device__ long long alive() const { return counters[SP_CNT_ALIVE]; }
    // slots that may have to be sent in direction dir (0 = down, 1 = up): the three owned layers next to that side
    __device__ void send_window(int dir, long long* lo, long long* hi) const {
        if (full) {
            *lo = 0;
            *hi = alive();
            return;
        }
        if (dir == 0) {
            *lo = layer_begin(SLAB_W);
            *hi = layer_begin(min(SLAB_W + 3, nl - SLAB_W));
        } else {
            *lo = layer_begin(max(SLAB_W, nl - SLAB_W - 3));
            *hi = layer_begin(nl - SLAB_W);
        }
    }
};

__device__ __forceinline__ bool slab_layer_of(const SpGrid& g, double xa, long long* layer) {
    const double q = floor(__ddiv_rn(xa, g.h));
    if (!(q == q) || !(fabs(q) < 9.0e18)) return false;
    *layer = (long long)q - g.phase[g.slab_axis];
    return true;
}
// is slot s part of the message in direction dir?  owned, alive, and its CURRENT layer within the first (last) two owned
// layers or beyond
__device__ __forceinline__ bool slab_selected(const SpGrid& g, const double* x, long long cap, const double* ghost, int nl,
                                              int dir, long long s) {
    if (ghost[s] != 0.0) return false;
    const double x0 = x[s];
    if (!(x0 == x0)) return false;
    long long l;
    if (!slab_layer_of(g, x[(size_t)g.slab_axis * cap + s], &l)) return false;
    return dir == 0 ? (l < 2 * SLAB_W) : (l >= nl - 2 * SLAB_W);
}

// old ghosts die: x = NaN (the build culls them)
__global__ void k_slab_kill_ghosts(SlabWin w, double* x, const double* ghost) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (w.full) {
        const long long n = w.alive();
        for (long long s = t0; s < n; s += stride)
            if (ghost[s] != 0.0) x[s] = nan("");
        return;
    }
    const long long a0 = w.layer_begin(0), a1 = w.layer_begin(SLAB_W);
    const long long b0 = w.layer_begin(w.nl - SLAB_W), b1 = w.layer_begin(w.nl);
    for (long long s = a0 + t0; s < a1; s += stride) x[s] = nan("");
    for (long long s = b0 + t0; s < b1; s += stride) x[s] = nan("");
}

// pass 1: how many selected slots in the chunk of block b (blockIdx.y = direction)
__global__ void __launch_bounds__(256) k_slab_sel_count(SpGrid g, SlabWin w, const double* x, long long cap, const double* ghost,
                                                        int* blk) {
    const int dir = blockIdx.y;
    long long lo, hi;
    w.send_window(dir, &lo, &hi);
    const long long len = max(hi - lo, 0LL);
    const long long chunk = (len + gridDim.x - 1) / gridDim.x;
    const long long b = lo + (long long)blockIdx.x * chunk, e = min(b + chunk, hi);
    int mine = 0;
    for (long long s = b + threadIdx.x; s < e; s += blockDim.x) mine += slab_selected(g, x, cap, ghost, w.nl, dir, s) ? 1 : 0;
    __shared__ int sm[256];
    sm[threadIdx.x] = mine;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if (threadIdx.x < d) sm[threadIdx.x] += sm[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) blk[dir * SLAB_NB + blockIdx.x] = sm[0];
}
// pass 2 (one block): exclusive scan of the block counts, totals, overflow against the message capacity
__global__ void __launch_bounds__(SLAB_NB) k_slab_sel_scan(int* blk, int* d_cnt, long long cap_dn, long long cap_up) {
    __shared__ int sm[SLAB_NB];
    for (int dir = 0; dir < 2; dir++) {
        const int v = blk[dir * SLAB_NB + threadIdx.x];
        sm[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < SLAB_NB; d <<= 1) {
            int t = threadIdx.x >= d ? sm[threadIdx.x - d] : 0;
            __syncthreads();
            sm[threadIdx.x] += t;
            __syncthreads();
        }
        blk[dir * SLAB_NB + threadIdx.x] = sm[threadIdx.x] - v;  // exclusive
        if (threadIdx.x == SLAB_NB - 1) {
            const long long capd = dir == 0 ? cap_dn : cap_up;
            const int total = sm[threadIdx.x];
            if (capd >= 0 && total > capd) {
                d_cnt[4] = 1;  // overflow: the message cannot hold the boundary particles (reported by a later rebuild)
                d_cnt[dir] = (int)capd;
            } else
                d_cnt[dir] = total;
        }
        __syncthreads();
    }
}
// pass 3: ordered pack of every plane of the selected slots into the message (plane c at SLAB_HDR + c*msg_cap)
__global__ void __launch_bounds__(256) k_slab_sel_pack(SpGrid g, SlabWin w, const double* x, long long cap, const double* ghost,
                                                       const int* blk, const int* d_cnt, SlabPlanes tab, int plane0,
                                                       double* buf_dn, long long stride_dn, double* buf_up, long long stride_up) {
    // buf_*: the NCCL staging buffer, or — on a peer-to-peer link — the neighbour's receive buffer itself (remote stores
    // over NVLink); stride_* = entries per plane of that buffer
    const int dir = blockIdx.y;
    double* buf = dir == 0 ? buf_dn : buf_up;
    const long long stride = dir == 0 ? stride_dn : stride_up;
    if (!buf) return;
    const long long mcap = d_cnt[dir];  // the (clamped) message count
    long long lo, hi;
    w.send_window(dir, &lo, &hi);
    const long long len = max(hi - lo, 0LL);
    const long long chunk = (len + gridDim.x - 1) / gridDim.x;
    const long long b = lo + (long long)blockIdx.x * chunk, e = min(b + chunk, hi);
    if (blockIdx.x == 0 && threadIdx.x == 0 && plane0 == 0) buf[0] = (double)d_cnt[dir];
    __shared__ int warp_tot[8];
    __shared__ int run_sm;
    if (threadIdx.x == 0) run_sm = blk[dir * SLAB_NB + blockIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = b; base < e; base += blockDim.x) {
        const long long s = base + threadIdx.x;
        const bool sel = s < e && slab_selected(g, x, cap, ghost, w.nl, dir, s);
        const unsigned bal = __ballot_sync(0xffffffffu, sel);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int off = run_sm;
        for (int q = 0; q < warp; q++) off += warp_tot[q];
        const long long m = off + __popc(bal & ((1u << lane) - 1u));
        if (sel && m < mcap)
            for (int c = 0; c < tab.count; c++) buf[SLAB_HDR + (size_t)(plane0 + c) * stride + m] = tab.p[c][s];
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int q = 0; q < 8; q++) t += warp_tot[q];
            run_sm += t;
        }
        __syncthreads();
    }
}
// Arrival layout, decided on the device from the two message headers: d_cnt[2], d_cnt[3] = particles from below / above
// (clamped to what the unpack launches cover: an excess raises the overflow flag), d_cnt[13] = slot offset of the second
// message, d_cnt[14] = slots appended in total.  A message that came through NCCL occupies its full capacity (padding
// becomes dead slots), a peer-to-peer message exactly its count.
__global__ void k_slab_arrival_layout(const double* hdr_lo, long long cap_lo, int lo_padded, const double* hdr_hi, long long cap_hi,
                                      int hi_padded, int* d_cnt) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long c_lo = hdr_lo ? (long long)hdr_lo[0] : 0, c_hi = hdr_hi ? (long long)hdr_hi[0] : 0;
    if (c_lo > cap_lo || c_hi > cap_hi || c_lo < 0 || c_hi < 0) d_cnt[4] = 1;
    c_lo = max(0LL, min(c_lo, cap_lo));
    c_hi = max(0LL, min(c_hi, cap_hi));
    d_cnt[2] = (int)c_lo;
    d_cnt[3] = (int)c_hi;
    const long long span_lo = hdr_lo ? (lo_padded ? cap_lo : c_lo) : 0;
    const long long span_hi = hdr_hi ? (hi_padded ? cap_hi : c_hi) : 0;
    d_cnt[13] = (int)span_lo;
    d_cnt[14] = (int)(span_lo + span_hi);
}
// arrivals go behind the alive slots: slot = alive + off + t; in a padded (NCCL) message the entries beyond the count
// become dead slots
__global__ void k_slab_unpack(SlabPlanes tab, int plane0, const double* buf, long long stride, long long launch_cap, int padded,
                              double shift, const int* counters, int* ref, const int* d_cnt, int which) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= launch_cap) return;
    const long long count = d_cnt[2 + which];
    const long long slot = (long long)counters[SP_CNT_ALIVE] + (which ? d_cnt[13] : 0) + t;
    if (t < count) {
        for (int c = 0; c < tab.count; c++) {
            double v = buf[SLAB_HDR + (size_t)(plane0 + c) * stride + t];
            if (c == tab.axis_plane) v += shift;
            tab.p[c][slot] = v;
        }
    } else if (padded) {
        // padding: a dead slot (NaN position, culled by the build).  Its other planes are cleared: a field the library
        // knows to be zero everywhere is not permuted by the build, so no stale value may sit in a slot below the new
        // alive count
        for (int c = 0; c < tab.count; c++) tab.p[c][slot] = c == tab.x_plane ? nan("") : 0.0;
    } else
        return;
    if (plane0 == 0) ref[slot] = (int)slot;
}
// ---- peer-to-peer flags.  Flag words (unsigned long long) at the head of a rank's block:
//   [0] sequence number of the newest message in my "from below" buffer   (written by the rank below)
//   [1] ... in my "from above" buffer                                      (written by the rank above)
//   [2] newest message of mine the rank below has consumed                 (written by the rank below)
//   [3] ... the rank above has consumed                                    (written by the rank above)
#define SLAB_P2P_FLAGS 64
// one thread waits until *flag >= value; gives up after `limit` device clocks (slab_wait_limit: 30 s unless
// SP_SLAB_TIMEOUT_S says otherwise) and raises d_cnt[6] (reported by a later rebuild) instead of hanging the GPU
// (two flags per launch — one per direction; a null pointer is skipped)
__global__ void k_slab_wait(const volatile unsigned long long* flag_a, unsigned long long value_a,
                            const volatile unsigned long long* flag_b, unsigned long long value_b, int* d_cnt, long long limit) {
    if (threadIdx.x > 1 || blockIdx.x != 0) return;
    const volatile unsigned long long* flag = threadIdx.x == 0 ? flag_a : flag_b;
    const unsigned long long value = threadIdx.x == 0 ? value_a : value_b;
    if (flag) {
        const long long t0 = clock64();
        while (*flag < value) {
            __nanosleep(200);
            if (clock64() - t0 > limit) {
                d_cnt[6] = 1;
                break;
            }
        }
    }
    __threadfence_system();
}
__global__ void k_slab_signal(volatile unsigned long long* flag_a, unsigned long long value_a, volatile unsigned long long* flag_b,
                              unsigned long long value_b) {
    if (threadIdx.x > 1 || blockIdx.x != 0) return;
    volatile unsigned long long* flag = threadIdx.x == 0 ? flag_a : flag_b;
    __threadfence_system();
    if (flag) *flag = threadIdx.x == 0 ? value_a : value_b;
    __threadfence_system();
}
__global__ void k_slab_zero_tail(SlabPlanes tab, long long launch_count, const int* d_cnt, const int* counters) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= launch_count || t >= d_cnt[14]) return;
    const long long slot = (long long)counters[SP_CNT_ALIVE] + t;
    for (int c = 0; c < tab.count; c++) tab.p[c][slot] = 0.0;
}
__global__ void k_slab_add_alive(int* counters, const int* d_cnt) {
    if (threadIdx.x == 0 && blockIdx.x == 0) counters[SP_CNT_ALIVE] += d_cnt[14];
}
__global__ void k_slab_clear(int* d_cnt) {
    if (threadIdx.x < 4 && blockIdx.x == 0) d_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 8 && blockIdx.x == 0) d_cnt[8] = 0;
    if ((threadIdx.x == 13 || threadIdx.x == 14) && blockIdx.x == 0) d_cnt[threadIdx.x] = 0;
}
// particles the host added have no global id yet: rank * 2^44 + base + slot (unique, deterministic)
__global__ void k_slab_assign_gid(double* gid, const int* counters, double base) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s < counters[SP_CNT_ALIVE] && gid[s] == 0.0) gid[s] = base + (double)s;
}
// after the build: "_ghost" from the cell layer (1 = below the owned layers, 2 = above), ref = slot, owned count
__global__ void __launch_bounds__(256) k_slab_post(const int* key, double* ghost, int* ref, long long L, int nl,
                                                   const int* counters, int* d_cnt) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int owned = 0;
    if (s < counters[SP_CNT_ALIVE]) {
        const long long l = ((long long)key[s] - 1) / L;
        const double gflag = l < SLAB_W ? 1.0 : (l >= nl - SLAB_W ? 2.0 : 0.0);
        ghost[s] = gflag;
        ref[s] = (int)s;
        owned = gflag == 0.0;
    }
    const int total = __syncthreads_count(owned);  // one atomic per CTA
    if (threadIdx.x == 0 && total) atomicAdd(d_cnt + 8, total);
}
// halo refresh: the two boundary layers of the owner and the two ghost layers of its neighbour are the same slot
// sequence (see the header comment), so a refresh is a copy of slot ranges
__global__ void k_slab_refresh_pack(SlabWin w, const double* f, long long cap, int ncomp, int comp0, double* buf_dn,
                                    long long cap_dn, double* buf_up, long long cap_up) {
    const int dir = blockIdx.y;
    double* buf = dir == 0 ? buf_dn : buf_up;
    if (!buf) return;
    const long long mcap = dir == 0 ? cap_dn : cap_up;
    const long long lo = dir == 0 ? w.layer_begin(SLAB_W) : w.layer_begin(w.nl - 2 * SLAB_W);
    const long long hi = dir == 0 ? w.layer_begin(2 * SLAB_W) : w.layer_begin(w.nl - SLAB_W);
    const long long len = min(hi - lo, mcap);
    if (blockIdx.x == 0 && threadIdx.x == 0) buf[0] = (double)(hi - lo);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < len; t += (long long)gridDim.x * blockDim.x)
        for (int c = 0; c < ncomp; c++) buf[SLAB_HDR + (size_t)(comp0 + c) * mcap + t] = f[(size_t)c * cap + lo + t];
}
__global__ void k_slab_refresh_unpack(SlabWin w, double* f, long long cap, int ncomp, int comp0, const double* buf_lo,
                                      long long cap_lo, const double* buf_hi, long long cap_hi, int axis_comp,
                                      double shift_lo, double shift_hi, int* d_cnt) {
    const int side = blockIdx.y;  // 0: from below into the lower ghost layers, 1: from above into the upper ones
    const double* buf = side == 0 ? buf_lo : buf_hi;
    if (!buf) return;
    const long long mcap = side == 0 ? cap_lo : cap_hi;
    const long long lo = side == 0 ? w.layer_begin(0) : w.layer_begin(w.nl - SLAB_W);
    const long long hi = side == 0 ? w.layer_begin(SLAB_W) : w.layer_begin(w.nl);
    const long long count = (long long)buf[0];
    if (count != hi - lo) {  // the two ends disagree about the boundary set: never expected, reported by the next rebuild
        if (blockIdx.x == 0 && threadIdx.x == 0) d_cnt[5] = 1;
        return;
    }
    const double shift = side == 0 ? shift_lo : shift_hi;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < count; t += (long long)gridDim.x * blockDim.x)
        for (int c = 0; c < ncomp; c++) {
            double v = buf[SLAB_HDR + (size_t)(comp0 + c) * mcap + t];
            if (c == axis_comp) v += shift;
            f[(size_t)c * cap + lo + t] = v;
        }
}

// ------------------------------------------------------------------ host helpers
static int slab_planes(sp_system* s, std::vector<SlabPlanes>& tabs, int* nplanes, bool skip_zero = true, bool only_zero = false) {
    tabs.clear();
    SlabPlanes cur;
    cur.count = 0;
    cur.axis_plane = cur.x_plane = -1;
    int total = 0;
    for (size_t f = 0; f < s->fields.size(); f++) {
        SpField& fl = s->fields[f];
        if (fl.transient) continue;
        // a field that is +0.0 in every slot (the operators that reset a field say so: Dv after move!, Drho after
        // find_pressure!) does not travel: the receiver's copy is zero as well — every rank runs the same call sequence —
        // and the dead tail the arrivals land in is kept zero for such fields (see k_slab_unpack / zero_tabs)
        if (skip_zero && fl.known_zero) continue;
        if (only_zero && !fl.known_zero) continue;
        for (int c = 0; c < fl.ncomp; c++) {
            if (f == 0 && c == s->slab->axis) cur.axis_plane = cur.count;
            if (f == 0 && c == 0) cur.x_plane = cur.count;
            cur.p[cur.count++] = fl.d + (size_t)c * s->cap;
            total++;
            if (cur.count == SLAB_PLANES) {
                tabs.push_back(cur);
                cur.count = 0;
                cur.axis_plane = cur.x_plane = -1;
            }
        }
    }
    if (cur.count) tabs.push_back(cur);
    *nplanes = total;
    return SP_OK;
}

static int slab_ensure_buffers(sp_system* s, long long doubles) {
    SlabState* sl = s->slab;
    if (doubles <= sl->buf_len) return SP_OK;
    const long long want = doubles + doubles / 4 + 4096;
    for (int d = 0; d < 2; d++) {
        if (sl->sendbuf[d]) SP_CUDA(s, sp_dfree(s, sl->sendbuf[d]));
        if (sl->recvbuf[d]) SP_CUDA(s, sp_dfree(s, sl->recvbuf[d]));
        sl->sendbuf[d] = sl->recvbuf[d] = nullptr;
    }
    sl->buf_len = 0;
    for (int d = 0; d < 2; d++) {
        SP_CUDA(s, sp_dmalloc(&sl->sendbuf[d], (size_t)want * sizeof(double)));
        SP_CUDA(s, sp_dmalloc(&sl->recvbuf[d], (size_t)want * sizeof(double)));
    }
    sl->buf_len = want;
    return SP_OK;
}

// neighbours along the slab axis (-1 = none)
static void slab_peers(const SlabState* sl, int* below, int* above) {
    *below = sl->rank - 1;
    *above = sl->rank + 1;
    if (sl->periodic) {
        *below = (sl->rank + sl->nranks - 1) % sl->nranks;
        *above = (sl->rank + 1) % sl->nranks;
    } else {
        if (*above >= sl->nranks) *above = -1;
    }
}

static SlabWin slab_win(sp_system* s, bool full) {
    SlabWin w;
    w.cell_start = s->cell_start;
    w.counters = s->counters;
    w.L = s->g.lim[0] * (s->slab->axis == 2 ? s->g.lim[1] : 1);
    w.nl = (int)s->g.lim[s->slab->axis];
    w.full = full ? 1 : 0;
    return w;
}

// The bootstrap rebuilds exchange the send counts (already in d_cnt[0], d_cnt[1]) with both neighbours and bring all
// four numbers to the host with ONE synchronisation: h[0],[1] = my counts down/up, h[2],[3] = counts arriving from
// below / above.
static int slab_exchange_counts(sp_system* s, long long* n_dn, long long* n_up, long long* from_below, long long* from_above) {
    SlabState* sl = s->slab;
    int below, above;
    slab_peers(sl, &below, &above);
    int* scratch = sl->d_cnt + 12;  // [12],[13] receive slots
    int* h = sl->h_cnt + SLAB_RING * 16;
    // Call order matters when below == above (1 or 2 ranks, periodic): messages between one pair of ranks are
    // matched in issue order, and what I send DOWN arrives at my lower neighbour FROM ABOVE.  So: send down,
    // send up, then receive from above, receive from below.
    SP_NCCL(s, g_nccl.GroupStart());
    if (below >= 0) SP_NCCL(s, g_nccl.Send(sl->d_cnt + 0, 1, ncclInt32, below, sl->comm, s->stream));
    if (above >= 0) SP_NCCL(s, g_nccl.Send(sl->d_cnt + 1, 1, ncclInt32, above, sl->comm, s->stream));
    if (above >= 0) SP_NCCL(s, g_nccl.Recv(scratch + 1, 1, ncclInt32, above, sl->comm, s->stream));
    if (below >= 0) SP_NCCL(s, g_nccl.Recv(scratch + 0, 1, ncclInt32, below, sl->comm, s->stream));
    SP_NCCL(s, g_nccl.GroupEnd());
    SP_CUDA(s, cudaMemcpyAsync(h, sl->d_cnt, 2 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    SP_CUDA(s, cudaMemcpyAsync(h + 2, scratch, 2 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    *n_dn = below >= 0 ? h[0] : 0;
    *n_up = above >= 0 ? h[1] : 0;
    *from_below = below >= 0 ? h[2] : 0;
    *from_above = above >= 0 ? h[3] : 0;
    return SP_OK;
}

static int slab_exchange_payload(sp_system* s, long long send_dn, long long send_up, long long recv_lo, long long recv_hi) {
    SlabState* sl = s->slab;
    int below, above;
    slab_peers(sl, &below, &above);
    // same call order as slab_exchange_counts; sizes in doubles, header included; a link without traffic is skipped on
    // both sides (size 0)
    SP_NCCL(s, g_nccl.GroupStart());
    if (below >= 0 && send_dn) SP_NCCL(s, g_nccl.Send(sl->sendbuf[0], (size_t)send_dn, ncclFloat64, below, sl->comm, s->stream));
    if (above >= 0 && send_up) SP_NCCL(s, g_nccl.Send(sl->sendbuf[1], (size_t)send_up, ncclFloat64, above, sl->comm, s->stream));
    if (above >= 0 && recv_hi) SP_NCCL(s, g_nccl.Recv(sl->recvbuf[1], (size_t)recv_hi, ncclFloat64, above, sl->comm, s->stream));
    if (below >= 0 && recv_lo) SP_NCCL(s, g_nccl.Recv(sl->recvbuf[0], (size_t)recv_lo, ncclFloat64, below, sl->comm, s->stream));
    SP_NCCL(s, g_nccl.GroupEnd());
    return SP_OK;
}

// message capacity for a link over which `count` particles went two rebuilds ago (both ends evaluate this)
// A boundary zone is two cell layers = about four lattice planes at h = 2 dr, and on an aligned lattice a whole plane
// can cross a cell boundary in one step (measured on the 10 M dam break: 216 808 -> 165 458 -> 216 808 particles in three
// consecutive rebuilds), so the head room is a factor of two over the larger of the last two known counts.
static long long slab_capacity(long long count) {
    long long c = 2 * count + 4096;
    return (c + 1023) / 1024 * 1024;
}

// (Re)create this rank's peer-to-peer block and map the neighbours' blocks.  Collective over the slab communicator
// (every rank calls it in the same rebuild: a bootstrap rebuild, whose counts size the buffers).  The handles travel over
// NCCL like the bootstrap counts; afterwards the message payload never touches NCCL on a link both ends mapped.
struct SlabP2PHello {
    cudaIpcMemHandle_t handle;
    long long cap, planes, ok, pid_rank;
};
static int slab_p2p_setup(sp_system* s, long long want_cap, long long planes) {
    SlabState* sl = s->slab;
    static const bool off = getenv("SP_SLAB_P2P") && atoi(getenv("SP_SLAB_P2P")) == 0;
    int below, above;
    slab_peers(sl, &below, &above);
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    // drop the old mappings and block
    for (int d = 0; d < 2; d++) {
        if (sl->p2p_mapped[d] && sl->p2p_peer[d] && !(d == 1 && sl->p2p_peer[1] == sl->p2p_peer[0])) cudaIpcCloseMemHandle(sl->p2p_peer[d]);
        sl->p2p_peer[d] = nullptr;
        sl->p2p_mapped[d] = false;
        sl->p2p_peer_cap[d] = sl->p2p_peer_planes[d] = 0;
        sl->p2p_send_seq[d] = sl->p2p_recv_seq[d] = 0;
    }
    if (sl->p2p_block) cudaFree(sl->p2p_block);
    sl->p2p_block = nullptr;
    cudaGetLastError();
    SlabP2PHello mine;
    memset(&mine, 0, sizeof mine);
    mine.cap = want_cap;
    mine.planes = planes;
    mine.pid_rank = sl->rank;
    mine.ok = 0;
    if (!off) {
        const size_t bytes = SLAB_P2P_FLAGS * sizeof(unsigned long long) + 2 * (size_t)(SLAB_HDR + want_cap * planes) * sizeof(double);
        if (cudaMalloc(&sl->p2p_block, bytes) == cudaSuccess && cudaMemset(sl->p2p_block, 0, bytes) == cudaSuccess &&
            cudaIpcGetMemHandle(&mine.handle, sl->p2p_block) == cudaSuccess)
            mine.ok = 1;
        else {
            cudaGetLastError();
            if (sl->p2p_block) cudaFree(sl->p2p_block);
            sl->p2p_block = nullptr;
        }
    }
    sl->p2p_cap = mine.ok ? want_cap : 0;
    sl->p2p_planes = mine.ok ? planes : 0;
    // exchange the hellos with both neighbours (same order rule as every exchange here)
    int rc = sp_ensure_stage(s, (long long)(3 * sizeof(SlabP2PHello) / sizeof(double)) + 8);
    if (rc) return rc;
    char* d_buf = reinterpret_cast<char*>(s->stage);
    SlabP2PHello theirs[2];
    memset(theirs, 0, sizeof theirs);
    SP_CUDA(s, cudaMemcpyAsync(d_buf, &mine, sizeof mine, cudaMemcpyHostToDevice, s->stream));
    SP_NCCL(s, g_nccl.GroupStart());
    if (below >= 0) SP_NCCL(s, g_nccl.Send(d_buf, sizeof mine, ncclInt8, below, sl->comm, s->stream));
    if (above >= 0) SP_NCCL(s, g_nccl.Send(d_buf, sizeof mine, ncclInt8, above, sl->comm, s->stream));
    if (above >= 0) SP_NCCL(s, g_nccl.Recv(d_buf + 2 * sizeof mine, sizeof mine, ncclInt8, above, sl->comm, s->stream));
    if (below >= 0) SP_NCCL(s, g_nccl.Recv(d_buf + sizeof mine, sizeof mine, ncclInt8, below, sl->comm, s->stream));
    SP_NCCL(s, g_nccl.GroupEnd());
    SP_CUDA(s, cudaMemcpyAsync(theirs, d_buf + sizeof mine, 2 * sizeof mine, cudaMemcpyDeviceToHost, s->stream));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    const int peer[2] = {below, above};
    for (int d = 0; d < 2; d++) {
        if (peer[d] < 0 || !mine.ok || !theirs[d].ok) continue;  // this link stays on NCCL
        if (peer[d] == sl->rank) {
            sl->p2p_peer[d] = sl->p2p_block;  // a periodic slab exchanging with itself
        } else if (d == 1 && peer[1] == peer[0] && sl->p2p_peer[0]) {
            sl->p2p_peer[1] = sl->p2p_peer[0];  // two ranks, periodic: the same neighbour on both sides
            sl->p2p_mapped[1] = sl->p2p_mapped[0];
        } else {
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, theirs[d].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                continue;  // (the other end cannot know: see the agreement round below)
            }
            sl->p2p_peer[d] = reinterpret_cast<unsigned long long*>(ptr);
            sl->p2p_mapped[d] = true;
        }
        sl->p2p_peer_cap[d] = theirs[d].cap;
        sl->p2p_peer_planes[d] = theirs[d].planes;
    }
    // agreement: a link is peer-to-peer only if BOTH ends mapped the other's block.  One more tiny exchange.
    long long mapped[2] = {sl->p2p_peer[0] ? 1 : 0, sl->p2p_peer[1] ? 1 : 0}, other[2] = {0, 0};
    SP_CUDA(s, cudaMemcpyAsync(d_buf, mapped, sizeof mapped, cudaMemcpyHostToDevice, s->stream));
    long long* d_l = reinterpret_cast<long long*>(d_buf);
    SP_NCCL(s, g_nccl.GroupStart());
    if (below >= 0) SP_NCCL(s, g_nccl.Send(d_l + 0, 1, ncclInt64, below, sl->comm, s->stream));
    if (above >= 0) SP_NCCL(s, g_nccl.Send(d_l + 1, 1, ncclInt64, above, sl->comm, s->stream));
    if (above >= 0) SP_NCCL(s, g_nccl.Recv(d_l + 3, 1, ncclInt64, above, sl->comm, s->stream));
    if (below >= 0) SP_NCCL(s, g_nccl.Recv(d_l + 2, 1, ncclInt64, below, sl->comm, s->stream));
    SP_NCCL(s, g_nccl.GroupEnd());
    SP_CUDA(s, cudaMemcpyAsync(other, d_l + 2, sizeof other, cudaMemcpyDeviceToHost, s->stream));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    for (int d = 0; d < 2; d++)
        if (sl->p2p_peer[d] && !other[d]) {
            if (sl->p2p_mapped[d] && !(d == 1 && sl->p2p_peer[1] == sl->p2p_peer[0])) cudaIpcCloseMemHandle(sl->p2p_peer[d]);
            sl->p2p_peer[d] = nullptr;
            sl->p2p_mapped[d] = false;
        }
    sl->p2p_state = 1;
    return SP_OK;
}

static int slab_rebuild(sp_system* s) {
    SlabState* sl = s->slab;
    const int B = 256;
    int rc;
    int below, above;
    slab_peers(sl, &below, &above);
    const bool trace = slab_trace_on();
    // SP_SLAB_TRACE=1: device time of the phases from events on the stream (one synchronisation at the end of the call)
    static thread_local cudaEvent_t tev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (trace) {
        for (auto& e : tev)
            if (!e) cudaEventCreate(&e);
        cudaEventRecord(tev[0], s->stream);
    }
    const bool steady = sl->history >= 2 && s->have_cells;
    long long cap_send[2] = {0, 0}, cap_recv[2] = {0, 0};
    if (steady) {
        // counts of rebuild b-2: long finished unless the host is more than two steps ahead of the device
        const int slot = (int)((sl->build_no - 2) % SLAB_RING);
        SP_CUDA(s, cudaEventSynchronize(sl->ring_ev[slot]));
        const int* h = sl->h_cnt + slot * 16;
        if (h[4]) {
            char buf[400];
            snprintf(buf, sizeof buf,
                     "slab exchange overflow at rebuild %lld (rank %d): a boundary message outgrew its capacity (sent %d down / %d "
                     "up with capacities %lld / %lld, i.e. the boundary population more than doubled within three rebuilds); particles "
                     "were lost",
                     sl->build_no - 2, sl->rank, h[0], h[1], (long long)h[10], (long long)h[11]);
            return sp_fail(s, SP_ERR_STATE, buf);
        }
        if (h[5]) return sp_fail(s, SP_ERR_STATE, "slab halo refresh: owner and ghost layers disagree");
        if (h[6]) return sp_fail(s, SP_ERR_STATE, "slab exchange: a peer-to-peer message or acknowledgement did not arrive within "
                                                  "the time limit (a neighbouring rank stopped, or the ranks do not issue the same calls)");
        if (h[12]) {
            char buf[300];
            snprintf(buf, sizeof buf,
                     "slab exchange (rank %d): %d particle(s) moved more than the two ghost layers between two rebuilds and "
                     "left this rank's window inside the global box; they were culled (time step too large?)", sl->rank, h[12]);
            return sp_fail(s, SP_ERR_STATE, buf);
        }
        if (trace)
            fprintf(stderr, "[slab trace rank %d] rebuild %lld: sent %d dn / %d up, received %d lo / %d hi, owned %d, alive %d\n",
                    sl->rank, sl->build_no - 2, h[0], h[1], h[2], h[3], h[8], h[9]);
        // both ends of a link hold the counts that went over it in rebuilds b-2 and b-3
        int big[4] = {h[0], h[1], h[2], h[3]};
        if (sl->history >= 3) {
            const int* h3 = sl->h_cnt + (int)((sl->build_no - 3) % SLAB_RING) * 16;
            for (int i = 0; i < 4; i++) big[i] = std::max(big[i], h3[i]);
        }
        cap_send[0] = below >= 0 ? slab_capacity(big[0]) : 0;
        cap_send[1] = above >= 0 ? slab_capacity(big[1]) : 0;
        cap_recv[0] = below >= 0 ? slab_capacity(big[2]) : 0;
        cap_recv[1] = above >= 0 ? slab_capacity(big[3]) : 0;
        // slot bound before this rebuild: alive after rebuild b-2 plus everything rebuild b-1 may have appended
        s->n = (long long)h[9] + sl->cap_recv[0] + sl->cap_recv[1];
        s->n_exact = false;
        s->count_pending = false;
    } else {
        if ((rc = sp_settle(s))) return rc;
    }
    const SlabWin w = slab_win(s, !steady);
    double* X = s->fields[0].d;
    double* ghost = s->fields[sl->f_ghost].d;
    SP_LAUNCH(s, k_slab_clear, 1, 32, 0, sl->d_cnt);
    if (s->n > 0) {
        if (sl->fresh_particles) {
            SP_LAUNCH(s, k_slab_assign_gid, sp_blocks(s->n, B), B, 0, s->fields[sl->f_gid].d, s->counters,
                      (double)sl->rank * 17592186044416.0 + (double)sl->gid_base);
            sl->gid_base += s->n;
            sl->fresh_particles = false;
        }
        // 1: old ghosts die
        SP_LAUNCH(s, k_slab_kill_ghosts, 1184, B, 0, w, X, ghost);
        // 2: selection: counts per block, one scan
        SP_LAUNCH(s, k_slab_sel_count, dim3(SLAB_NB, 2), B, 0, s->g, w, X, s->cap, ghost, sl->d_cnt + 16);
        // (no capacity to check in the bootstrap rebuilds — they size the messages from the counts — nor towards a side
        // without a neighbour, whose selection is never sent)
        SP_LAUNCH(s, k_slab_sel_scan, 1, SLAB_NB, 0, sl->d_cnt + 16, sl->d_cnt, steady && below >= 0 ? cap_send[0] : -1LL,
                  steady && above >= 0 ? cap_send[1] : -1LL);
    }
    if (!steady) {
        long long n_dn = 0, n_up = 0, fb = 0, fa = 0;
        if ((rc = slab_exchange_counts(s, &n_dn, &n_up, &fb, &fa))) return rc;
        cap_send[0] = below >= 0 ? std::max<long long>(n_dn, 1) : 0;
        cap_send[1] = above >= 0 ? std::max<long long>(n_up, 1) : 0;
        cap_recv[0] = below >= 0 ? std::max<long long>(fb, 1) : 0;
        cap_recv[1] = above >= 0 ? std::max<long long>(fa, 1) : 0;
    }
    std::vector<SlabPlanes> tabs;
    int nplanes = 0;
    slab_planes(s, tabs, &nplanes);
    // peer-to-peer blocks: (re)created in the first bootstrap rebuild after the host touched the particles, sized from
    // its counts (4x head room) — every rank is in that rebuild at the same time, so the handshake is collective
    if (!steady && sl->history == 0) {
        const long long most = std::max(std::max(cap_send[0], cap_send[1]), std::max(cap_recv[0], cap_recv[1]));
        if ((rc = slab_p2p_setup(s, std::max<long long>(4 * most, 65536), nplanes + 4))) return rc;
    }
    // which links carry this rebuild's messages peer to peer: both ends evaluate the same numbers
    bool p2p_send[2], p2p_recv[2];
    for (int d = 0; d < 2; d++) {
        p2p_send[d] = cap_send[d] && sl->p2p_peer[d] && cap_send[d] <= sl->p2p_peer_cap[d] && nplanes <= sl->p2p_peer_planes[d];
        p2p_recv[d] = cap_recv[d] && sl->p2p_peer[d] && cap_recv[d] <= sl->p2p_cap && nplanes <= sl->p2p_planes;
    }
    long long nccl_most = 0;
    for (int d = 0; d < 2; d++) {
        if (!p2p_send[d]) nccl_most = std::max(nccl_most, cap_send[d]);
        if (!p2p_recv[d]) nccl_most = std::max(nccl_most, cap_recv[d]);
    }
    if (nccl_most && (rc = slab_ensure_buffers(s, SLAB_HDR + nccl_most * nplanes + 16))) return rc;
    // room for the arrivals (growing reallocates every plane: rare, and it waits for the stream)
    const long long n_new = s->n + cap_recv[0] + cap_recv[1];
    if (n_new > s->cap) {
        if ((rc = sp_ensure_capacity(s, n_new + n_new / 8))) return rc;
        slab_planes(s, tabs, &nplanes);
        X = s->fields[0].d;
        ghost = s->fields[sl->f_ghost].d;
    }
    // a rank's block: [flags][buffer "from below"][buffer "from above"]
    auto p2p_buffer = [](unsigned long long* block, long long cap, long long planes, int which) -> double* {
        return reinterpret_cast<double*>(block + SLAB_P2P_FLAGS) + (size_t)which * (size_t)(SLAB_HDR + cap * planes);
    };
    // what I send DOWN lands in the lower neighbour's "from above" buffer (1), what I send UP in the upper one's (0)
    double* dst[2] = {nullptr, nullptr};
    long long dst_stride[2] = {0, 0};
    for (int d = 0; d < 2; d++) {
        if (!cap_send[d]) continue;
        if (p2p_send[d]) {
            dst[d] = p2p_buffer(sl->p2p_peer[d], sl->p2p_peer_cap[d], sl->p2p_peer_planes[d], d == 0 ? 1 : 0);
            dst_stride[d] = sl->p2p_peer_cap[d];
        } else {
            dst[d] = sl->sendbuf[d];
            dst_stride[d] = cap_send[d];
        }
    }
    // the neighbour must have consumed my previous message before I overwrite it: it writes my flag 2 (below) / 3 (above)
    if (p2p_send[0] || p2p_send[1])
        SP_LAUNCH(s, k_slab_wait, 1, 32, 0, p2p_send[0] ? sl->p2p_block + 2 : (unsigned long long*)nullptr,
                  (unsigned long long)sl->p2p_send_seq[0], p2p_send[1] ? sl->p2p_block + 3 : (unsigned long long*)nullptr,
                  (unsigned long long)sl->p2p_send_seq[1], sl->d_cnt, slab_wait_limit(s));
    // 3: ordered pack (all planes), headers
    if (s->n > 0 && (cap_send[0] || cap_send[1])) {
        int plane0 = 0;
        for (SlabPlanes& t : tabs) {
            SP_LAUNCH(s, k_slab_sel_pack, dim3(SLAB_NB, 2), B, 0, s->g, w, X, s->cap, ghost, sl->d_cnt + 16, sl->d_cnt, t, plane0,
                      dst[0], dst_stride[0], dst[1], dst_stride[1]);
            plane0 += t.count;
        }
    } else {
        // nothing selected on an empty system: the headers still have to say so
        for (int d = 0; d < 2; d++)
            if (dst[d]) SP_CUDA(s, cudaMemsetAsync(dst[d], 0, SLAB_HDR * sizeof(double), s->stream));
    }
    for (int d = 0; d < 2; d++)
        if (p2p_send[d]) sl->p2p_send_seq[d]++;
    if (p2p_send[0] || p2p_send[1])  // "message k is in your buffer": flag 1 of the rank below (from above), flag 0 of the rank above
        SP_LAUNCH(s, k_slab_signal, 1, 32, 0, p2p_send[0] ? sl->p2p_peer[0] + 1 : (unsigned long long*)nullptr,
                  (unsigned long long)sl->p2p_send_seq[0], p2p_send[1] ? sl->p2p_peer[1] + 0 : (unsigned long long*)nullptr,
                  (unsigned long long)sl->p2p_send_seq[1]);
    if (trace) cudaEventRecord(tev[1], s->stream);
    // 4: one exchange: NCCL for the links that are not peer to peer, flag waits for those that are
    if ((rc = slab_exchange_payload(s, cap_send[0] && !p2p_send[0] ? SLAB_HDR + cap_send[0] * nplanes : 0,
                                    cap_send[1] && !p2p_send[1] ? SLAB_HDR + cap_send[1] * nplanes : 0,
                                    cap_recv[0] && !p2p_recv[0] ? SLAB_HDR + cap_recv[0] * nplanes : 0,
                                    cap_recv[1] && !p2p_recv[1] ? SLAB_HDR + cap_recv[1] * nplanes : 0)))
        return rc;
    const double* src[2] = {nullptr, nullptr};
    long long src_stride[2] = {0, 0};
    for (int d = 0; d < 2; d++) {
        if (!cap_recv[d]) continue;
        if (p2p_recv[d]) {
            sl->p2p_recv_seq[d]++;
            src[d] = p2p_buffer(sl->p2p_block, sl->p2p_cap, sl->p2p_planes, d);
            src_stride[d] = sl->p2p_cap;
        } else {
            src[d] = sl->recvbuf[d];
            src_stride[d] = cap_recv[d];
        }
    }
    if (p2p_recv[0] || p2p_recv[1])
        SP_LAUNCH(s, k_slab_wait, 1, 32, 0, p2p_recv[0] ? sl->p2p_block + 0 : (unsigned long long*)nullptr,
                  (unsigned long long)sl->p2p_recv_seq[0], p2p_recv[1] ? sl->p2p_block + 1 : (unsigned long long*)nullptr,
                  (unsigned long long)sl->p2p_recv_seq[1], sl->d_cnt, slab_wait_limit(s));
    if (trace) cudaEventRecord(tev[2], s->stream);
    // 5: arrivals behind the alive slots; a particle that crossed the periodic boundary is shifted by one period
    const double shift_lo = (sl->periodic && sl->rank == 0) ? -sl->period : 0.0;              // came from the top rank
    const double shift_hi = (sl->periodic && sl->rank == sl->nranks - 1) ? sl->period : 0.0;  // came from rank 0
    if (cap_recv[0] + cap_recv[1]) {
        SP_LAUNCH(s, k_slab_arrival_layout, 1, 32, 0, src[0], cap_recv[0], p2p_recv[0] ? 0 : 1, src[1], cap_recv[1], p2p_recv[1] ? 0 : 1,
                  sl->d_cnt);
        int plane0 = 0;
        for (SlabPlanes& t : tabs) {
            if (cap_recv[0])
                SP_LAUNCH(s, k_slab_unpack, sp_blocks(cap_recv[0], B), B, 0, t, plane0, src[0], src_stride[0], cap_recv[0],
                          p2p_recv[0] ? 0 : 1, shift_lo, s->counters, s->ref, sl->d_cnt, 0);
            if (cap_recv[1])
                SP_LAUNCH(s, k_slab_unpack, sp_blocks(cap_recv[1], B), B, 0, t, plane0, src[1], src_stride[1], cap_recv[1],
                          p2p_recv[1] ? 0 : 1, shift_hi, s->counters, s->ref, sl->d_cnt, 1);
            plane0 += t.count;
        }
        // fields that did not travel (zero everywhere): the arrival slots must hold zero too
        std::vector<SlabPlanes> ztabs;
        int nz = 0;
        slab_planes(s, ztabs, &nz, false, true);
        for (SlabPlanes& t : ztabs)
            SP_LAUNCH(s, k_slab_zero_tail, sp_blocks(cap_recv[0] + cap_recv[1], B), B, 0, t, cap_recv[0] + cap_recv[1], sl->d_cnt,
                      s->counters);
        SP_LAUNCH(s, k_slab_add_alive, 1, 32, 0, s->counters, sl->d_cnt);
        // tell the senders their buffers are free again: the message from below was the lower neighbour's UP message
        // (its flag 3), the one from above the upper neighbour's DOWN message (its flag 2)
        if (p2p_recv[0] || p2p_recv[1])
            SP_LAUNCH(s, k_slab_signal, 1, 32, 0, p2p_recv[0] ? sl->p2p_peer[0] + 3 : (unsigned long long*)nullptr,
                      (unsigned long long)sl->p2p_recv_seq[0], p2p_recv[1] ? sl->p2p_peer[1] + 2 : (unsigned long long*)nullptr,
                      (unsigned long long)sl->p2p_recv_seq[1]);
    }
    s->n = n_new;
    s->n_exact = false;
    s->count_pending = false;
    s->x_version++;
    for (int d = 0; d < 2; d++) {
        sl->cap_send[d] = cap_send[d];
        sl->cap_recv[d] = cap_recv[d];
    }
    if (trace) cudaEventRecord(tev[3], s->stream);
    // 6: the ordinary build on the local window (in-cell order by descending "_gid"), then ghost flags
    if ((rc = sp_build_cells(s))) return rc;
    if (s->n)
        SP_LAUNCH(s, k_slab_post, sp_blocks(s->n, B), B, 0, s->key, s->fields[sl->f_ghost].d, s->ref, w.L, w.nl, s->counters,
                  sl->d_cnt);
    s->identity_order = true;
    // 7: this rebuild's counts go to the ring: [0..8] slab counts, [9] alive
    {
        const int slot = (int)(sl->build_no % SLAB_RING);
        int* h = sl->h_cnt + slot * 16;
        SP_CUDA(s, cudaMemcpyAsync(h, sl->d_cnt, 9 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        SP_CUDA(s, cudaMemcpyAsync(h + 9, s->counters + SP_CNT_ALIVE, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        SP_CUDA(s, cudaMemcpyAsync(h + 12, s->counters + SP_CNT_LOST, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        h[10] = (int)cap_send[0];  // (host-side notes for the error message; written before the event can complete)
        h[11] = (int)cap_send[1];
        SP_CUDA(s, cudaEventRecord(sl->ring_ev[slot], s->stream));
    }
    sl->build_no++;
    sl->history++;
    if (trace) {
        cudaEventRecord(tev[4], s->stream);
        cudaEventSynchronize(tev[4]);
        float ms[4] = {0, 0, 0, 0};
        for (int k = 0; k < 4; k++) cudaEventElapsedTime(&ms[k], tev[k], tev[k + 1]);
        for (int k = 0; k < 4; k++) sl->trace_s[k] += ms[k];
        if (++sl->trace_calls % 10 == 0) {
            fprintf(stderr, "[slab trace rank %d] calls=%lld (last 10) select+pack=%.3f ms exchange=%.3f ms unpack=%.3f ms build+post=%.3f ms "
                            "(bound n=%lld, caps %lld %lld | %lld %lld)\n",
                    sl->rank, sl->trace_calls, sl->trace_s[0] / 10, sl->trace_s[1] / 10, sl->trace_s[2] / 10, sl->trace_s[3] / 10,
                    (long long)s->n, cap_send[0], cap_send[1], cap_recv[0], cap_recv[1]);
            for (int k = 0; k < 4; k++) sl->trace_s[k] = 0;
        }
    }
    return SP_OK;
}

// ------------------------------------------------------------------ C ABI
extern "C" {

int32_t sp_slab_unique_id(uint8_t id[128]) {
    if (!id) return SP_ERR_INVALID;
    if (!nccl_load()) return sp_fail(nullptr, SP_ERR_NCCL, g_nccl.err);
    ncclUniqueId u;
    ncclResult_t r = g_nccl.GetUniqueId(&u);
    if (r != ncclSuccess) return sp_fail(nullptr, SP_ERR_NCCL, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r));
    static_assert(sizeof(u) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id, &u, 128);
    return SP_OK;
}

int32_t sp_slab_init(sp_system* s, const uint8_t id[128], int32_t rank, int32_t nranks, int32_t periodic) {
    return sp_slab_init_cuts(s, id, rank, nranks, periodic, nullptr);
}

int32_t sp_slab_init_cuts(sp_system* s, const uint8_t id[128], int32_t rank, int32_t nranks, int32_t periodic,
                          const int64_t* cuts) {
    if (!s || !id) return SP_ERR_INVALID;
    if (nranks < 1 || rank < 0 || rank >= nranks) return sp_fail(s, SP_ERR_INVALID, "bad rank / nranks");
    if (s->slab) return sp_fail(s, SP_ERR_STATE, "slab already initialised");
    if (s->n != 0) return sp_fail(s, SP_ERR_STATE, "sp_slab_init must be called before particles are added");
    SP_CUDA(s, cudaSetDevice(s->device));
    if (!nccl_load()) return sp_fail(s, SP_ERR_NCCL, g_nccl.err);
    SpGrid& g = s->g;
    const int axis = g.dim == 2 ? 1 : 2;  // slowest key axis
    // owned layers [c0, c1): equal layer counts, or the caller's cut planes (count-balanced slabs: SURVEY 8(e))
    long long c0, c1;
    if (cuts) {
        if (cuts[0] != 0 || cuts[nranks] != g.lim[axis]) return sp_fail(s, SP_ERR_INVALID, "cuts must start at 0 and end at key_lim");
        for (int r = 0; r < nranks; r++)
            if (cuts[r + 1] <= cuts[r]) return sp_fail(s, SP_ERR_INVALID, "cuts must be increasing");
        c0 = cuts[rank];
        c1 = cuts[rank + 1];
    } else {
        const long long base = g.lim[axis] / nranks, rem = g.lim[axis] % nranks;
        c0 = rank * base + std::min<long long>(rank, rem);
        c1 = c0 + base + (rank < rem ? 1 : 0);
    }
    // three owned layers per rank: the two boundary layers facing one neighbour must not receive migrants from the other
    // (that is what makes a rank's boundary layers and its neighbour's ghost layers the same particle sets)
    if (nranks > 1 || periodic) {
        bool ok = true;
        if (cuts)
            for (int r = 0; r < nranks; r++) ok = ok && cuts[r + 1] - cuts[r] >= 3;
        else
            ok = g.lim[axis] >= 3LL * nranks;
        if (!ok) return sp_fail(s, SP_ERR_INVALID, "fewer than three cell layers per rank along the slab axis");
    }
    SlabState* sl = new SlabState();
    sl->rank = rank;
    sl->nranks = nranks;
    sl->periodic = periodic ? 1 : 0;
    sl->axis = axis;
    sl->gphase = g.phase[axis];
    sl->glim = g.lim[axis];
    sl->c0 = c0;
    sl->c1 = c1;
    sl->period = (double)sl->glim * g.h;
    s->slab = sl;
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclResult_t r = g_nccl.CommInitRank(&sl->comm, nranks, u, rank);
    if (r != ncclSuccess) {
        s->slab = nullptr;
        delete sl;
        return sp_fail(s, SP_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
    }
    SP_CUDA(s, sp_dmalloc(&sl->d_cnt, (16 + 2 * SLAB_NB) * sizeof(int)));
    SP_CUDA(s, cudaMemset(sl->d_cnt, 0, (16 + 2 * SLAB_NB) * sizeof(int)));
    SP_CUDA(s, cudaHostAlloc(&sl->h_cnt, (SLAB_RING * 16 + 16) * sizeof(int), cudaHostAllocDefault));
    memset(sl->h_cnt, 0, (SLAB_RING * 16 + 16) * sizeof(int));
    for (int k = 0; k < SLAB_RING; k++) SP_CUDA(s, cudaEventCreateWithFlags(&sl->ring_ev[k], cudaEventDisableTiming));
    // local cell window: owned layers [c0, c1) plus SLAB_W ghost layers per side
    g.phase[axis] = sl->gphase + sl->c0 - SLAB_W;
    g.lim[axis] = (sl->c1 - sl->c0) + 2 * SLAB_W;
    g.key_max = g.lim[0] * g.lim[1] * g.lim[2];
    // the local window (with its ghost layers) can be larger than the global grid when nranks is small
    SP_CUDA(s, sp_dfree(s, s->cell_start));
    SP_CUDA(s, sp_dfree(s, s->cell_fill));
    s->cell_start = s->cell_fill = nullptr;
    SP_CUDA(s, sp_dmalloc(&s->cell_start, (size_t)(g.key_max + 3) * sizeof(int)));
    SP_CUDA(s, sp_dmalloc(&s->cell_fill, (size_t)(g.key_max + 3) * sizeof(int)));
    SP_CUDA(s, cudaMemset(s->cell_start, 0, (size_t)(g.key_max + 3) * sizeof(int)));
    if (s->scan_tmp_len < (g.key_max + 3) / 1024 + 1024) {
        SP_CUDA(s, sp_dfree(s, s->scan_tmp));
        s->scan_tmp = nullptr;
        s->scan_tmp_len = (g.key_max + 3) / 1024 + 2048;
        SP_CUDA(s, sp_dmalloc(&s->scan_tmp, (size_t)s->scan_tmp_len * sizeof(int)));
    }
    g.slab_axis = axis;
    g.slab_periodic = sl->periodic;
    if (g.dim == 3) {
        int k = 0;
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++)
                for (int dk = -1; dk <= 1; dk++) s->key_diff[k++] = di + g.lim[0] * (dj + g.lim[1] * dk);
    }
    int32_t fid;
    int rc;
    if ((rc = sp_add_field(s, "_ghost", 1, &fid))) return rc;
    sl->f_ghost = fid;
    if ((rc = sp_add_field(s, "_gid", 1, &fid))) return rc;
    sl->f_gid = fid;
    return SP_OK;
}

int32_t sp_slab_range(sp_system* s, int64_t* cell_lo, int64_t* cell_hi, double* coord_lo, double* coord_hi, int32_t* axis) {
    if (!s || !s->slab) return sp_fail(s, SP_ERR_STATE, "not a slab system");
    SlabState* sl = s->slab;
    if (cell_lo) *cell_lo = sl->c0;
    if (cell_hi) *cell_hi = sl->c1;
    // a particle belongs to this rank iff floor(x_axis/h) - key_phase_global lies in [cell_lo, cell_hi)
    if (coord_lo) *coord_lo = (double)(sl->gphase + sl->c0) * s->g.h;
    if (coord_hi) *coord_hi = (double)(sl->gphase + sl->c1) * s->g.h;
    if (axis) *axis = sl->axis;
    return SP_OK;
}

int32_t sp_slab_create_cell_list(sp_system* s) {
    if (!s) return SP_ERR_INVALID;
    if (!s->slab) return sp_fail(s, SP_ERR_STATE, "not a slab system");
    SP_CUDA(s, cudaSetDevice(s->device));
    int rc = sp_time_begin(s);
    if (rc) return rc;
    if ((rc = slab_rebuild(s))) return rc;
    return sp_time_end(s);
}

int32_t sp_slab_halo_refresh(sp_system* s, const int32_t* fields, int32_t nfields) {
    if (!s || !fields) return SP_ERR_INVALID;
    if (!s->slab) return sp_fail(s, SP_ERR_STATE, "not a slab system");
    if (!s->have_cells) return sp_fail(s, SP_ERR_STATE, "halo refresh before sp_slab_create_cell_list");
    SP_CUDA(s, cudaSetDevice(s->device));
    SlabState* sl = s->slab;
    int rc = sp_time_begin(s);
    if (rc) return rc;
    int ncomp_total = 0;
    for (int k = 0; k < nfields; k++) {
        if (fields[k] < 0 || fields[k] >= (int)s->fields.size()) return sp_fail(s, SP_ERR_INVALID, "bad field id");
        ncomp_total += s->fields[fields[k]].ncomp;
    }
    int below, above;
    slab_peers(sl, &below, &above);
    // the boundary layers hold at most what the last rebuild sent (that message also carried the migrants), the ghost
    // layers at most what it received: the capacities of that rebuild fit, and both ends know them
    const long long cs0 = below >= 0 ? sl->cap_send[0] : 0, cs1 = above >= 0 ? sl->cap_send[1] : 0;
    const long long cr0 = below >= 0 ? sl->cap_recv[0] : 0, cr1 = above >= 0 ? sl->cap_recv[1] : 0;
    const long long biggest = std::max(std::max(cs0, cs1), std::max(cr0, cr1));
    if (biggest == 0) return sp_time_end(s);
    if ((rc = slab_ensure_buffers(s, SLAB_HDR + biggest * ncomp_total + 16))) return rc;
    const SlabWin w = slab_win(s, false);
    const int B = 256;
    int c0 = 0;
    for (int k = 0; k < nfields; k++) {
        SpField& f = s->fields[fields[k]];
        SP_LAUNCH(s, k_slab_refresh_pack, dim3(296, 2), B, 0, w, f.d, s->cap, f.ncomp, c0, cs0 ? sl->sendbuf[0] : (double*)nullptr,
                  cs0, cs1 ? sl->sendbuf[1] : (double*)nullptr, cs1);
        c0 += f.ncomp;
    }
    if ((rc = slab_exchange_payload(s, cs0 ? SLAB_HDR + cs0 * ncomp_total : 0, cs1 ? SLAB_HDR + cs1 * ncomp_total : 0,
                                    cr0 ? SLAB_HDR + cr0 * ncomp_total : 0, cr1 ? SLAB_HDR + cr1 * ncomp_total : 0)))
        return rc;
    const double shift_lo = (sl->periodic && sl->rank == 0) ? -sl->period : 0.0;
    const double shift_hi = (sl->periodic && sl->rank == sl->nranks - 1) ? sl->period : 0.0;
    c0 = 0;
    for (int k = 0; k < nfields; k++) {
        SpField& f = s->fields[fields[k]];
        const int axis_comp = fields[k] == 0 ? sl->axis : -1;
        f.version++;  // ghost values change (a field that is zero everywhere stays zero: known_zero is kept)
        if (fields[k] == 0) s->x_version++;
        SP_LAUNCH(s, k_slab_refresh_unpack, dim3(296, 2), B, 0, w, f.d, s->cap, f.ncomp, c0,
                  cr0 ? sl->recvbuf[0] : (const double*)nullptr, cr0, cr1 ? sl->recvbuf[1] : (const double*)nullptr, cr1, axis_comp,
                  shift_lo, shift_hi, sl->d_cnt);
        c0 += f.ncomp;
    }
    return sp_time_end(s);
}

int32_t sp_slab_num_owned(sp_system* s, int64_t* n_owned) {
    if (!s || !n_owned) return SP_ERR_INVALID;
    if (!s->slab) return sp_fail(s, SP_ERR_STATE, "not a slab system");
    SP_CUDA(s, cudaSetDevice(s->device));
    SlabState* sl = s->slab;
    int* h = sl->h_cnt + SLAB_RING * 16 + 8;
    SP_CUDA(s, cudaMemcpyAsync(h, sl->d_cnt + 8, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    *n_owned = sl->build_no ? *h : s->n;  // before the first rebuild everything the host added is owned
    return SP_OK;
}

int32_t sp_slab_allreduce(sp_system* s, double* inout, int32_t count, int32_t is_max) {
    if (!s || !inout || count < 0 || count > 1024) return SP_ERR_INVALID;
    if (!s->slab) return sp_fail(s, SP_ERR_STATE, "not a slab system");
    SP_CUDA(s, cudaSetDevice(s->device));
    int rc = sp_ensure_stage(s, count + 8);
    if (rc) return rc;
    SP_CUDA(s, cudaMemcpyAsync(s->stage, inout, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    if ((rc = sp_slab_allreduce_device(s, s->stage, count, is_max))) return rc;
    SP_CUDA(s, cudaMemcpyAsync(inout, s->stage, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    return SP_OK;
}

}  // extern "C"
