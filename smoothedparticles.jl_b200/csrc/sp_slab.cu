// sp_slab.cu — slab decomposition of one ParticleSystem over the GPUs of a node (one process per GPU).
//
// The reference is single-process shared memory; this is the multi-GPU extension SURVEY §8(e) specifies.
// The global cell grid of sp_create is cut along the SLOWEST key axis (z in 3-D, y in 2-D) into `nranks`
// slabs of whole cell layers.  Because that axis is the slowest in the linear key, after the local sort a
// rank's two boundary cell layers and its two ghost layers are contiguous slot ranges.
//
// sp_slab_create_cell_list (replaces create_cell_list! on a slab system):
//   1. drop last step's ghosts; classify owned particles by the cell layer of their CURRENT position
//   2. MIGRATION: particles whose layer left [c0, c1) are packed (every field) and ncclSend/ncclRecv'd to
//      the lower / upper neighbour (wrapping, with the coordinate shifted by the period, if periodic)
//   3. GHOST HALO: copies of the owned particles in layers c0 and c1-1 go to the neighbours, which append
//      them as ghosts (flag field "_ghost" = 1 from below / 2 from above, "_hidx" = index in the message)
//   4. the ordinary cell-list build over owned + ghost particles on the LOCAL cell window
// sp_slab_halo_refresh re-sends chosen fields of the same boundary particles in message order (the senders
// remember the message index in "_sdn"/"_sup"), e.g. rho and P after find_pressure!.
// Reductions and CG dot products skip ghosts and are summed with ncclAllReduce.
//
// NCCL is loaded lazily with dlopen, so the single-GPU library has no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstring>

#include <time.h>
#include <cstdio>

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

struct SlabState {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, periodic = 0, axis = 2;
    long long gphase = 0, glim = 1;  // global key_phase / key_lim along the axis
    long long c0 = 0, c1 = 1;        // owned global cell layers [c0, c1), 0-based from gphase
    double period = 0.0;
    int f_ghost = -1, f_hidx = -1, f_sdn = -1, f_sup = -1;
    double* sendbuf[2] = {nullptr, nullptr};
    double* recvbuf[2] = {nullptr, nullptr};
    long long buf_len = 0;  // doubles per buffer
    int* d_cnt = nullptr;   // [0],[1] send counts down/up, [2],[3] received counts from below/above
    int* h_cnt = nullptr;
    long long n_send[2] = {0, 0};   // ghost message sizes sent down / up at the last rebuild
    long long n_ghost[2] = {0, 0};  // ghosts received from below / above
    long long n_owned = 0;
    // slot window of the selection passes: after a rebuild the slots are sorted by cell layer, and a particle moves
    // less than one cell per step, so only slots [0, sel_a) and [sel_b, n) (three cell layers per side) can hold
    // old ghosts, migrants or new boundary particles.  sel_valid is dropped whenever the host touched positions.
    long long sel_a = 0, sel_b = 0;
    bool sel_valid = false;
    // SP_SLAB_TRACE=1: host wall-clock of the phases of sp_slab_create_cell_list (each ends in a stream sync)
    double trace_s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long trace_calls = 0;
};

static bool slab_trace_on() {
    static const bool on = getenv("SP_SLAB_TRACE") && atoi(getenv("SP_SLAB_TRACE"));
    return on;
}
static double slab_now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
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
    if (sl->h_cnt) cudaFreeHost(sl->h_cnt);
    delete sl;
    s->slab = nullptr;
}

void sp_slab_host_touched(sp_system* s) {
    if (s->slab) s->slab->sel_valid = false;
}

const double* sp_slab_ghost_mask(sp_system* s) {
    if (!s->slab) return nullptr;
    return s->fields[s->slab->f_ghost].d;
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
};

// selection passes run over t in [0, n_sel): slot = t below sel_a, sel_b + (t - sel_a) above
struct SlabSel {
    long long a, b, n_sel;
    __host__ __device__ long long slot(long long t) const { return t < a ? t : b + (t - a); }
};

// flags: dn[t] = 1 if the particle migrates to the lower neighbour, up[t] likewise; old ghosts are killed (x = NaN)
__global__ void k_slab_classify(SpGrid g, int rank, int nranks, long long gphase, long long c0, long long c1,
                                double* x, long long cap, const double* ghost, SlabSel sel, int* dn, int* up) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= sel.n_sel) return;
    const long long s = sel.slot(t);
    int fd = 0, fu = 0;
    if (ghost[s] != 0.0) {
        x[s] = nan("");  // dropped by the build
    } else {
        const double xa = x[(size_t)g.slab_axis * cap + s];
        const double q = floor(__ddiv_rn(xa, g.h));
        if (q == q && fabs(q) < 9.0e18) {
            const long long ca = (long long)q - gphase;
            if (ca < c0) fd = (g.slab_periodic || rank > 0) ? 1 : 0;          // else: left the global domain, culled
            else if (ca >= c1) fu = (g.slab_periodic || rank < nranks - 1) ? 1 : 0;
        }
    }
    dn[t] = fd;
    up[t] = fu;
}
// owned, alive particles in the first / last owned layer are sent as ghosts down / up
__global__ void k_slab_boundary(SpGrid g, int rank, int nranks, long long gphase, long long c0, long long c1,
                                const double* x, long long cap, const double* ghost, SlabSel sel, int* dn, int* up) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= sel.n_sel) return;
    const long long s = sel.slot(t);
    int fd = 0, fu = 0;
    const double x0 = x[s];
    if (ghost[s] == 0.0 && x0 == x0) {
        const double q = floor(__ddiv_rn(x[(size_t)g.slab_axis * cap + s], g.h));
        if (q == q && fabs(q) < 9.0e18) {
            const long long ca = (long long)q - gphase;
            if (ca == c0) fd = (g.slab_periodic || rank > 0) ? 1 : 0;
            if (ca == c1 - 1) fu = (g.slab_periodic || rank < nranks - 1) ? 1 : 0;
        }
    }
    dn[t] = fd;
    up[t] = fu;
}
// message index of every selected slot (exclusive scans in posd/posu), -1 otherwise, as Float64 fields
__global__ void k_slab_fill2(double* a, double* b, double v, long long n) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s < n) {
        a[s] = v;
        b[s] = v;
    }
}
__global__ void k_slab_record(const int* fd, const int* fu, const int* posd, const int* posu, double* sdn, double* sup,
                              SlabSel sel) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= sel.n_sel) return;
    const long long s = sel.slot(t);
    sdn[s] = fd[t] ? (double)posd[t] : -1.0;
    sup[s] = fu[t] ? (double)posu[t] : -1.0;
}
__global__ void k_slab_pack(SlabPlanes tab, const int* flag, const int* pos, SlabSel sel, double* buf, long long count) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= sel.n_sel || !flag[t]) return;
    const long long s = sel.slot(t);
    const long long m = pos[t];
    for (int c = 0; c < tab.count; c++) buf[(size_t)c * count + m] = tab.p[c][s];
}
__global__ void k_slab_kill(const int* fd, const int* fu, double* x, SlabSel sel) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < sel.n_sel && (fd[t] | fu[t])) x[sel.slot(t)] = nan("");
}
__global__ void k_slab_unpack(SlabPlanes tab, const double* buf, long long count, long long base, double shift, int* ref) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= count) return;
    for (int c = 0; c < tab.count; c++) {
        double v = buf[(size_t)c * count + t];
        if (c == tab.axis_plane) v += shift;
        tab.p[c][base + t] = v;
    }
    if (ref) ref[base + t] = (int)(base + t);
}
__global__ void k_slab_mark(double* ghost, double* hidx, double* sdn, double* sup, long long base, long long count,
                            double side) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= count) return;
    ghost[base + t] = side;
    hidx[base + t] = (double)t;
    sdn[base + t] = -1.0;
    sup[base + t] = -1.0;
}
// halo refresh: boundary owners write the field into message order; ghosts read it back by message index
__global__ void k_slab_refresh_pack(const double* f, long long cap, int ncomp, const double* sel, long long n, double* buf,
                                    long long count) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= n) return;
    const double v = sel[s];
    if (!(v >= 0.0)) return;
    const long long t = (long long)v;
    for (int c = 0; c < ncomp; c++) buf[(size_t)c * count + t] = f[(size_t)c * cap + s];
}
__global__ void k_slab_refresh_unpack(double* f, long long cap, int ncomp, const double* ghost, const double* hidx,
                                      long long n, const double* buf_lo, long long cnt_lo, const double* buf_hi,
                                      long long cnt_hi, int axis_comp, double shift_lo, double shift_hi) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= n) return;
    const double gflag = ghost[s];
    if (gflag == 0.0) return;
    const long long t = (long long)hidx[s];
    const bool lo = gflag == 1.0;
    const double* buf = lo ? buf_lo : buf_hi;
    const long long cnt = lo ? cnt_lo : cnt_hi;
    for (int c = 0; c < ncomp; c++) {
        double v = buf[(size_t)c * cnt + t];
        if (c == axis_comp) v += lo ? shift_lo : shift_hi;
        f[(size_t)c * cap + s] = v;
    }
}
__global__ void k_slab_iota(int* ref, long long n) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s < n) ref[s] = (int)s;
}
__global__ void k_slab_count_owned(const double* ghost, long long n, int* out) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int owned = (s < n && ghost[s] == 0.0) ? 1 : 0;
    const int total = __syncthreads_count(owned);  // one atomic per CTA: 300 k same-address atomics cost 0.15 ms
    if (threadIdx.x == 0 && total) atomicAdd(out, total);
}

// ------------------------------------------------------------------ host helpers
static int slab_planes(sp_system* s, std::vector<SlabPlanes>& tabs, int* nplanes) {
    tabs.clear();
    SlabPlanes cur;
    cur.count = 0;
    cur.axis_plane = -1;
    int total = 0;
    for (size_t f = 0; f < s->fields.size(); f++) {
        SpField& fl = s->fields[f];
        if (fl.transient) continue;
        for (int c = 0; c < fl.ncomp; c++) {
            if (f == 0 && c == s->slab->axis) cur.axis_plane = cur.count;
            cur.p[cur.count++] = fl.d + (size_t)c * s->cap;
            total++;
            if (cur.count == SLAB_PLANES) {
                tabs.push_back(cur);
                cur.count = 0;
                cur.axis_plane = -1;
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

// totals of the two selections on the device: d_cnt[0] = to send down, d_cnt[1] = to send up (last exclusive-scan
// value + last flag), d_cnt[2], d_cnt[3] cleared for the receive counts
__global__ void k_slab_totals(const int* fd, const int* fu, const int* posd, const int* posu, long long ns, int* d_cnt) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        d_cnt[0] = ns > 0 ? posd[ns - 1] + fd[ns - 1] : 0;
        d_cnt[1] = ns > 0 ? posu[ns - 1] + fu[ns - 1] : 0;
        d_cnt[2] = 0;
        d_cnt[3] = 0;
    }
}

// exchange the send counts (already in d_cnt[0], d_cnt[1]) with both neighbours and bring all four numbers to the
// host with ONE synchronisation: h_cnt[0],[1] = my counts down/up, h_cnt[2],[3] = counts arriving from below / above
static int slab_exchange_counts(sp_system* s, long long* n_dn, long long* n_up, long long* from_below, long long* from_above) {
    SlabState* sl = s->slab;
    int below, above;
    slab_peers(sl, &below, &above);
    // Call order matters when below == above (1 or 2 ranks, periodic): messages between one pair of ranks are
    // matched in issue order, and what I send DOWN arrives at my lower neighbour FROM ABOVE.  So: send down,
    // send up, then receive from above, receive from below.
    SP_NCCL(s, g_nccl.GroupStart());
    if (below >= 0) SP_NCCL(s, g_nccl.Send(sl->d_cnt + 0, 1, ncclInt32, below, sl->comm, s->stream));
    if (above >= 0) SP_NCCL(s, g_nccl.Send(sl->d_cnt + 1, 1, ncclInt32, above, sl->comm, s->stream));
    if (above >= 0) SP_NCCL(s, g_nccl.Recv(sl->d_cnt + 3, 1, ncclInt32, above, sl->comm, s->stream));
    if (below >= 0) SP_NCCL(s, g_nccl.Recv(sl->d_cnt + 2, 1, ncclInt32, below, sl->comm, s->stream));
    SP_NCCL(s, g_nccl.GroupEnd());
    SP_CUDA(s, cudaMemcpyAsync(sl->h_cnt, sl->d_cnt, 4 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    *n_dn = sl->h_cnt[0];
    *n_up = sl->h_cnt[1];
    *from_below = below >= 0 ? sl->h_cnt[2] : 0;
    *from_above = above >= 0 ? sl->h_cnt[3] : 0;
    return SP_OK;
}

static int slab_exchange_payload(sp_system* s, long long send_dn, long long send_up, long long recv_lo, long long recv_hi,
                                 int nplanes) {
    SlabState* sl = s->slab;
    int below, above;
    slab_peers(sl, &below, &above);
    // same call order as slab_exchange_counts; empty messages are skipped on both sides (counts are known)
    SP_NCCL(s, g_nccl.GroupStart());
    if (below >= 0 && send_dn)
        SP_NCCL(s, g_nccl.Send(sl->sendbuf[0], (size_t)send_dn * nplanes, ncclFloat64, below, sl->comm, s->stream));
    if (above >= 0 && send_up)
        SP_NCCL(s, g_nccl.Send(sl->sendbuf[1], (size_t)send_up * nplanes, ncclFloat64, above, sl->comm, s->stream));
    if (above >= 0 && recv_hi)
        SP_NCCL(s, g_nccl.Recv(sl->recvbuf[1], (size_t)recv_hi * nplanes, ncclFloat64, above, sl->comm, s->stream));
    if (below >= 0 && recv_lo)
        SP_NCCL(s, g_nccl.Recv(sl->recvbuf[0], (size_t)recv_lo * nplanes, ncclFloat64, below, sl->comm, s->stream));
    SP_NCCL(s, g_nccl.GroupEnd());
    return SP_OK;
}

// select (flags fd/fu) -> scan -> pack all planes -> exchange -> append; returns counts
static int slab_round(sp_system* s, bool ghosts, long long* n_recv_lo, long long* n_recv_hi, long long* n_sent_dn,
                      long long* n_sent_up) {
    SlabState* sl = s->slab;
    const int B = 256;
    const long long n = s->n;
    int* fd = s->flags;
    int* fu = s->key_alt;
    int* posd = s->perm;
    int* posu = s->tmp_slot;
    double* X = s->fields[0].d;
    const double* ghost = s->fields[sl->f_ghost].d;
    SlabSel sel;
    if (sl->sel_valid && sl->sel_a < sl->sel_b && sl->sel_b <= n) {
        sel.a = sl->sel_a;
        sel.b = sl->sel_b;
    } else {
        sel.a = n;  // everything
        sel.b = n;
    }
    sel.n_sel = sel.a + (n - sel.b);
    const long long ns = sel.n_sel;
    if (ns > 0) {
        if (!ghosts)
            SP_LAUNCH(s, k_slab_classify, sp_blocks(ns, B), B, 0, s->g, sl->rank, sl->nranks, sl->gphase, sl->c0, sl->c1, X,
                      s->cap, ghost, sel, fd, fu);
        else
            SP_LAUNCH(s, k_slab_boundary, sp_blocks(ns, B), B, 0, s->g, sl->rank, sl->nranks, sl->gphase, sl->c0, sl->c1, X,
                      s->cap, ghost, sel, fd, fu);
        SP_CUDA(s, cudaMemcpyAsync(posd, fd, (size_t)ns * sizeof(int), cudaMemcpyDeviceToDevice, s->stream));
        SP_CUDA(s, cudaMemcpyAsync(posu, fu, (size_t)ns * sizeof(int), cudaMemcpyDeviceToDevice, s->stream));
        int rc = sp_exclusive_scan_i32(s, posd, ns);
        if (rc) return rc;
        if ((rc = sp_exclusive_scan_i32(s, posu, ns))) return rc;
    }
    // totals stay on the device: they go to the neighbours from there, and one synchronisation brings the send and
    // receive counts to the host together
    SP_LAUNCH(s, k_slab_totals, 1, 32, 0, fd, fu, posd, posu, ns, sl->d_cnt);
    long long send_dn = 0, send_up = 0, recv_lo = 0, recv_hi = 0;
    int rc = slab_exchange_counts(s, &send_dn, &send_up, &recv_lo, &recv_hi);
    if (rc) return rc;
    if (ghosts && n > 0) {
        SP_LAUNCH(s, k_slab_fill2, sp_blocks(n, B), B, 0, s->fields[sl->f_sdn].d, s->fields[sl->f_sup].d, -1.0, n);
        if (ns > 0)
            SP_LAUNCH(s, k_slab_record, sp_blocks(ns, B), B, 0, fd, fu, posd, posu, s->fields[sl->f_sdn].d,
                      s->fields[sl->f_sup].d, sel);
    }
    std::vector<SlabPlanes> tabs;
    int nplanes = 0;
    slab_planes(s, tabs, &nplanes);
    const long long biggest = std::max(std::max(send_dn, send_up), std::max(recv_lo, recv_hi));
    if ((rc = slab_ensure_buffers(s, biggest * nplanes + 16))) return rc;
    // pack
    int plane0 = 0;
    for (SlabPlanes& t : tabs) {
        if (send_dn) SP_LAUNCH(s, k_slab_pack, sp_blocks(ns, B), B, 0, t, fd, posd, sel, sl->sendbuf[0] + (size_t)plane0 * send_dn, send_dn);
        if (send_up) SP_LAUNCH(s, k_slab_pack, sp_blocks(ns, B), B, 0, t, fu, posu, sel, sl->sendbuf[1] + (size_t)plane0 * send_up, send_up);
        plane0 += t.count;
    }
    if (!ghosts && (send_dn || send_up)) SP_LAUNCH(s, k_slab_kill, sp_blocks(ns, B), B, 0, fd, fu, X, sel);
    if ((rc = slab_exchange_payload(s, send_dn, send_up, recv_lo, recv_hi, nplanes))) return rc;
    // append arrivals after the current particles
    const long long n_new = n + recv_lo + recv_hi;
    if (n_new > s->cap) {
        // growing reallocates every plane: finish the exchange first, then rebuild the plane tables
        SP_CUDA(s, cudaStreamSynchronize(s->stream));
        if ((rc = sp_ensure_capacity(s, n_new))) return rc;
        slab_planes(s, tabs, &nplanes);
    }
    // a particle that crossed the periodic boundary is shifted by one period on arrival
    const double shift_lo = (sl->periodic && sl->rank == 0) ? -sl->period : 0.0;              // came from the top rank
    const double shift_hi = (sl->periodic && sl->rank == sl->nranks - 1) ? sl->period : 0.0;  // came from rank 0
    plane0 = 0;
    for (SlabPlanes& t : tabs) {
        int* ref = plane0 == 0 ? s->ref : nullptr;  // arrivals are numbered by their slot
        if (recv_lo) SP_LAUNCH(s, k_slab_unpack, sp_blocks(recv_lo, B), B, 0, t, sl->recvbuf[0] + (size_t)plane0 * recv_lo, recv_lo, n, shift_lo, ref);
        if (recv_hi) SP_LAUNCH(s, k_slab_unpack, sp_blocks(recv_hi, B), B, 0, t, sl->recvbuf[1] + (size_t)plane0 * recv_hi, recv_hi, n + recv_lo, shift_hi, ref);
        plane0 += t.count;
    }
    if (ghosts) {
        double* gh = s->fields[sl->f_ghost].d;
        double* hx = s->fields[sl->f_hidx].d;
        double* sd = s->fields[sl->f_sdn].d;
        double* su = s->fields[sl->f_sup].d;
        if (recv_lo) SP_LAUNCH(s, k_slab_mark, sp_blocks(recv_lo, B), B, 0, gh, hx, sd, su, n, recv_lo, 1.0);
        if (recv_hi) SP_LAUNCH(s, k_slab_mark, sp_blocks(recv_hi, B), B, 0, gh, hx, sd, su, n + recv_lo, recv_hi, 2.0);
    }
    s->n = n_new;
    s->x_version++;
    *n_recv_lo = recv_lo;
    *n_recv_hi = recv_hi;
    *n_sent_dn = send_dn;
    *n_sent_up = send_up;
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
    if (!s || !id) return SP_ERR_INVALID;
    if (nranks < 1 || rank < 0 || rank >= nranks) return sp_fail(s, SP_ERR_INVALID, "bad rank / nranks");
    if (s->slab) return sp_fail(s, SP_ERR_STATE, "slab already initialised");
    if (s->n != 0) return sp_fail(s, SP_ERR_STATE, "sp_slab_init must be called before particles are added");
    SP_CUDA(s, cudaSetDevice(s->device));
    if (!nccl_load()) return sp_fail(s, SP_ERR_NCCL, g_nccl.err);
    SpGrid& g = s->g;
    const int axis = g.dim == 2 ? 1 : 2;  // slowest key axis
    if (g.lim[axis] < nranks) return sp_fail(s, SP_ERR_INVALID, "fewer cell layers along the slab axis than ranks");
    SlabState* sl = new SlabState();
    sl->rank = rank;
    sl->nranks = nranks;
    sl->periodic = periodic ? 1 : 0;
    sl->axis = axis;
    sl->gphase = g.phase[axis];
    sl->glim = g.lim[axis];
    const long long base = sl->glim / nranks, rem = sl->glim % nranks;
    sl->c0 = rank * base + std::min<long long>(rank, rem);
    sl->c1 = sl->c0 + base + (rank < rem ? 1 : 0);
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
    SP_CUDA(s, sp_dmalloc(&sl->d_cnt, 16 * sizeof(int)));
    SP_CUDA(s, cudaHostAlloc(&sl->h_cnt, 16 * sizeof(int), cudaHostAllocDefault));
    // local cell window: owned layers [c0, c1) plus one ghost layer per side
    g.phase[axis] = sl->gphase + sl->c0 - 1;
    g.lim[axis] = (sl->c1 - sl->c0) + 2;
    g.key_max = g.lim[0] * g.lim[1] * g.lim[2];
    // the local window (with its two ghost layers) can be larger than the global grid when nranks is small
    SP_CUDA(s, sp_dfree(s, s->cell_start));
    SP_CUDA(s, sp_dfree(s, s->cell_fill));
    s->cell_start = s->cell_fill = nullptr;
    SP_CUDA(s, sp_dmalloc(&s->cell_start, (size_t)(g.key_max + 3) * sizeof(int)));
    SP_CUDA(s, sp_dmalloc(&s->cell_fill, (size_t)(g.key_max + 3) * sizeof(int)));
    SP_CUDA(s, cudaMemset(s->cell_start, 0, (size_t)(g.key_max + 3) * sizeof(int)));
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
    if ((rc = sp_add_field(s, "_hidx", 1, &fid))) return rc;
    sl->f_hidx = fid;
    if ((rc = sp_add_field(s, "_sdn", 1, &fid))) return rc;
    sl->f_sdn = fid;
    if ((rc = sp_add_field(s, "_sup", 1, &fid))) return rc;
    sl->f_sup = fid;
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
    SlabState* sl = s->slab;
    int rc = sp_time_begin(s);
    if (rc) return rc;
    long long rl, rh, sd, su;
    const bool trace = slab_trace_on();
    double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    if (trace) {
        cudaStreamSynchronize(s->stream);
        t0 = slab_now();
    }
    // 1+2: drop old ghosts, migrate
    if ((rc = slab_round(s, false, &rl, &rh, &sd, &su))) return rc;
    if (trace) {
        cudaStreamSynchronize(s->stream);
        t1 = slab_now();
    }
    // 3: ghost halo
    if ((rc = slab_round(s, true, &rl, &rh, &sd, &su))) return rc;
    if (trace) {
        cudaStreamSynchronize(s->stream);
        t2 = slab_now();
    }
    sl->n_ghost[0] = rl;
    sl->n_ghost[1] = rh;
    sl->n_send[0] = sd;
    sl->n_send[1] = su;
    // 4: local build; the reference numbering has no meaning across ranks: slots are renumbered in place
    // (ref is already the slot order unless the host added or re-ordered particles since the last rebuild)
    if (s->n && !sl->sel_valid) SP_LAUNCH(s, k_slab_iota, sp_blocks(s->n, 256), 256, 0, s->ref, s->n);
    if ((rc = sp_build_cells(s))) return rc;
    if (s->n) SP_LAUNCH(s, k_slab_iota, sp_blocks(s->n, 256), 256, 0, s->ref, s->n);
    s->identity_order = true;
    if (trace) {
        cudaStreamSynchronize(s->stream);
        t3 = slab_now();
    }
    // owned count, and the slot window of the next rebuild's selection passes (three cell layers per side)
    SP_CUDA(s, cudaMemsetAsync(sl->d_cnt + 8, 0, sizeof(int), s->stream));
    if (s->n)
        SP_LAUNCH(s, k_slab_count_owned, sp_blocks(s->n, 256), 256, 0, s->fields[sl->f_ghost].d, s->n, sl->d_cnt + 8);
    SP_CUDA(s, cudaMemcpyAsync(sl->h_cnt + 8, sl->d_cnt + 8, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    const long long layer = s->g.lim[0] * (sl->axis == 2 ? s->g.lim[1] : 1);  // cells per layer
    const long long nl = s->g.lim[sl->axis];
    const bool window = nl >= 8;
    if (window) {
        SP_CUDA(s, cudaMemcpyAsync(sl->h_cnt + 9, s->cell_start + 3 * layer + 1, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        SP_CUDA(s, cudaMemcpyAsync(sl->h_cnt + 10, s->cell_start + (nl - 3) * layer + 1, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    }
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    sl->sel_valid = window;
    sl->sel_a = sl->h_cnt[9];
    sl->sel_b = sl->h_cnt[10];
    sl->n_owned = sl->h_cnt[8];
    if (trace) {
        const double t4 = slab_now();
        sl->trace_s[0] += t1 - t0;
        sl->trace_s[1] += t2 - t1;
        sl->trace_s[2] += t3 - t2;
        sl->trace_s[3] += t4 - t3;
        if (++sl->trace_calls % 20 == 0)
            fprintf(stderr, "[slab trace rank %d] calls=%lld migrate=%.3f ms ghosts=%.3f ms build=%.3f ms tail=%.3f ms (n=%lld ghosts=%lld+%lld)\n",
                    sl->rank, sl->trace_calls, 1e3 * sl->trace_s[0] / sl->trace_calls, 1e3 * sl->trace_s[1] / sl->trace_calls,
                    1e3 * sl->trace_s[2] / sl->trace_calls, 1e3 * sl->trace_s[3] / sl->trace_calls, (long long)s->n, rl, rh);
    }
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
    const long long sd = sl->n_send[0], su = sl->n_send[1], rl = sl->n_ghost[0], rh = sl->n_ghost[1];
    const long long biggest = std::max(std::max(sd, su), std::max(rl, rh));
    if ((rc = slab_ensure_buffers(s, biggest * ncomp_total + 16))) return rc;
    const int B = 256;
    const long long n = s->n;
    const double* sdn = s->fields[sl->f_sdn].d;
    const double* sup = s->fields[sl->f_sup].d;
    int c0 = 0;
    for (int k = 0; k < nfields && n > 0; k++) {
        SpField& f = s->fields[fields[k]];
        if (sd) SP_LAUNCH(s, k_slab_refresh_pack, sp_blocks(n, B), B, 0, f.d, s->cap, f.ncomp, sdn, n, sl->sendbuf[0] + (size_t)c0 * sd, sd);
        if (su) SP_LAUNCH(s, k_slab_refresh_pack, sp_blocks(n, B), B, 0, f.d, s->cap, f.ncomp, sup, n, sl->sendbuf[1] + (size_t)c0 * su, su);
        c0 += f.ncomp;
    }
    if ((rc = slab_exchange_payload(s, sd, su, rl, rh, ncomp_total))) return rc;
    const double shift_lo = (sl->periodic && sl->rank == 0) ? -sl->period : 0.0;
    const double shift_hi = (sl->periodic && sl->rank == sl->nranks - 1) ? sl->period : 0.0;
    c0 = 0;
    for (int k = 0; k < nfields && n > 0; k++) {
        SpField& f = s->fields[fields[k]];
        const int axis_comp = fields[k] == 0 ? sl->axis : -1;
        f.version++;  // ghost values change (a field that is zero everywhere stays zero: known_zero is kept)
        if (fields[k] == 0) s->x_version++;
        if (rl || rh)
            SP_LAUNCH(s, k_slab_refresh_unpack, sp_blocks(n, B), B, 0, f.d, s->cap, f.ncomp, s->fields[sl->f_ghost].d,
                      s->fields[sl->f_hidx].d, n, sl->recvbuf[0] + (size_t)c0 * rl, rl, sl->recvbuf[1] + (size_t)c0 * rh, rh,
                      axis_comp, shift_lo, shift_hi);
        c0 += f.ncomp;
    }
    return sp_time_end(s);
}

int32_t sp_slab_num_owned(sp_system* s, int64_t* n_owned) {
    if (!s || !n_owned) return SP_ERR_INVALID;
    if (!s->slab) return sp_fail(s, SP_ERR_STATE, "not a slab system");
    *n_owned = s->slab->n_owned;
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
