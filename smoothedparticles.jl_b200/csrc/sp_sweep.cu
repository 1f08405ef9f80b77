// sp_sweep.cu — apply! / apply_binary! / apply_unary! on the device (reference src/core.jl:94-161).
//
// Gather formulation: one thread owns particle p (a slot of the cell-sorted SoA planes), scans the
// reference's 9/27 linear-offset cells, applies the exact distance predicate and accumulates in
// registers; nothing of q is written, so the sweep is race-free by construction (core.jl:122-123).
//
// Candidate cells are `key + dkey` with only the global range check 1 <= key <= key_max
// (core.jl:97-98) — including the row wrap-around at domain edges.  Cells with consecutive keys are
// contiguous in the sorted planes, so the three di = -1,0,1 cells of one (dj,dk) row form ONE
// contiguous slot range; the default order walks those 3 (2-D) / 9 (3-D) ranges, SP_FLAG_STRICT_ORDER
// walks the 9/27 cells in key_diff order (di outermost), which with the descending in-cell order is
// exactly the reference's accumulation order.
//
// Kernels in this file, in the order a time step meets them (DESIGN.md §4 has the measurements):
//   k_prefilter_coords       FP32 cell-unit coordinates for the conservative pre-filter (also written by the cell-list permute)
//   k_nbr_build_sweep<Op>    DEFAULT for the first pair sweep after a position change: FP32x2 candidate scan with one
//                            conservative threshold -> the "maybe" slots go to the target's own column of the warp-tiled list ->
//                            replay of that column with the exact FP64 predicate (d2 > T2) inside the pair body, compacted in
//                            place: leaves the exact neighbour list in visiting order AND the operator's result
//   k_nbr_build              the list alone (operators without FUSED_BUILD, SP_FLAG_UNFUSED_BUILD): two thresholds, only the
//                            thin shell between them takes the FP64 test
//   k_sweep_list<Op>         every later sweep of the same position version: replay of the cached list (ids prefetched one
//                            trip ahead)
//   k_sweep<Op,STRICT>       no lists: the candidate scan with the exact predicate per candidate; STRICT = reference
//                            accumulation order (SP_FLAG_STRICT_ORDER), also the fallback for targets whose list overflowed
//   k_sweep_tile<Op>         the one alternative kept (SP_FLAG_TILE_KERNEL): shared-memory tile staged by TMA bulk copies
//   k_unary<Op>              apply_unary! (core.jl:138-142)
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include <type_traits>
#include "sp_internal.cuh"
#include "sp_ops.cuh"

struct SweepCtx {
    const double *x, *y, *z;
    const float *ux, *uy, *uz;  // cell-unit FP32 coordinates (x - lo)/h for the conservative pre-filter
    const int* cell_start;
    float thr;                  // pre-filter threshold on the FP32 squared distance in cell units: above = not a neighbour
    float thr_lo;               // at or below = certainly a neighbour (k_nbr_build)
    int capk;                   // entries per target in the cached neighbour lists (multiple of 32)
    int n;                      // slots the launch covers
    const int* alive;           // device: slots [0, *alive) are alive, [*alive, n) is the dead tail of culled particles
    // slab systems: the outermost ghost layer on either side only FEEDS sums (its own neighbourhood is incomplete), so its
    // particles are not swept as targets: targets are the slots of the cells [tgt_key_lo, tgt_key_hi) (0, 0 = all)
    int tgt_key_lo, tgt_key_hi;
};
__device__ __forceinline__ bool sp_is_target(const SweepCtx& c, int i) {
    if (i >= *c.alive) return false;  // the dead tail [alive, n) takes no part (sp_internal.cuh)
    if (c.tgt_key_hi && (i < c.cell_start[c.tgt_key_lo] || i >= c.cell_start[c.tgt_key_hi])) return false;
    return true;
}

// Visit every candidate slot j of particle (xi,yi,zi): f(j, dx, dy, dz, d2).
template <bool STRICT, class F>
__device__ __forceinline__ void sp_for_candidates(const SpGrid& g, const SweepCtx& c, double xi, double yi, double zi,
                                                  F&& f) {
    const long long key = sp_find_key(g, xi, yi, zi);  // core.jl:95 recomputes the key from the current x
    const long long L1 = g.lim[0], L12 = g.lim[0] * g.lim[1];
    const int nk = (g.dim == 2) ? 1 : 3;
    if (STRICT) {
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++)
                for (int dk = (nk == 1 ? 0 : -1); dk <= (nk == 1 ? 0 : 1); dk++) {
                    const long long nkey = key + di + L1 * dj + L12 * dk;
                    if (nkey < 1 || nkey > g.key_max) continue;
                    const int jb = c.cell_start[nkey], je = c.cell_start[nkey + 1];
                    for (int j = jb; j < je; j++) {
                        double dx = __dsub_rn(xi, c.x[j]), dy = __dsub_rn(yi, c.y[j]), dz = __dsub_rn(zi, c.z[j]);
                        f(j, dx, dy, dz, sp_d2(dx, dy, dz));
                    }
                }
    } else {
        for (int dk = (nk == 1 ? 0 : -1); dk <= (nk == 1 ? 0 : 1); dk++)
            for (int dj = -1; dj <= 1; dj++) {
                const long long mid = key + L1 * dj + L12 * dk;
                long long klo = mid - 1, khi = mid + 1;
                if (klo < 1) klo = 1;
                if (khi > g.key_max) khi = g.key_max;
                if (klo > khi) continue;
                const int jb = c.cell_start[klo], je = c.cell_start[khi + 1];
                for (int j = jb; j < je; j++) {
                    double dx = __dsub_rn(xi, c.x[j]), dy = __dsub_rn(yi, c.y[j]), dz = __dsub_rn(zi, c.z[j]);
                    f(j, dx, dy, dz, sp_d2(dx, dy, dz));
                }
            }
    }
}

// q-field accessor of the reference-order kernel: plane k of the neighbour at slot j, through L1/L2
template <int NQ>
struct QGlobal {
    const double* const* qp;
    int j;
    __device__ __forceinline__ double operator()(int k) const { return qp[k][j]; }
};

// q-field values of one neighbour already in registers (list replay kernel)
template <int NQ>
struct QRegs {
    double v[NQ > 0 ? NQ : 1];
    __device__ __forceinline__ double operator()(int k) const { return v[k]; }
};

// ---- reference-order kernel (SP_FLAG_STRICT_ORDER, and the parity views): one thread per particle,
// candidates read straight from the sorted planes.
__device__ __forceinline__ double sp_sqrt_fast(double a);

template <class Op, bool STRICT>
__global__ void __launch_bounds__(128) k_sweep(SpGrid g, SweepCtx c, typename Op::Params P, int self_flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (!sp_is_target(c, i)) return;
    if (!Op::active(P, i)) return;
    const double xi = c.x[i], yi = c.y[i], zi = c.z[i];
    typename Op::PS p;
    typename Op::Acc acc;
    Op::load(P, i, xi, yi, zi, p, acc);
    sp_for_candidates<STRICT>(g, c, xi, yi, zi, [&](int j, double dx, double dy, double dz, double d2) {
        // (r > h || p == q) && continue   (core.jl:105), r = sqrt_rn(d2)  <=>  d2 > T2
        if (d2 > g.T2 || j == i) return;
        QGlobal<Op::NQ> q{P.qp, j};
        Op::pair(P, p, q, dx, dy, dz, STRICT ? sqrt(d2) : sp_sqrt_fast(d2), acc);
    });
    if (self_flag & 1) Op::self(P, p, acc);
    Op::store(P, i, p, acc);
}

// ---- tile kernel (experimental, SP_FLAG_TILE_KERNEL).
// A CTA owns TP consecutive slots of the cell-sorted planes (its targets, one per thread).  For each of
// the 3 (2-D) / 9 (3-D) stencil rows the candidates of ALL its targets form one contiguous slot range;
// the ranges are cut into segments and staged, batch by batch, into shared memory as SoA Float64 planes
// (x, y, z + the NQ planes the operator reads of q) by TMA 1-D bulk copies (cp.async.bulk -> UBLKCP)
// completing on an mbarrier: every plane segment is contiguous in HBM, so one elected thread issues
// nseg*(3+NQ) bulk copies and the LSU pipe stays free.  Per batch:
//   phase 1  every thread walks its own cells' sub-range of each staged segment, evaluates the exact
//            un-fused distance predicate (4 candidates in flight), and appends accepted candidates
//            (16-bit tile index) to a private list in shared memory — a tight, branch-free loop;
//   phase 2  every thread runs the operator body over its list — the expensive part (sqrt, kernel, FMAs)
//            executes with nearly all lanes active instead of ~15 % of them.
// Any density works: a row that does not fit is split over several batches, a list that fills is flushed.
#define TILE_MAX_SEG 16
struct TileSeg {
    int pos;   // first global slot staged (even: 16-byte aligned for the bulk copy)
    int cnt;   // slots staged (even)
    int soff;  // offset of the segment in the tile (even)
    int row;   // stencil row index (0..8)
};

__device__ __forceinline__ unsigned sp_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sp_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sp_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void sp_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sp_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sp_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     sp_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(sp_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void sp_mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(sp_smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

template <int NQ, int CAP>
struct QTile {
    const double* sq;  // [NQ][CAP] planes of the current buffer, already offset by the tile index
    __device__ __forceinline__ double operator()(int k) const { return sq[k * CAP]; }
};

// r = sqrt(d2) for the operator bodies of the tile kernel: MUFU.RSQ64H seed + two Newton steps, ~1 ulp,
// branch-free (the neighbour DECISION never uses it: that is the exact d2 > T2 test).
__device__ __forceinline__ double sp_sqrt_fast(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double h = 0.5 * a;
    y = y * fma(-h, y * y, 1.5);
    y = y * fma(-h, y * y, 1.5);
    const double r = a * y;
    return a > 0.0 ? r : 0.0;  // coincident particles: r = 0 (rsqrt(0) = inf)
}

template <class Op, int TP, int LCAP, int CAPB, int NROWS, int RPB, int MINB>
__global__ void __launch_bounds__(TP, MINB) k_sweep_tile(SpGrid g, SweepCtx c, typename Op::Params P, int self_flag) {
    constexpr int NQ = Op::NQ;
    constexpr int NPL = 6 + NQ;                                    // planes per buffer: 3+NQ doubles, 3 floats
    constexpr size_t BUF_BYTES = (size_t)CAPB * ((3 + NQ) * 8 + 12);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned short* list = reinterpret_cast<unsigned short*>(smem_raw + 2 * BUF_BYTES);  // [LCAP][TP]
    __shared__ TileSeg segs[2][TILE_MAX_SEG];
    __shared__ int s_nseg[2], s_fill[2];
    __shared__ int s_rowpos[NROWS], s_rowend[NROWS];
    __shared__ long long s_kmin, s_kmax;
    __shared__ int s_any;
    __shared__ __align__(8) unsigned long long s_bar[2];

    const int tid = threadIdx.x;
    const int i = blockIdx.x * TP + tid;
    const bool act = (i < *c.alive) && Op::active(P, i);
    double xi = 0.0, yi = 0.0, zi = 0.0;
    float ui = 0.f, vi = 0.f, wi = 0.f;
    long long key = 0;
    typename Op::PS p;
    typename Op::Acc acc;
    if (tid == 0) {
        s_kmin = 0x7fffffffffffffffLL;
        s_kmax = -0x7fffffffffffffffLL;
        s_any = 0;
        sp_mbar_init(&s_bar[0], 1);
        sp_mbar_init(&s_bar[1], 1);
    }
    __syncthreads();
    if (act) {
        xi = c.x[i]; yi = c.y[i]; zi = c.z[i];
        ui = c.ux[i]; vi = c.uy[i]; wi = c.uz[i];
        key = sp_find_key(g, xi, yi, zi);  // core.jl:95
    }
    {
        // key span of the active targets: warp reduction, then one shared atomic per warp
        long long kmn = act ? key : 0x7fffffffffffffffLL, kmx = act ? key : -0x7fffffffffffffffLL;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const long long a = __shfl_xor_sync(0xffffffffu, kmn, d), b = __shfl_xor_sync(0xffffffffu, kmx, d);
            kmn = a < kmn ? a : kmn;
            kmx = b > kmx ? b : kmx;
        }
        if ((tid & 31) == 0 && kmn <= kmx) {
            atomicMin(&s_kmin, kmn);
            atomicMax(&s_kmax, kmx);
            s_any = 1;
        }
    }
    // own candidate slot range of every stencil row (registers; rows are unrolled at compile time)
    const long long L1 = g.lim[0], L12 = g.lim[0] * g.lim[1];
    int jb[NROWS], je[NROWS];
#pragma unroll
    for (int row = 0; row < NROWS; row++) {
        const int dj = row % 3 - 1, dk = (NROWS == 3) ? 0 : row / 3 - 1;
        const long long mid = key + L1 * dj + L12 * dk;
        long long klo = mid - 1, khi = mid + 1;
        if (klo < 1) klo = 1;
        if (khi > g.key_max) khi = g.key_max;
        jb[row] = je[row] = 0;
        if (act && klo <= khi) {
            jb[row] = c.cell_start[klo];
            je[row] = c.cell_start[khi + 1];
        }
    }
    if (act) Op::load(P, i, xi, yi, zi, p, acc);
    __syncthreads();
    if (!s_any) return;  // e.g. a block of wall particles under a fluid-only operator
    // slot span of the whole block for every row: NROWS threads fetch them concurrently
    if (tid < NROWS) {
        const int dj = tid % 3 - 1, dk = (NROWS == 3) ? 0 : tid / 3 - 1;
        long long klo = s_kmin - 1 + L1 * dj + L12 * dk, khi = s_kmax + 1 + L1 * dj + L12 * dk;
        if (klo < 1) klo = 1;
        if (khi > g.key_max) khi = g.key_max;
        int a = 0, b = 0;
        if (klo <= khi) {
            a = c.cell_start[klo];
            b = c.cell_start[khi + 1];
        }
        s_rowpos[tid] = a;
        s_rowend[tid] = b;
    }
    __syncthreads();

    const double T2 = g.T2;
    const float thr = c.thr;
    unsigned short* lp = list + tid;  // next free list entry of this thread
    int self_s = -1;                  // tile index of p itself in the current batch
    double *sx, *sy, *sz, *sq;        // planes of the buffer being processed
    float *fx, *fy, *fz;
    auto set_buf = [&](int b) {
        sx = reinterpret_cast<double*>(smem_raw + (size_t)b * BUF_BYTES);
        sy = sx + CAPB;
        sz = sy + CAPB;
        sq = sz + CAPB;
        fx = reinterpret_cast<float*>(sq + (size_t)NQ * CAPB);
        fy = fx + CAPB;
        fz = fy + CAPB;
    };

    // ---- phase 2: exact un-fused predicate (core.jl:104-105) and the operator body, 4 list entries at a time,
    // branch-free: rejected / padding entries are computed and then discarded by a select.
    auto flush = [&]() {
        const int cnt = (int)((lp - (list + tid)) / TP);
        lp = list + tid;
        if (self_flag & 256) return;
        for (int k = 0; k < cnt; k += 4) {
            int s4[4];
            bool ok[4];
            double dx[4], dy[4], dz[4], d2[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                ok[u] = (k + u) < cnt;
                s4[u] = ok[u] ? (int)list[(k + u) * TP + tid] : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                dx[u] = __dsub_rn(xi, sx[s4[u]]);
                dy[u] = __dsub_rn(yi, sy[s4[u]]);
                dz[u] = __dsub_rn(zi, sz[s4[u]]);
                d2[u] = sp_d2(dx[u], dy[u], dz[u]);
                // (r > h || p == q) && continue  <=>  d2 > T2 with r = sqrt_rn(d2)
                ok[u] = ok[u] && !(d2[u] > T2) && s4[u] != self_s;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                typename Op::Acc t = acc;
                QTile<NQ, CAPB> q{sq + s4[u]};
                Op::pair(P, p, q, dx[u], dy[u], dz[u], sp_sqrt_fast(d2[u]), t);
                if (ok[u]) acc = t;
            }
        }
    };

    // ---- batch planner (block-uniform): up to RPB stencil rows, or as much of them as fits one buffer
    int row = 0, pos = 0;
    auto plan = [&](int b) -> int {
        int fill = 0, nseg = 0;
        while (row < NROWS && nseg == 0) {  // skip row groups without candidates
            const int row_stop = min(NROWS, (row / RPB + 1) * RPB);
            while (row < row_stop && nseg < TILE_MAX_SEG) {
                const int rend = s_rowend[row];
                if (pos < s_rowpos[row]) pos = s_rowpos[row];
                if (pos >= rend) {
                    row++;
                    pos = 0;
                    continue;
                }
                const int room = (CAPB - fill) & ~3;
                if (room < 8) break;
                const int pos_al = pos & ~3;                // 16-byte aligned sources for the 8 B and 4 B planes
                const int want = (rend - pos_al + 3) & ~3;  // multiple of 4 slots
                const int take = min(want, room);
                if (tid == 0) segs[b][nseg] = TileSeg{pos_al, take, fill, row};
                nseg++;
                fill += take;
                pos = pos_al + take;
            }
        }
        if (tid == 0) {
            s_nseg[b] = nseg;
            s_fill[b] = fill;
        }
        return nseg;
    };
    // TMA bulk copies of one batch: one per (segment, plane), spread over the lanes of warp 0
    auto issue = [&](int b, int nseg, int fill) {
        if (tid < 32) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                sp_mbar_expect_tx(&s_bar[b], (unsigned)fill * ((3 + NQ) * 8u + 12u));
            }
            __syncwarp();
            unsigned char* base = smem_raw + (size_t)b * BUF_BYTES;
            for (int w = tid; w < nseg * NPL; w += 32) {
                const int sgi = w / NPL, pl = w % NPL;
                const TileSeg sg = segs[b][sgi];
                if (pl < 3 + NQ) {
                    const double* src = pl == 0 ? c.x : pl == 1 ? c.y : c.z;
#pragma unroll
                    for (int k = 0; k < NQ; k++)
                        if (pl == 3 + k) src = P.qp[k];  // static indices: keeps Params in the constant bank
                    sp_bulk_g2s(reinterpret_cast<double*>(base) + (size_t)pl * CAPB + sg.soff, src + sg.pos,
                                (unsigned)sg.cnt * 8u, &s_bar[b]);
                } else {
                    const int f = pl - (3 + NQ);
                    const float* src = f == 0 ? c.ux : f == 1 ? c.uy : c.uz;
                    sp_bulk_g2s(reinterpret_cast<float*>(base + (size_t)(3 + NQ) * CAPB * 8) + (size_t)f * CAPB + sg.soff,
                                src + sg.pos, (unsigned)sg.cnt * 4u, &s_bar[b]);
                }
            }
        }
    };

    int cur = 0;
    unsigned parity[2] = {0u, 0u};
    int nseg_cur = plan(0);
    __syncthreads();  // segs[0] visible to warp 0
    if (nseg_cur) issue(0, nseg_cur, s_fill[0]);
    while (nseg_cur) {
        // prefetch the next batch into the other buffer (it was released by the barrier ending the last round)
        const int nseg_next = plan(cur ^ 1);
        __syncthreads();  // segs[cur^1] visible; also orders this round after the previous flush
        if (nseg_next) issue(cur ^ 1, nseg_next, s_fill[cur ^ 1]);
        sp_mbar_wait(&s_bar[cur], parity[cur]);
        parity[cur] ^= 1u;
        set_buf(cur);
        if (act && !(self_flag & 512)) {
            self_s = -1;
            int sgi = 0;
#pragma unroll
            for (int r = 0; r < NROWS; r++) {
                while (sgi < nseg_cur && segs[cur][sgi].row == r) {
                    const TileSeg sg = segs[cur][sgi];
                    sgi++;
                    const int lo = max(jb[r], sg.pos), hi = min(je[r], sg.pos + sg.cnt);
                    if (lo >= hi) continue;
                    const int shift = sg.soff - sg.pos;
                    if (r == NROWS / 2 && i >= lo && i < hi) self_s = i + shift;
                    // ---- phase 1: conservative FP32 pre-filter, 4 candidates per 128-bit shared load
                    int s_lo = lo + shift;
                    const int s_hi = hi + shift;
                    while (s_lo < s_hi) {
                        int s_end = s_hi;
                        const int have = (int)((lp - (list + tid)) / TP);
                        if (have + (s_end - s_lo) > LCAP) {
                            if (have > 0) {
                                flush();
                                continue;
                            }
                            s_end = s_lo + LCAP;
                        }
#pragma unroll 2
                        for (int gs = s_lo & ~3; gs < s_end; gs += 4) {
                            const float4 qx = *reinterpret_cast<const float4*>(fx + gs);
                            const float4 qy = *reinterpret_cast<const float4*>(fy + gs);
                            const float4 qz = *reinterpret_cast<const float4*>(fz + gs);
                            const float ax[4] = {qx.x, qx.y, qx.z, qx.w}, ay[4] = {qy.x, qy.y, qy.z, qy.w},
                                        az[4] = {qz.x, qz.y, qz.z, qz.w};
                            int pass[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const float ddx = ui - ax[u], ddy = vi - ay[u], ddz = wi - az[u];
                                const float dd = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));
                                const int s = gs + u;
                                // !(d2 > thr) keeps NaN distances, which the reference also lets through
                                pass[u] = (!(dd > thr) && s >= s_lo && s < s_end) ? 1 : 0;
                            }
                            const int o1 = pass[0], o2 = o1 + pass[1], o3 = o2 + pass[2];
                            if (pass[0]) lp[0] = (unsigned short)gs;
                            if (pass[1]) lp[o1 * TP] = (unsigned short)(gs + 1);
                            if (pass[2]) lp[o2 * TP] = (unsigned short)(gs + 2);
                            if (pass[3]) lp[o3 * TP] = (unsigned short)(gs + 3);
                            lp += (o3 + pass[3]) * TP;
                        }
                        s_lo = s_end;
                    }
                }
            }
            flush();
        }
        __syncthreads();  // everyone is done with buffer `cur` and its segment table before they are re-planned
        cur ^= 1;
        nseg_cur = nseg_next;
    }
    if (act) {
        if (self_flag & 1) Op::self(P, p, acc);
        Op::store(P, i, p, acc);
    }
}

// ---- neighbour-list cache (default path): build once per position version, replay per operator.
// Positions do not change between the pair sweeps of one time step (WCSPH: balance_of_mass! and internal_force!;
// ISPH: viscous_force!, div_L_lambda!, every CG mat-vec, internal_force!), so the op-independent part of
// apply_binary! (core.jl:94-110: key, candidate cells, dist, r > h, identity) is evaluated ONCE by k_nbr_build and
// its result — the exact accepted neighbour slots of every target, in visiting order — is stored in HBM:
//   ids[((i >> 5) * CAPK + k) * 32 + (i & 31)]   k-th neighbour of target slot i (warp-tiled: a warp's k-th
//                                                 entries are one 128-byte line)
//   cnt[i]                                        number of neighbours (may exceed CAPK: such a target is swept
//                                                 by the exact candidate scan instead, nothing is dropped)
// k_sweep_list<Op> then runs only the pair bodies: no candidate loop, no predicate, ~88 % active lanes.
// The cache is keyed on sp_system::x_version, which every position write / re-sort / resize bumps.
#define SP_NBR_CAPK_DEFAULT 64 /* grows by itself when a build reports longer lists (3-D runs with h = 3 dr: ~113) */
#define SP_NBR_CAPK_MAX 512

// The FP32 distance (cell units) classifies a candidate three ways — the rounding bound delta of the pre-filter is
// symmetric (see launch_sweep): dd <= 1 - delta is certainly a neighbour, dd > 1 + delta certainly is not, and only
// the thin shell in between (~0.1 % of the candidates) needs the exact FP64 predicate.  So the build kernel reads
// FP64 positions only for those few.
__global__ void __launch_bounds__(128, 5) k_nbr_build(SpGrid g, SweepCtx c, int* __restrict__ cnt, int* __restrict__ ids,
                                                      int* __restrict__ max_cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *c.alive) return;  // the dead tail [alive, n) takes no part (sp_internal.cuh)
    if (!sp_is_target(c, i)) {  // outermost ghost layer of a slab: no list
        cnt[i] = 0;
        return;
    }
    const float ui = c.ux[i], vi = c.uy[i], wi = c.uz[i];
    const double T2 = g.T2;
    unsigned long long ui2, vi2, wi2, thr2, thr_lo2;
    asm("mov.b64 %0, {%1,%1};" : "=l"(ui2) : "f"(ui));
    asm("mov.b64 %0, {%1,%1};" : "=l"(vi2) : "f"(vi));
    asm("mov.b64 %0, {%1,%1};" : "=l"(wi2) : "f"(wi));
    asm("mov.b64 %0, {%1,%1};" : "=l"(thr2) : "f"(c.thr));
    asm("mov.b64 %0, {%1,%1};" : "=l"(thr_lo2) : "f"(c.thr_lo));
    // entry k of target i lives at ((i / 32) * CAPK + k) * 32 + i % 32
    const int capk = c.capk;
    int* col = ids + ((size_t)(i >> 5) * capk << 5) + (i & 31);
    int n_out = 0;
    unsigned m0 = 0u, m1 = 0u, m2 = 0u;  // pending chunks: candidates that may be neighbours (newest in m0)
    unsigned s0 = 0u, s1 = 0u, s2 = 0u;  // ... of which certainly neighbours
    int b0 = 0, b1 = 0, b2 = 0;
    const double xi = c.x[i], yi = c.y[i], zi = c.z[i];
    // exact FP64 predicate for the ambiguous candidates of a chunk: returns the bits that are NOT neighbours.
    // (r > h) && continue (core.jl:105)  <=>  d2 > T2 with r = sqrt_rn(d2), decided in FP64 as the reference does
    auto reject = [&](unsigned amb, int base) -> unsigned {
        unsigned out = 0u;
        while (amb) {
            const unsigned bit = amb & (0u - amb);
            const int j = base + __ffs(amb) - 1;
            amb ^= bit;
            const double dx = __dsub_rn(xi, c.x[j]), dy = __dsub_rn(yi, c.y[j]), dz = __dsub_rn(zi, c.z[j]);
            if (sp_d2(dx, dy, dz) > T2) out |= bit;
        }
        return out;
    };
    auto append = [&](unsigned m, int base) {
        while (m) {
            const int j = base + __ffs(m) - 1;
            m &= m - 1;
            if (n_out < capk) col[n_out << 5] = j;
            n_out++;
        }
    };
    auto drain = [&]() {
        // first the few exact tests of all three chunks (kept out of the append loop: on a lattice at rest ~6 of 32
        // neighbours sit at r == h and are ambiguous in FP32), then the appends, which touch no particle data
        m2 &= ~reject(m2 & ~s2, b2);
        m1 &= ~reject(m1 & ~s1, b1);
        m0 &= ~reject(m0 & ~s0, b0);
        append(m2, b2);
        append(m1, b1);
        append(m0, b0);
        m0 = m1 = m2 = 0u;
    };
    const long long key = sp_find_key(g, xi, yi, zi);  // core.jl:95
    const long long L1 = g.lim[0], L12 = g.lim[0] * g.lim[1];
    const int nk = (g.dim == 2) ? 0 : 1;
    for (int dk = -nk; dk <= nk; dk++) {
        for (int dj = -1; dj <= 1; dj++) {
            const long long mid = key + L1 * dj + L12 * dk;
            long long klo = mid - 1, khi = mid + 1;
            if (klo < 1) klo = 1;
            if (khi > g.key_max) khi = g.key_max;
            if (klo > khi) continue;
            const int jb = c.cell_start[klo], je = c.cell_start[khi + 1];
            // chunks of 32 slots starting at a multiple of 4: one LDG.128 per plane brings 4 candidates, the distance
            // runs in packed FP32x2 (FADD2/FMUL2/FFMA2), and the two classifications are read off the SIGN of
            // thr - dd and thr_lo - dd, shifted into the masks by one funnel shift per candidate (no compare/select)
            for (int j0 = jb & ~3; j0 < je; j0 += 32) {
                const int nj = min(32, je - j0);
                const int nq = (nj + 3) >> 2;
                unsigned not_maybe = 0u, is_sure = 0u;  // candidate k of the chunk lands in bit 4*nq-1-k
                const ulonglong2* px = reinterpret_cast<const ulonglong2*>(c.ux + j0);
                const ulonglong2* py = reinterpret_cast<const ulonglong2*>(c.uy + j0);
                const ulonglong2* pz = reinterpret_cast<const ulonglong2*>(c.uz + j0);
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    if (q >= nq) break;
                    const ulonglong2 qx = __ldg(px + q), qy = __ldg(py + q), qz = __ldg(pz + q);
                    const unsigned long long cx[2] = {qx.x, qx.y}, cy[2] = {qy.x, qy.y}, cz[2] = {qz.x, qz.y};
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        unsigned long long dx, dy, dz, dd, ta, tb;
                        asm("sub.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(ui2), "l"(cx[e]));
                        asm("sub.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(vi2), "l"(cy[e]));
                        asm("sub.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(wi2), "l"(cz[e]));
                        asm("mul.f32x2 %0, %1, %1;" : "=l"(dd) : "l"(dx));
                        asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(dd) : "l"(dy), "l"(dd));
                        asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(dd) : "l"(dz), "l"(dd));
                        // sign(thr - dd) = 1  <=>  dd > thr  (certainly not a neighbour); a NaN distance gives the
                        // canonical positive NaN: it stays a candidate and the reference accepts it as well
                        asm("sub.f32x2 %0, %1, %2;" : "=l"(ta) : "l"(thr2), "l"(dd));
                        // sign(dd - thr_lo) = 1  <=>  dd < thr_lo: certainly a neighbour.  A NaN distance (a coordinate
                        // outside the range the rounding bound covers, k_prefilter_coords) has sign 0 in both tests: it is
                        // neither certainly outside nor certainly inside and goes to the exact predicate
                        asm("sub.f32x2 %0, %1, %2;" : "=l"(tb) : "l"(dd), "l"(thr_lo2));
                        unsigned a0, a1, c0, c1;
                        asm("mov.b64 {%0,%1}, %2;" : "=r"(a0), "=r"(a1) : "l"(ta));
                        asm("mov.b64 {%0,%1}, %2;" : "=r"(c0), "=r"(c1) : "l"(tb));
                        not_maybe = __funnelshift_l(a0, not_maybe, 1);
                        not_maybe = __funnelshift_l(a1, not_maybe, 1);
                        is_sure = __funnelshift_l(c0, is_sure, 1);
                        is_sure = __funnelshift_l(c1, is_sure, 1);
                    }
                }
                const int sh = 32 - 4 * nq;
                unsigned m = __brev(~not_maybe) >> sh;
                unsigned sure = __brev(is_sure) >> sh;
                // only slots of this row range count (the aligned chunk may start up to 3 slots early / end late)
                const int lo_bit = max(jb - j0, 0);
                unsigned valid = nj >= 32 ? 0xffffffffu : ((1u << nj) - 1u);
                valid &= ~((1u << lo_bit) - 1u);
                m &= valid;
                const unsigned self_off = (unsigned)(i - j0);  // p == q (core.jl:105): drop the own slot's bit
                if (self_off < 32u) m &= ~(1u << self_off);
                if (m2) drain();
                m2 = m1; s2 = s1; b2 = b1;
                m1 = m0; s1 = s0; b1 = b0;
                m0 = m;  s0 = sure; b0 = j0;
            }
        }
        drain();
    }
    cnt[i] = n_out;
    if (n_out > capk) atomicMax(max_cnt, n_out);  // rare: lets the host grow the lists for the next build
}

// Replay: one thread per target walks its list column (the k-th entries of a warp are one 128-byte line).  (Several
// lanes per target, combined by a butterfly, were measured in round 1: same L1 wavefronts, more requests — removed.)
template <class Op, int MINB = 6>
__global__ void __launch_bounds__(128, MINB) k_sweep_list(SpGrid g, SweepCtx c, const int* __restrict__ cnt,
                                                       const int* __restrict__ ids, typename Op::Params P, int self_flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (!sp_is_target(c, i)) return;
    if (!Op::active(P, i)) return;
    const double xi = c.x[i], yi = c.y[i], zi = c.z[i];
    typename Op::PS p;
    typename Op::Acc acc;
    Op::load(P, i, xi, yi, zi, p, acc);
    const int n_nb = cnt[i];
    if (n_nb <= c.capk) {
        const int* col = ids + ((size_t)(i >> 5) * c.capk << 5) + (i & 31);
        const int n_it = n_nb;
        // U pairs per trip: all their loads (ids first, then the gathers) are issued before the first pair body, so
        // each warp keeps U*(3+NQ) gathers in flight — the kernel is bound by the latency / L1 cost of these loads
        constexpr int U = Op::NQ <= 1 ? 4 : 2;
        int k = 0;
        // the ids of the next trip are requested one trip ahead (two dependent round trips per trip otherwise)
        int jn[U];
#pragma unroll
        for (int u = 0; u < U; u++) jn[u] = (u < n_it) ? __ldcs(col + (u << 5)) : i;
        for (; k + U <= n_it; k += U) {
            int jj[U];
#pragma unroll
            for (int u = 0; u < U; u++) jj[u] = jn[u];
#pragma unroll
            for (int u = 0; u < U; u++) jn[u] = (k + U + u < n_it) ? __ldcs(col + ((k + U + u) << 5)) : i;
            double qx[U], qy[U], qz[U];
            QRegs<Op::NQ> q[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                qx[u] = c.x[jj[u]];
                qy[u] = c.y[jj[u]];
                qz[u] = c.z[jj[u]];
#pragma unroll
                for (int e = 0; e < Op::NQ; e++) q[u].v[e] = P.qp[e][jj[u]];
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const double dx = __dsub_rn(xi, qx[u]), dy = __dsub_rn(yi, qy[u]), dz = __dsub_rn(zi, qz[u]);
                Op::pair(P, p, q[u], dx, dy, dz, sp_sqrt_fast(sp_d2(dx, dy, dz)), acc);
            }
        }
#pragma unroll
        for (int u = 0; u < U - 1; u++) {  // the tail (fewer than U entries) came with the last prefetch
            if (k + u >= n_it) break;
            const int j = jn[u];
            const double dx = __dsub_rn(xi, c.x[j]), dy = __dsub_rn(yi, c.y[j]), dz = __dsub_rn(zi, c.z[j]);
            QGlobal<Op::NQ> q{P.qp, j};
            Op::pair(P, p, q, dx, dy, dz, sp_sqrt_fast(sp_d2(dx, dy, dz)), acc);
        }
    } else {
        // more neighbours than the cache holds per target: the exact candidate scan, same visiting order
        sp_for_candidates<false>(g, c, xi, yi, zi, [&](int j, double dx, double dy, double dz, double d2) {
            if (d2 > g.T2 || j == i) return;
            QGlobal<Op::NQ> q{P.qp, j};
            Op::pair(P, p, q, dx, dy, dz, sp_sqrt_fast(d2), acc);
        });
    }
    if (self_flag & 1) Op::self(P, p, acc);
    Op::store(P, i, p, acc);
}

// ---- fused list build + first replay (default for the operators that declare FUSED_BUILD, i.e. the first pair sweep
// after create_cell_list! in the WCSPH / ISPH time loops).
// Phase A is the FP32 candidate scan of k_nbr_build with ONE threshold: every candidate that MAY be a neighbour
// (d2_f32 <= 1 + delta) is appended to the target's own column of the warp-tiled list.  Phase B replays that column
// exactly like k_sweep_list, and since the pair body needs the exact FP64 d2 anyway, the reference's predicate
// (r > h) && continue  <=>  d2 > T2  (core.jl:105) is decided THERE, for free: the few false "maybes" of the thin FP32
// shell are dropped and the column is compacted in place, so what stays in HBM is the exact neighbour list in visiting
// order — the same lists k_nbr_build writes, for the later sweeps of the step.  Against build + replay as two kernels
// this saves the second classification (2 of 7.25 instructions per candidate), the three-chunk reject queue, one
// 1.2 GB read of the lists from HBM and a launch; and the issue-bound scan of some warps overlaps with the
// L1-bound gathers of others on the same SM.
template <class Op, int MINB>
__global__ void __launch_bounds__(128, MINB) k_nbr_build_sweep(SpGrid g, SweepCtx c, int* cnt, int* ids, int* max_cnt,
                                                               typename Op::Params P, int self_flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *c.alive) return;  // the dead tail [alive, n) takes no part (sp_internal.cuh)
    if (!sp_is_target(c, i)) {  // outermost ghost layer of a slab: no list, no sweep
        cnt[i] = 0;
        return;
    }
    const int capk = c.capk;
    int* col = ids + ((size_t)(i >> 5) * capk << 5) + (i & 31);
    const double xi = c.x[i], yi = c.y[i], zi = c.z[i];
    int n_maybe = 0;
    {
        const float ui = c.ux[i], vi = c.uy[i], wi = c.uz[i];
        unsigned long long ui2, vi2, wi2, thr2;
        asm("mov.b64 %0, {%1,%1};" : "=l"(ui2) : "f"(ui));
        asm("mov.b64 %0, {%1,%1};" : "=l"(vi2) : "f"(vi));
        asm("mov.b64 %0, {%1,%1};" : "=l"(wi2) : "f"(wi));
        asm("mov.b64 %0, {%1,%1};" : "=l"(thr2) : "f"(c.thr));
        const long long key = sp_find_key(g, xi, yi, zi);  // core.jl:95
        const long long L1 = g.lim[0], L12 = g.lim[0] * g.lim[1];
        const int nk = (g.dim == 2) ? 0 : 1;
        for (int dk = -nk; dk <= nk; dk++) {
            for (int dj = -1; dj <= 1; dj++) {
                const long long mid = key + L1 * dj + L12 * dk;
                long long klo = mid - 1, khi = mid + 1;
                if (klo < 1) klo = 1;
                if (khi > g.key_max) khi = g.key_max;
                if (klo > khi) continue;
                const int jb = c.cell_start[klo], je = c.cell_start[khi + 1];
                for (int j0 = jb & ~3; j0 < je; j0 += 32) {
                    const int nj = min(32, je - j0);
                    const int nq = (nj + 3) >> 2;
                    unsigned not_maybe = 0u;  // candidate k of the chunk lands in bit 4*nq-1-k
                    const ulonglong2* px = reinterpret_cast<const ulonglong2*>(c.ux + j0);
                    const ulonglong2* py = reinterpret_cast<const ulonglong2*>(c.uy + j0);
                    const ulonglong2* pz = reinterpret_cast<const ulonglong2*>(c.uz + j0);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        if (q >= nq) break;
                        const ulonglong2 qx = __ldg(px + q), qy = __ldg(py + q), qz = __ldg(pz + q);
                        const unsigned long long cx[2] = {qx.x, qx.y}, cy[2] = {qy.x, qy.y}, cz[2] = {qz.x, qz.y};
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            unsigned long long dx, dy, dz, dd, ta;
                            asm("sub.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(ui2), "l"(cx[e]));
                            asm("sub.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(vi2), "l"(cy[e]));
                            asm("sub.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(wi2), "l"(cz[e]));
                            asm("mul.f32x2 %0, %1, %1;" : "=l"(dd) : "l"(dx));
                            asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(dd) : "l"(dy), "l"(dd));
                            asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(dd) : "l"(dz), "l"(dd));
                            // sign(thr - dd) = 1  <=>  dd > thr: certainly not a neighbour.  A NaN distance gives the
                            // canonical positive NaN: it stays a candidate, and the exact test accepts it as the
                            // reference does (NaN > h is false)
                            asm("sub.f32x2 %0, %1, %2;" : "=l"(ta) : "l"(thr2), "l"(dd));
                            unsigned a0, a1;
                            asm("mov.b64 {%0,%1}, %2;" : "=r"(a0), "=r"(a1) : "l"(ta));
                            not_maybe = __funnelshift_l(a0, not_maybe, 1);
                            not_maybe = __funnelshift_l(a1, not_maybe, 1);
                        }
                    }
                    unsigned m = __brev(~not_maybe) >> (32 - 4 * nq);
                    // only slots of this row range count (the aligned chunk may start up to 3 slots early / end late)
                    const int lo_bit = max(jb - j0, 0);
                    unsigned valid = nj >= 32 ? 0xffffffffu : ((1u << nj) - 1u);
                    valid &= ~((1u << lo_bit) - 1u);
                    m &= valid;
                    const unsigned self_off = (unsigned)(i - j0);  // p == q (core.jl:105): drop the own slot's bit
                    if (self_off < 32u) m &= ~(1u << self_off);
                    while (m) {
                        const int j = j0 + __ffs(m) - 1;
                        m &= m - 1;
                        if (n_maybe < capk) col[n_maybe << 5] = j;
                        n_maybe++;
                    }
                }
            }
        }
    }
    // ---- phase B: replay the own column with the exact predicate, compacting it in place
    const double T2 = g.T2;
    const bool act = Op::active(P, i);
    typename Op::PS p;
    typename Op::Acc acc;
    if (act) Op::load(P, i, xi, yi, zi, p, acc);
    int n_out = 0;
    if (n_maybe <= capk) {
        constexpr int U = Op::NQ <= 1 ? 4 : 2;
        int k = 0;
        // the ids of the NEXT trip are requested before this trip's gathers are consumed: they come back from L2 (phase A
        // wrote them through), and two dependent round trips per trip is what this loop would otherwise wait on
        int jn[U];
#pragma unroll
        for (int u = 0; u < U; u++) jn[u] = (u < n_maybe) ? col[u << 5] : i;
        for (; k + U <= n_maybe; k += U) {
            int jj[U];
#pragma unroll
            for (int u = 0; u < U; u++) jj[u] = jn[u];
#pragma unroll
            for (int u = 0; u < U; u++) jn[u] = (k + U + u < n_maybe) ? col[(k + U + u) << 5] : i;
            double qx[U], qy[U], qz[U];
            QRegs<Op::NQ> q[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                qx[u] = c.x[jj[u]];
                qy[u] = c.y[jj[u]];
                qz[u] = c.z[jj[u]];
#pragma unroll
                for (int e = 0; e < Op::NQ; e++) q[u].v[e] = P.qp[e][jj[u]];
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const double dx = __dsub_rn(xi, qx[u]), dy = __dsub_rn(yi, qy[u]), dz = __dsub_rn(zi, qz[u]);
                const double d2 = sp_d2(dx, dy, dz);
                if (d2 > T2) continue;  // (r > h) && continue, core.jl:105
                // compaction: position n_out <= k + u, and everything up to k + 2U - 1 is already in registers
                if (n_out != k + u) col[n_out << 5] = jj[u];
                n_out++;
                if (act) Op::pair(P, p, q[u], dx, dy, dz, sp_sqrt_fast(d2), acc);
            }
        }
        // the tail entries (fewer than U) were prefetched with the last full trip
#pragma unroll
        for (int u = 0; u < U - 1; u++) {
            if (k + u >= n_maybe) break;
            const int j = jn[u];
            const double dx = __dsub_rn(xi, c.x[j]), dy = __dsub_rn(yi, c.y[j]), dz = __dsub_rn(zi, c.z[j]);
            const double d2 = sp_d2(dx, dy, dz);
            if (d2 > T2) continue;
            if (n_out != k + u) col[n_out << 5] = j;
            n_out++;
            if (act) {
                QGlobal<Op::NQ> q{P.qp, j};
                Op::pair(P, p, q, dx, dy, dz, sp_sqrt_fast(d2), acc);
            }
        }
    } else {
        // more candidates than the lists hold per target: the exact candidate scan in the same visiting order; the first
        // capk accepted slots are still recorded, so a target whose TRUE count fits is replayed from its list later
        sp_for_candidates<false>(g, c, xi, yi, zi, [&](int j, double dx, double dy, double dz, double d2) {
            if (d2 > T2 || j == i) return;
            if (n_out < capk) col[n_out << 5] = j;
            n_out++;
            if (act) {
                QGlobal<Op::NQ> q{P.qp, j};
                Op::pair(P, p, q, dx, dy, dz, sp_sqrt_fast(d2), acc);
            }
        });
    }
    cnt[i] = n_out;
    if (n_out > capk) atomicMax(max_cnt, n_out);  // rare: lets the host grow the lists for the next build
    if (!act) return;
    if (self_flag & 1) Op::self(P, p, acc);
    Op::store(P, i, p, acc);
}

template <class U>
__global__ void __launch_bounds__(256) k_unary(typename U::Params P, const int* __restrict__ alive) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *alive) U::apply(P, i);
}

#define TILE_TP 128
#define TILE_LCAP 40
#define TILE_RPB 3                     /* stencil rows per batch (3-D: 3 batches of 3 rows; 2-D: 1 row each) */
#define TILE_CAPB3 (3 * (TILE_TP + 24)) /* slots per buffer, 3-D */
#define TILE_CAPB2 (TILE_TP + 96)       /* slots per buffer, 2-D (rows are ~3x longer per cell) */

// FP32 cell-unit coordinates u = (x - lo)/h of every slot, the input of the tile kernel's pre-filter.
// The rounding bound delta of the pre-filter holds for |u| <= U (sp_ensure_prefilter).  A particle that has drifted
// further out since the last create_cell_list! (several move! calls without a rebuild, a runaway particle) gets NaN
// coordinates instead: a NaN distance is never "certainly outside" nor "certainly inside", so every pair it takes part
// in is decided by the exact FP64 predicate from its true position, as the reference decides all of them.
__global__ void __launch_bounds__(256) k_prefilter_coords(SpGrid g, const double* __restrict__ x, long long cap,
                                                          float* __restrict__ u, long long n, float U) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ih = 1.0 / g.h;
    float a = (float)((x[i] - g.lo[0]) * ih);
    float b = (float)((x[cap + i] - g.lo[1]) * ih);
    float c = (float)((x[2 * cap + i] - g.lo[2]) * ih);
    if (!(fabsf(a) <= U) || !(fabsf(b) <= U) || !(fabsf(c) <= U)) a = b = c = nanf("");
    u[i] = a;
    u[cap + i] = b;
    u[2 * cap + i] = c;
}

template <class Op, int CAPB, int NROWS, int RPB>
static int launch_tile(sp_system* s, const SweepCtx& c, const typename Op::Params& P, int self_flag) {
    const size_t buf = (size_t)CAPB * ((3 + Op::NQ) * sizeof(double) + 3 * sizeof(float));
    const size_t smem = 2 * buf + (size_t)TILE_LCAP * TILE_TP * sizeof(unsigned short);
    constexpr int MINB = 3;
    auto kern = k_sweep_tile<Op, TILE_TP, TILE_LCAP, CAPB, NROWS, RPB, MINB>;
    static bool configured = false;  // per instantiation
    if (!configured) {
        SP_CUDA(s, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    kern<<<sp_blocks(s->n, TILE_TP), TILE_TP, smem, s->stream>>>(s->g, c, P, self_flag);
    s->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sp_fail_cuda(s, e, "k_sweep_tile", __FILE__, __LINE__);
    return SP_OK;
}

// the balance_of_mass -> internal_force cache (OpBalanceOfMassAux) is used on the default path only
static bool sp_pair_aux_enabled(int flags) {
    static const bool on = !(getenv("SP_PAIR_AUX") && atoi(getenv("SP_PAIR_AUX")) == 0);
    return on && !(flags & (SP_FLAG_STRICT_ORDER | SP_FLAG_TILE_KERNEL));
}
// SP_FUSED_BUILD=0 builds the lists with k_nbr_build and replays them with k_sweep_list even for the operators that
// have the fused kernel (A/B timing and the parity tests of the two-kernel path)
static bool sp_fused_build_enabled() {
    static const bool on = !(getenv("SP_FUSED_BUILD") && atoi(getenv("SP_FUSED_BUILD")) == 0);
    return on;
}

static void sp_sweep_ctx(sp_system* s, SweepCtx& c) {
    const double* X = s->fields[0].d;
    c.x = X;
    c.y = X + s->cap;
    c.z = X + 2 * s->cap;
    c.ux = c.uy = c.uz = nullptr;
    c.cell_start = s->cell_start;
    c.thr = c.thr_lo = 0.f;
    c.capk = s->nbr_capk;
    c.n = (int)s->n;
    c.alive = sp_alive(s);
    c.tgt_key_lo = c.tgt_key_hi = 0;
    if (s->g.slab_axis >= 0 && s->have_cells) {
        const long long L = s->g.lim[0] * (s->g.slab_axis == 2 ? s->g.lim[1] : 1);  // cells per layer
        const long long nl = s->g.lim[s->g.slab_axis];
        c.tgt_key_lo = (int)(1 * L + 1);
        c.tgt_key_hi = (int)((nl - 1) * L + 1);
    }
}

// FP32 pre-filter inputs: coordinates in cell units u = (x - lo)/h and thresholds that can never misclassify.
// |u| <= U; rounding x -> u costs <= U*2^-24 per coordinate, the FP32 difference and the three products/sums a
// few 2^-24 relative more:   |d2_f32 - d2/h^2| <= delta = 8*2^-23*U + 1e-6   for d2 <~ h^2.
//   d2_f32 > 1 + delta   certainly not a neighbour        (thr)
//   d2_f32 <= 1 - delta  certainly a neighbour            (thr_lo, used by k_nbr_build only)
// Everything in between is decided by the exact FP64 predicate, so the neighbour set is the reference's bit for bit.
// U bounds |u| from the GLOBAL box (on a slab system key_lim is only the local window along the slab axis)
float sp_prefilter_range(const sp_system* s) {
    double U = (double)std::max(std::max(s->g.lim[0], s->g.lim[1]), s->g.lim[2]);
    for (int a = 0; a < 3; a++) U = std::max(U, std::ceil((s->g.hi[a] - s->g.lo[a]) / s->g.h) + 1.0);
    return (float)(U + 2.0);
}
static int sp_ensure_prefilter(sp_system* s, SweepCtx& c) {
    if (!s->ucoord || s->ucoord_cap != s->cap) {
        if (s->ucoord) SP_CUDA(s, sp_dfree(s, s->ucoord));
        s->ucoord = nullptr;
        // + 64: the pre-filter reads whole aligned groups of 4 (pairs) of candidates, up to 3 slots past slot n-1
        SP_CUDA(s, sp_dmalloc(&s->ucoord, ((size_t)3 * s->cap + 64) * sizeof(float)));
        s->ucoord_cap = s->cap;
        s->ucoord_version = 0;
    }
    const double U = (double)sp_prefilter_range(s);
    if (s->ucoord_version != s->x_version) {  // positions unchanged since the last sweep: keep the planes
        SP_LAUNCH(s, k_prefilter_coords, sp_blocks(s->n, 256), 256, 0, s->g, s->fields[0].d, s->cap, s->ucoord, s->n, (float)U);
        s->ucoord_version = s->x_version;
    }
    c.ux = s->ucoord;
    c.uy = s->ucoord + s->cap;
    c.uz = s->ucoord + 2 * s->cap;
    c.thr = (float)(1.0 + 8.0 * U / 8388608.0 + 1e-6);
    c.thr = std::nextafter(c.thr, 2.0f);
    // 1e-6 more for the FP64 rounding of d2 and of T2 itself
    c.thr_lo = (float)(1.0 - 8.0 * U / 8388608.0 - 2e-6);
    c.thr_lo = std::nextafter(c.thr_lo, -1.0f);
    return SP_OK;
}

// Make room for the cached neighbour lists and tell whether they have to be (re)built: positions / slot order changed
// since they were built.  The caller launches the build (k_nbr_build, or the fused k_nbr_build_sweep) and then calls
// sp_nbr_built.
static int sp_nbr_prepare(sp_system* s, SweepCtx& c, bool* need_build) {
    int rc = sp_ensure_prefilter(s, c);
    if (rc) return rc;
    const bool rebuild = s->nbr_version != s->x_version || s->nbr_n != s->n || !s->nbr_ids || s->nbr_cap != s->cap;
    if (rebuild && !s->capturing && s->nbr_max_pending && cudaEventQuery(s->ev_nbr) == cudaSuccess) {
        // the previous build met targets with more neighbours than the lists hold (they were swept by the exact scan):
        // give the lists room before building them again
        s->nbr_max_pending = false;
        const int mx = s->h_counters[40];
        if (mx > s->nbr_capk && s->nbr_capk < SP_NBR_CAPK_MAX) {
            int want = std::min(SP_NBR_CAPK_MAX, ((mx + mx / 4) + 31) & ~31);
            size_t free_b = 0, total_b = 0;
            SP_CUDA(s, cudaMemGetInfo(&free_b, &total_b));
            if ((size_t)s->cap * want * sizeof(int) < free_b / 4) {
                s->nbr_capk = want;
                if (s->nbr_ids) SP_CUDA(s, sp_dfree(s, s->nbr_ids));
                s->nbr_ids = nullptr;
            }
        }
    }
    if (!s->nbr_ids || s->nbr_cap != s->cap) {
        if (s->nbr_ids) SP_CUDA(s, sp_dfree(s, s->nbr_ids));
        if (s->nbr_cnt) SP_CUDA(s, sp_dfree(s, s->nbr_cnt));
        s->nbr_ids = s->nbr_cnt = nullptr;
        SP_CUDA(s, sp_dmalloc(&s->nbr_ids, (size_t)s->cap * s->nbr_capk * sizeof(int)));
        SP_CUDA(s, sp_dmalloc(&s->nbr_cnt, (size_t)s->cap * sizeof(int)));
        s->nbr_cap = s->cap;
        s->nbr_version = 0;
    }
    c.capk = s->nbr_capk;
    *need_build = s->nbr_version != s->x_version || s->nbr_n != s->n;
    // (counters[SP_CNT_NBRMAX] is a running maximum over all builds: the lists only ever grow)
    return SP_OK;
}
static int sp_nbr_built(sp_system* s) {
    // the longest list of this build travels to the host asynchronously; it is looked at before the NEXT build
    if (!s->capturing) {
        SP_CUDA(s, cudaMemcpyAsync(s->h_counters + 40, s->counters + 40, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        SP_CUDA(s, cudaEventRecord(s->ev_nbr, s->stream));
        s->nbr_max_pending = true;
    }
    s->nbr_version = s->x_version;
    s->nbr_n = s->n;
    return SP_OK;
}
static int sp_ensure_nbr_cache(sp_system* s, SweepCtx& c) {
    bool need = false;
    int rc = sp_nbr_prepare(s, c, &need);
    if (rc || !need) return rc;
    SP_LAUNCH(s, k_nbr_build, sp_blocks(s->n, 128), 128, 0, s->g, c, s->nbr_cnt, s->nbr_ids, s->counters + 40);
    return sp_nbr_built(s);
}

// Operators that declare `static constexpr bool LISTS_ONLY = true` (the example zoo beyond the BASELINE configs) are
// built for the two production kernels only — the cached-list replay and the strict-order scan.  The experimental
// kernels kept from the exploration (tile, packed records, hit masks; profiles/r1_sweep_exploration.md) are
// instantiated for the WCSPH / ISPH / collision operators they were measured and parity-tested with: ~10 kernels
// x 3 kernel families per operator is what the build time of this file is made of.
template <class Op, class = void>
struct SpListsOnly : std::false_type {};
template <class Op>
struct SpListsOnly<Op, std::void_t<decltype(Op::LISTS_ONLY)>> : std::true_type {};

template <class Op, class = void>
struct SpFusedBuild : std::false_type {};
template <class Op>
struct SpFusedBuild<Op, std::void_t<decltype(Op::FUSED_BUILD)>> : std::true_type {};

template <class Op>
static int launch_sweep(sp_system* s, const typename Op::Params& P, int flags) {
    if (s->n == 0) return SP_OK;
    SweepCtx c;
    sp_sweep_ctx(s, c);
    const int self_flag = (flags & SP_FLAG_SELF) ? 1 : 0;
    if (flags & SP_FLAG_STRICT_ORDER) {
        SP_LAUNCH(s, (k_sweep<Op, true>), sp_blocks(s->n, 128), 128, 0, s->g, c, P, self_flag);
        return SP_OK;
    }
    if (flags & SP_FLAG_TILE_KERNEL) {
        // the shared-memory tile kernel (TMA-staged candidate rows), kept as the one alternative to the list path
        if constexpr (SpListsOnly<Op>::value) {
            return sp_fail(s, SP_ERR_INVALID,
                           "this operator is built for the cached-list and strict-order kernels only (no SP_FLAG_TILE_KERNEL)");
        } else {
            int rc = sp_ensure_prefilter(s, c);
            if (rc) return rc;
            if (s->g.dim == 2) return launch_tile<Op, TILE_CAPB2, 3, 1>(s, c, P, self_flag);
            return launch_tile<Op, TILE_CAPB3, 9, TILE_RPB>(s, c, P, self_flag);
        }
    }
    // default: cached neighbour lists
    bool need = false;
    int rc = sp_nbr_prepare(s, c, &need);
    if (rc) return rc;
    const unsigned nb = sp_blocks(s->n, 128);
    if (need) {
        if constexpr (SpFusedBuild<Op>::value) {
            if (sp_fused_build_enabled() && !(flags & SP_FLAG_UNFUSED_BUILD)) {
                // 80 registers / 6 CTAs per SM; 72 and 64 registers spill and were measured slower (3.23 / 3.79 vs 3.00 ms)
                SP_LAUNCH(s, (k_nbr_build_sweep<Op, 6>), nb, 128, 0, s->g, c, s->nbr_cnt, s->nbr_ids, s->counters + 40, P, self_flag);
                return sp_nbr_built(s);
            }
        }
        SP_LAUNCH(s, k_nbr_build, nb, 128, 0, s->g, c, s->nbr_cnt, s->nbr_ids, s->counters + 40);
        if ((rc = sp_nbr_built(s))) return rc;
    }
    if constexpr (Op::NQ <= 1) {
        // operators that gather one plane besides the position (the cached force sweep): 64 registers, 8 CTAs per SM —
        // measured 1.21 -> 1.06 ms on the 10 M dam break (7 CTAs: 1.11 ms); SP_LIST_MINB=6 for the A/B comparison
        static const int minb = getenv("SP_LIST_MINB") ? atoi(getenv("SP_LIST_MINB")) : 8;
        if (minb == 8) {
            SP_LAUNCH(s, (k_sweep_list<Op, 8>), nb, 128, 0, s->g, c, s->nbr_cnt, s->nbr_ids, P, self_flag);
            return SP_OK;
        }
    }
    SP_LAUNCH(s, (k_sweep_list<Op>), nb, 128, 0, s->g, c, s->nbr_cnt, s->nbr_ids, P, self_flag);
    return SP_OK;
}

template <template <class> class OpT, class MakeParams>
static int dispatch_kernel(sp_system* s, int kernel, double h, int flags, MakeParams&& mk) {
    SpKC kc;
    if (!sp_make_kc(kernel, h, &kc)) return sp_fail(s, SP_ERR_INVALID, "unknown SPH kernel id");
    switch (kernel) {
        case SP_KERNEL_WENDLAND1:
        case SP_KERNEL_WENDLAND2:
        case SP_KERNEL_WENDLAND3: {
            typename OpT<KWendland>::Params P;
            mk(P);
            P.kc = kc;
            return launch_sweep<OpT<KWendland>>(s, P, flags);
        }
        case SP_KERNEL_SPLINE23: {
            typename OpT<KSpline23>::Params P;
            mk(P);
            P.kc = kc;
            return launch_sweep<OpT<KSpline23>>(s, P, flags);
        }
        case SP_KERNEL_SPLINE24: {
            typename OpT<KSpline24>::Params P;
            mk(P);
            P.kc = kc;
            return launch_sweep<OpT<KSpline24>>(s, P, flags);
        }
    }
    return sp_fail(s, SP_ERR_INVALID, "unknown SPH kernel id");
}

template <class U>
static int launch_unary(sp_system* s, const typename U::Params& P) {
    if (s->n == 0) return SP_OK;
    SP_LAUNCH(s, (k_unary<U>), sp_blocks(s->n, 256), 256, 0, P, sp_alive(s));
    return SP_OK;
}

static inline RV3 rv3(sp_system* s, int fid) {
    const double* d = s->fields[fid].d;
    return RV3{d, d + s->cap, d + 2 * s->cap};
}
static inline WV3 wv3(sp_system* s, int fid) {
    double* d = s->fields[fid].d;
    return WV3{d, d + s->cap, d + 2 * s->cap};
}
static inline double* sc(sp_system* s, int fid) { return s->fields[fid].d; }
static inline void set_v3(sp_system* s, int fid, const double** qp) {
    const double* d = s->fields[fid].d;
    qp[0] = d;
    qp[1] = d + s->cap;
    qp[2] = d + 2 * s->cap;
}

template <class U>
static int launch_unary(sp_system* s, const typename U::Params& P);

// pr = P/rho^2 for every particle into the transient scratch field "_pr" (one IEEE division per particle,
// the same value the closure computes per pair)
static int pressure_over_rho2(sp_system* s, int fP, int frho, double** out) {
    int32_t fid;
    int rc = sp_add_field(s, "_pr", 1, &fid);
    if (rc) return rc;
    s->fields[fid].transient = true;
    UPressureOverRho2::Params P{s->fields[fP].d, s->fields[frho].d, s->fields[fid].d};
    if ((rc = launch_unary<UPressureOverRho2>(s, P))) return rc;
    *out = s->fields[fid].d;
    return SP_OK;
}

// Internal entry shared with the step programs (sp_program.cu) and the CG (sp_isph.cu).
int sp_apply_impl(sp_system* s, int32_t op, const int32_t* F, int32_t nf, const double* Pm, int32_t np, int32_t flags) {
#define NEED(NC, NP, ...)                                                                    \
    {                                                                                        \
        const int _nc[] = {__VA_ARGS__};                                                     \
        int _rc = sp_check_fields(s, F, nf, _nc, NC);                                        \
        if (_rc) return _rc;                                                                 \
        if (np != (NP) || !Pm) return sp_fail(s, SP_ERR_INVALID, "wrong number of parameters for this operator"); \
    }
#define NEED_CELLS()                                                                                          \
    {                                                                                                         \
        if (!s->have_cells) return sp_fail(s, SP_ERR_STATE, "binary operator before sp_create_cell_list");    \
        if (F[0] != 0) return sp_fail(s, SP_ERR_INVALID, "the first field of a binary operator must be x (field 0)"); \
    }
    switch (op) {
        case SP_OP_BALANCE_OF_MASS: {
            NEED(4, 4, 3, 3, 1, 1);
            NEED_CELLS();
            sp_wrote(s, F[3]);
            s->pair_aux.valid = false;
            if (sp_pair_aux_enabled(flags)) {
                int32_t fkx, fkv;
                int rc2 = sp_add_field(s, "_kx", 3, &fkx);
                if (!rc2) rc2 = sp_add_field(s, "_kv", 3, &fkv);
                if (rc2) return rc2;
                s->fields[fkx].transient = s->fields[fkv].transient = true;
                rc2 = dispatch_kernel<OpBalanceOfMassAux>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                    set_v3(s, F[1], P.qp);
                    P.qp[3] = sc(s, F[2]);
                    P.Drho = sc(s, F[3]);
                    P.kx = wv3(s, fkx);
                    P.kv = wv3(s, fkv);
                    P.m = Pm[1];
                    P.two_nu = Pm[3];
                });
                if (rc2) return rc2;
                auto& t = s->pair_aux;
                t.valid = true;
                t.x_version = s->x_version;
                t.v_fid = F[1];
                t.v_version = s->fields[F[1]].version;
                t.n = s->n;
                t.kernel = (int)Pm[0];
                t.m = Pm[1];
                t.h = Pm[2];
                t.f_kx = fkx;
                t.f_kv = fkv;
                return SP_OK;
            }
            return dispatch_kernel<OpBalanceOfMass>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                set_v3(s, F[1], P.qp);
                P.qp[3] = sc(s, F[2]);
                P.Drho = sc(s, F[3]);
                P.m = Pm[1];
                P.two_nu = Pm[3];
            });
        }
        case SP_OP_FIND_PRESSURE: {
            NEED(3, 4, 1, 1, 1);
            sp_wrote(s, F[0]);
            sp_zeroed(s, F[1]);
            sp_wrote(s, F[2]);
            UFindPressure::Params P{sc(s, F[0]), sc(s, F[1]), sc(s, F[2]), Pm[0], Pm[1], Pm[2], Pm[3]};
            return launch_unary<UFindPressure>(s, P);
        }
        case SP_OP_INTERNAL_FORCE: {
            NEED(6, 5, 3, 3, 1, 1, 3, 1);
            NEED_CELLS();
            sp_wrote(s, F[4]);
            double* pr = nullptr;
            if (flags & SP_FLAG_INTERNAL_PR_READY) {  // the step program's fused find_pressure pass has written _pr
                int32_t fid;
                if (sp_find_field(s, "_pr", &fid)) return sp_fail(s, SP_ERR_STATE, "_pr missing");
                pr = s->fields[fid].d;
                flags &= ~SP_FLAG_INTERNAL_PR_READY;
            } else {
                int rc2 = pressure_over_rho2(s, F[2], F[3], &pr);
                if (rc2) return rc2;
            }
            {
                const auto& t = s->pair_aux;
                if (t.valid && sp_pair_aux_enabled(flags) && t.x_version == s->x_version && t.v_fid == F[1] &&
                    t.v_version == s->fields[F[1]].version && t.n == s->n && t.kernel == (int)Pm[0] && t.m == Pm[1] &&
                    t.h == Pm[2]) {
                    return dispatch_kernel<OpInternalForceCached>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                        P.qp[0] = pr;
                        P.Dv = wv3(s, F[4]);
                        P.type = sc(s, F[5]);
                        P.kx = rv3(s, t.f_kx);
                        P.kv = rv3(s, t.f_kv);
                        P.m = Pm[1];
                        P.visc = 2 * Pm[3] / (Pm[4] * Pm[4]);
                    });
                }
            }
            return dispatch_kernel<OpInternalForce>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                set_v3(s, F[1], P.qp);
                P.qp[3] = pr;
                P.Dv = wv3(s, F[4]);
                P.type = sc(s, F[5]);
                P.m = Pm[1];
                P.visc = 2 * Pm[3] / (Pm[4] * Pm[4]);
            });
        }
        case SP_OP_INTERNAL_FORCE_CAVITY: {
            NEED(6, 6, 3, 3, 1, 1, 3, 1);
            NEED_CELLS();
            sp_wrote(s, F[4]);
            double* pr = nullptr;
            {
                int rc2 = pressure_over_rho2(s, F[2], F[3], &pr);
                if (rc2) return rc2;
            }
            return dispatch_kernel<OpInternalForceCavity>(s, SP_KERNEL_WENDLAND2, Pm[1], flags, [&](auto& P) {
                set_v3(s, F[1], P.qp);
                P.qp[3] = pr;
                P.qp[4] = sc(s, F[3]);
                P.qp[5] = sc(s, F[5]);
                P.Dv = wv3(s, F[4]);
                P.m = Pm[0];
                P.Re = Pm[2];
                P.vlid = Pm[3];
                P.ylid = Pm[4];
                P.lid = Pm[5];
                P.tenth_h = 0.1 * Pm[1];
                P.eps = 0.01 * (Pm[1] * Pm[1]);
            });
        }
        case SP_OP_MOVE: {
            NEED(4, 1, 3, 3, 3, 1);
            sp_wrote(s, F[0]);
            sp_zeroed(s, F[2]);
            UMove::Params P{wv3(s, F[0]), rv3(s, F[1]), wv3(s, F[2]), sc(s, F[3]), Pm[0]};
            return launch_unary<UMove>(s, P);
        }
        case SP_OP_ACCELERATE: {
            NEED(3, 4, 3, 3, 1);
            sp_wrote(s, F[0]);
            UAccelerate::Params P{wv3(s, F[0]), rv3(s, F[1]), sc(s, F[2]), Pm[0], Pm[1], Pm[2], Pm[3]};
            return launch_unary<UAccelerate>(s, P);
        }
        case SP_OP_DENSITY_SUM: {
            NEED(2, 3, 3, 1);
            NEED_CELLS();
            sp_wrote(s, F[1]);
            return dispatch_kernel<OpDensitySum>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                P.qp[0] = nullptr;
                P.out = sc(s, F[1]);
                P.m = Pm[1];
            });
        }
        case SP_OP_PRESSURE_FROM_RHO: {
            NEED(3, 1, 1, 1, 1);
            sp_wrote(s, F[0]);
            sp_wrote(s, F[1]);
            sp_wrote(s, F[2]);
            UPressureFromRho::Params P{sc(s, F[0]), sc(s, F[1]), sc(s, F[2]), Pm[0]};
            return launch_unary<UPressureFromRho>(s, P);
        }
        case SP_OP_INTERNAL_FORCE_SYM: {
            NEED(3, 4, 3, 1, 3);
            NEED_CELLS();
            sp_wrote(s, F[2]);
            return dispatch_kernel<OpInternalForceSym>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                P.qp[0] = sc(s, F[1]);
                P.a = wv3(s, F[2]);
                P.m = Pm[1];
                P.inv_rho0sq = 1.0 / (Pm[3] * Pm[3]);
            });
        }
        case SP_OP_FILL: {
            NEED(1, 1, 0);
            if (Pm[0] == 0.0 && !std::signbit(Pm[0])) sp_zeroed(s, F[0]); else sp_wrote(s, F[0]);
            UFill::Params P{sc(s, F[0]), s->cap, s->fields[F[0]].ncomp, Pm[0]};
            return launch_unary<UFill>(s, P);
        }
        case SP_OP_ADVECT: {
            NEED(2, 1, 3, 3);
            sp_wrote(s, F[0]);
            UAdvect::Params P{wv3(s, F[0]), rv3(s, F[1]), Pm[0]};
            return launch_unary<UAdvect>(s, P);
        }
        case SP_OP_KICK: {
            NEED(2, 1, 3, 3);
            sp_wrote(s, F[0]);
            UKick::Params P{wv3(s, F[0]), rv3(s, F[1]), Pm[0]};
            return launch_unary<UKick>(s, P);
        }
        case SP_OP_ISPH_INITIALIZE: {
            NEED(6, 4, 3, 3, 1, 1, 1, 1);
            for (int k = 0; k < 6; k++) sp_wrote(s, F[k]);
            UIsphInitialize::Params P{wv3(s, F[0]), wv3(s, F[1]), sc(s, F[2]), sc(s, F[3]), sc(s, F[4]),
                                      sc(s, F[5]),  Pm[0],        Pm[1],       Pm[2],       Pm[3]};
            return launch_unary<UIsphInitialize>(s, P);
        }
        case SP_OP_ISPH_VISCOUS_FORCE: {
            NEED(3, 5, 3, 3, 3);
            NEED_CELLS();
            sp_wrote(s, F[2]);
            return dispatch_kernel<OpIsphViscous>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                set_v3(s, F[1], P.qp);
                P.Dv = wv3(s, F[2]);
                P.coef = 2.0 * Pm[1] * Pm[3] / (Pm[4] * Pm[4]);
            });
        }
        case SP_OP_ISPH_DIV_L_LAMBDA: {
            NEED(5, 5, 3, 3, 1, 1, 1);
            NEED_CELLS();
            sp_wrote(s, F[2]);
            sp_wrote(s, F[3]);
            sp_wrote(s, F[4]);
            return dispatch_kernel<OpIsphDivLLambda>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                set_v3(s, F[1], P.qp);
                P.div = sc(s, F[2]);
                P.L = sc(s, F[3]);
                P.lambda = sc(s, F[4]);
                P.m = Pm[1];
                P.m_over_rho = Pm[1] / Pm[3];
                P.inv_dim = 1.0 / Pm[4];
            });
        }
        case SP_OP_ISPH_PROJECTION_VECTOR: {
            NEED(2, 2, 1, 1);
            sp_wrote(s, F[0]);
            sp_wrote(s, F[1]);
            UIsphProjectionVector::Params P{sc(s, F[0]), sc(s, F[1]), -(Pm[0] * Pm[0]), Pm[1]};
            return launch_unary<UIsphProjectionVector>(s, P);
        }
        case SP_OP_ISPH_INTERNAL_FORCE: {
            NEED(3, 4, 3, 1, 3);
            NEED_CELLS();
            sp_wrote(s, F[2]);
            return dispatch_kernel<OpIsphInternalForce>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                P.qp[0] = sc(s, F[1]);
                P.Dv = wv3(s, F[2]);
                P.coef = Pm[1] / (Pm[3] * Pm[3]);
            });
        }
        case SP_OP_ISPH_ACCELERATE: {
            NEED(3, 1, 3, 3, 1);
            sp_wrote(s, F[0]);
            sp_zeroed(s, F[1]);
            UIsphAccelerate::Params P{wv3(s, F[0]), wv3(s, F[1]), sc(s, F[2]), Pm[0]};
            return launch_unary<UIsphAccelerate>(s, P);
        }
        case SP_OP_SC_BALANCE_OF_MASS: {
            NEED(3, 4, 3, 3, 1);
            NEED_CELLS();
            sp_wrote(s, F[2]);
            return dispatch_kernel<OpScBalanceOfMass>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                set_v3(s, F[1], P.qp);
                P.rho = sc(s, F[2]);
                P.m = Pm[1];
                P.dt = Pm[3];
            });
        }
        case SP_OP_SC_INTERNAL_FORCE: {
            NEED(5, 6, 3, 3, 1, 3, 1);
            NEED_CELLS();
            sp_wrote(s, F[3]);
            int32_t fpr;
            {
                int rc2 = sp_add_field(s, "_pr", 1, &fpr);
                if (rc2) return rc2;
                s->fields[fpr].transient = true;
                UEosPressureOverRho2::Params Pp{sc(s, F[2]), s->fields[fpr].d, Pm[4], Pm[5]};
                if ((rc2 = launch_unary<UEosPressureOverRho2>(s, Pp))) return rc2;
            }
            return dispatch_kernel<OpScInternalForce>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                set_v3(s, F[1], P.qp);
                P.qp[3] = s->fields[fpr].d;
                P.qp[4] = sc(s, F[2]);
                P.a = wv3(s, F[3]);
                P.type = sc(s, F[4]);
                P.m = Pm[1];
                P.two_mu = 2 * Pm[3];
            });
        }
        case SP_OP_FIND_NORMAL: {
            NEED(2, 3, 3, 3);
            NEED_CELLS();
            sp_wrote(s, F[1]);
            return dispatch_kernel<OpFindNormal>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                P.qp[0] = nullptr;
                P.n = wv3(s, F[1]);
                P.coef = Pm[1];
            });
        }
        case SP_OP_NORMALIZE: {
            NEED(1, 1, 3);
            sp_wrote(s, F[0]);
            UNormalize::Params P{wv3(s, F[0]), Pm[0]};
            return launch_unary<UNormalize>(s, P);
        }
        case SP_OP_INTERNAL_FORCE_TENSION: {
            NEED(5, 6, 3, 3, 1, 3, 3);
            NEED_CELLS();
            sp_wrote(s, F[4]);
            return dispatch_kernel<OpInternalForceTension>(s, SP_KERNEL_WENDLAND3, Pm[1], flags, [&](auto& P) {
                set_v3(s, F[1], P.qp);
                P.qp[3] = sc(s, F[2]);
                set_v3(s, F[3], P.qp + 4);
                P.a = wv3(s, F[4]);
                P.m = Pm[0];
                P.mu = Pm[2];
                P.rho0sq = Pm[3] * Pm[3];
                P.tens = 2 * Pm[4] / (Pm[3] * Pm[3]);
                P.s0 = Pm[5];
            });
        }
        case SP_OP_MOVE_ALL: {
            NEED(3, 1, 3, 3, 3);
            sp_wrote(s, F[0]);
            sp_zeroed(s, F[2]);
            UMoveAll::Params P{wv3(s, F[0]), rv3(s, F[1]), wv3(s, F[2]), Pm[0]};
            return launch_unary<UMoveAll>(s, P);
        }
        case SP_OP_DENSITY_SUM_FLUID: {
            NEED(3, 3, 3, 1, 1);
            NEED_CELLS();
            sp_wrote(s, F[1]);
            return dispatch_kernel<OpDensitySumFluid>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                P.qp[0] = sc(s, F[2]);
                P.out = sc(s, F[1]);
                P.m = Pm[1];
            });
        }
        case SP_OP_INTERNAL_FORCE_LJ: {
            NEED(5, 8, 3, 1, 1, 3, 1);
            NEED_CELLS();
            sp_wrote(s, F[3]);
            double* pr = nullptr;
            if (Pm[3] == 0.0) {  // collapse_symplectic.jl:116: P/rho^2 of each particle
                int rc2 = pressure_over_rho2(s, F[1], F[2], &pr);
                if (rc2) return rc2;
            } else {  // Kepler_vortex.jl:158: P/rho0^2 with the constant rho0
                int32_t fpr;
                int rc2 = sp_add_field(s, "_pr", 1, &fpr);
                if (rc2) return rc2;
                s->fields[fpr].transient = true;
                pr = s->fields[fpr].d;
                UPressureOverConst::Params Pp{sc(s, F[1]), pr, Pm[3] * Pm[3]};
                if ((rc2 = launch_unary<UPressureOverConst>(s, Pp))) return rc2;
            }
            return dispatch_kernel<OpInternalForceLJ>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                P.qp[0] = pr;
                P.qp[1] = sc(s, F[4]);
                P.a = wv3(s, F[3]);
                P.m = Pm[1];
                P.wall = Pm[4];
                P.dr_wall = Pm[5];
                P.E_wall = Pm[6];
                P.eps = Pm[7];
            });
        }
        case SP_OP_MOVE_REV: {
            NEED(3, 1, 3, 3, 1);
            sp_wrote(s, F[0]);
            UMoveRev::Params P{wv3(s, F[0]), rv3(s, F[1]), sc(s, F[2]), Pm[0]};
            return launch_unary<UMoveRev>(s, P);
        }
        case SP_OP_ACCELERATE_REV: {
            NEED(3, 4, 3, 3, 1);
            sp_wrote(s, F[0]);
            UAccelerateRev::Params P{wv3(s, F[0]), rv3(s, F[1]), sc(s, F[2]), Pm[0], Pm[1], Pm[2], Pm[3]};
            return launch_unary<UAccelerateRev>(s, P);
        }
        case SP_OP_ACCELERATE_REV_CENTRAL: {
            NEED(4, 2, 3, 3, 3, 1);
            sp_wrote(s, F[1]);
            UAccelerateRevCentral::Params P{rv3(s, F[0]), wv3(s, F[1]), rv3(s, F[2]), sc(s, F[3]), Pm[0], Pm[1]};
            return launch_unary<UAccelerateRevCentral>(s, P);
        }
        case SP_OP_LJ_POTENTIAL: {
            NEED(3, 5, 3, 1, 1);
            NEED_CELLS();
            sp_wrote(s, F[1]);
            // no SPH kernel in this sum: any family serves the dispatch, only the cut-off h matters
            return dispatch_kernel<OpLJPotential>(s, SP_KERNEL_WENDLAND2, Pm[0], flags & ~SP_FLAG_SELF, [&](auto& P) {
                P.qp[0] = sc(s, F[2]);
                P.out = sc(s, F[1]);
                P.coef = Pm[1];
                P.wall = Pm[2];
                P.dr_wall = Pm[3];
                P.eps = Pm[4];
            });
        }
        case SP_OP_CYL_BALANCE_OF_MASS: {
            NEED(6, 3, 3, 3, 1, 1, 1, 1);
            NEED_CELLS();
            sp_wrote(s, F[3]);
            return dispatch_kernel<OpCylBalanceOfMass>(s, (int)Pm[0], Pm[1], flags, [&](auto& P) {
                set_v3(s, F[1], P.qp);
                P.qp[3] = sc(s, F[2]);
                P.qp[4] = sc(s, F[4]);
                P.qp[5] = sc(s, F[5]);
                P.Drho = sc(s, F[3]);
                P.two_nu = Pm[2];
            });
        }
        case SP_OP_CYL_FIND_PRESSURE: {
            NEED(4, 4, 3, 1, 1, 1);
            sp_wrote(s, F[1]);
            sp_zeroed(s, F[2]);
            sp_wrote(s, F[3]);
            UCylFindPressure::Params P{sc(s, F[0]), sc(s, F[1]), sc(s, F[2]), sc(s, F[3]), Pm[0], Pm[1], Pm[2], Pm[3]};
            return launch_unary<UCylFindPressure>(s, P);
        }
        case SP_OP_CYL_INTERNAL_FORCE: {
            NEED(6, 4, 3, 3, 1, 1, 3, 1);
            NEED_CELLS();
            sp_wrote(s, F[4]);
            double* pr = nullptr;
            {
                int rc2 = pressure_over_rho2(s, F[2], F[3], &pr);
                if (rc2) return rc2;
            }
            return dispatch_kernel<OpCylInternalForce>(s, (int)Pm[0], Pm[1], flags, [&](auto& P) {
                set_v3(s, F[1], P.qp);
                P.qp[3] = pr;
                P.qp[4] = sc(s, F[3]);
                P.qp[5] = sc(s, F[5]);
                P.a = wv3(s, F[4]);
                P.mu = Pm[2];
                P.eps2 = Pm[3];
            });
        }
        case SP_OP_MOVE_TYPES: {
            NEED(4, 3, 3, 3, 3, 1);
            sp_wrote(s, F[0]);
            sp_zeroed(s, F[2]);
            UMoveTypes::Params P{wv3(s, F[0]), rv3(s, F[1]), wv3(s, F[2]), sc(s, F[3]), Pm[0], Pm[1], Pm[2]};
            return launch_unary<UMoveTypes>(s, P);
        }
        case SP_OP_CYL_ACCELERATE: {
            NEED(4, 3, 3, 3, 3, 1);
            sp_wrote(s, F[1]);
            UCylAccelerate::Params P{rv3(s, F[0]), wv3(s, F[1]), rv3(s, F[2]), sc(s, F[3]), Pm[0], Pm[1], Pm[2]};
            return launch_unary<UCylAccelerate>(s, P);
        }
        case SP_OP_SET_INFLOW_SPEED: {
            NEED(3, 4, 3, 3, 1);
            sp_wrote(s, F[1]);
            USetInflowSpeed::Params P{rv3(s, F[0]), wv3(s, F[1]), sc(s, F[2]), Pm[0], Pm[1], Pm[2], Pm[3]};
            return launch_unary<USetInflowSpeed>(s, P);
        }
        case SP_OP_ROD_FIND_A: {
            NEED(4, 2, 3, 3, 9, 9);
            NEED_CELLS();
            sp_wrote(s, F[2]);
            sp_wrote(s, F[3]);
            return dispatch_kernel<OpRodFindA>(s, (int)Pm[0], Pm[1], flags, [&](auto& P) {
                P.qp[0] = sc(s, F[1]);
                P.qp[1] = sc(s, F[1]) + s->cap;
                P.A = sc(s, F[2]);
                P.H = sc(s, F[3]);
                P.cap = s->cap;
            });
        }
        case SP_OP_ROD_FIND_B: {
            NEED(3, 3, 9, 9, 9);
            sp_wrote(s, F[0]);
            sp_wrote(s, F[2]);
            URodFindB::Params P{sc(s, F[0]), sc(s, F[1]), sc(s, F[2]), s->cap, Pm[0], Pm[1] * Pm[1], Pm[2] * Pm[2]};
            return launch_unary<URodFindB>(s, P);
        }
        case SP_OP_ROD_FIND_F: {
            NEED(6, 4, 3, 3, 3, 9, 9, 3);
            NEED_CELLS();
            sp_wrote(s, F[5]);
            return dispatch_kernel<OpRodFindF>(s, (int)Pm[0], Pm[1], flags, [&](auto& P) {
                const double *A = sc(s, F[3]), *B = sc(s, F[4]), *X = sc(s, F[2]);
                const long long cap = s->cap;
                P.qp[0] = A; P.qp[1] = A + cap; P.qp[2] = A + 3 * cap; P.qp[3] = A + 4 * cap;
                P.qp[4] = B; P.qp[5] = B + cap; P.qp[6] = B + 3 * cap; P.qp[7] = B + 4 * cap;
                P.qp[8] = X; P.qp[9] = X + cap;
                set_v3(s, F[1], P.qp + 10);
                P.f = wv3(s, F[5]);
                P.two_m_vol = Pm[2];
                P.nu = Pm[3];
            });
        }
        case SP_OP_ROD_PULL: {
            NEED(2, 2, 3, 3);
            sp_wrote(s, F[1]);
            URodPull::Params P{sc(s, F[0]), sc(s, F[1]) + s->cap, Pm[0], Pm[1]};
            return launch_unary<URodPull>(s, P);
        }
        case SP_OP_ROD_UPDATE_V: {
            NEED(3, 3, 3, 3, 3);
            sp_wrote(s, F[0]);
            URodUpdateV::Params P{wv3(s, F[0]), rv3(s, F[1]), sc(s, F[2]), Pm[0], Pm[1], Pm[2]};
            return launch_unary<URodUpdateV>(s, P);
        }
        case SP_OP_ROD_UPDATE_X: {
            NEED(6, 1, 3, 3, 9, 9, 3, 1);
            sp_wrote(s, F[0]);
            sp_zeroed(s, F[2]);
            sp_zeroed(s, F[3]);
            sp_zeroed(s, F[4]);
            sp_zeroed(s, F[5]);
            URodUpdateX::Params P{wv3(s, F[0]), rv3(s, F[1]), sc(s, F[2]), sc(s, F[3]), wv3(s, F[4]), sc(s, F[5]), s->cap, Pm[0]};
            return launch_unary<URodUpdateX>(s, P);
        }
        case SP_OP_ROD_FIND_E: {
            NEED(4, 1, 3, 3, 9, 1);
            NEED_CELLS();
            sp_wrote(s, F[3]);
            return dispatch_kernel<OpRodFindE>(s, SP_KERNEL_WENDLAND2, Pm[0], flags, [&](auto& P) {
                P.qp[0] = sc(s, F[1]);
                P.qp[1] = sc(s, F[1]) + s->cap;
                P.A = sc(s, F[2]);
                P.e = sc(s, F[3]);
                P.cap = s->cap;
            });
        }
        case SP_OP_SHTC_FIND_STRESS: {
            NEED(3, 3, 9, 1, 9);
            sp_wrote(s, F[2]);
            UShtcFindStress::Params P{sc(s, F[0]), sc(s, F[1]), sc(s, F[2]), s->cap, Pm[0] * Pm[0], Pm[1] * Pm[1], Pm[2]};
            return launch_unary<UShtcFindStress>(s, P);
        }
        case SP_OP_SHTC_UPDATE_V: {
            NEED(5, 3, 3, 3, 1, 9, 1);
            NEED_CELLS();
            sp_wrote(s, F[1]);
            return dispatch_kernel<OpShtcUpdateV>(s, (int)Pm[0], Pm[1], flags, [&](auto& P) {
                for (int c = 0; c < 9; c++) P.qp[c] = sc(s, F[3]) + (size_t)c * s->cap;
                P.qp[9] = sc(s, F[2]);
                P.type = sc(s, F[4]);
                P.v = wv3(s, F[1]);
                P.dtm = Pm[2];
            });
        }
        case SP_OP_SHTC_UPDATE_RHO: {
            NEED(4, 3, 3, 3, 1, 1);
            NEED_CELLS();
            sp_wrote(s, F[2]);
            return dispatch_kernel<OpShtcUpdateRho>(s, (int)Pm[0], Pm[1], flags, [&](auto& P) {
                set_v3(s, F[1], P.qp);
                P.type = sc(s, F[3]);
                P.rho = sc(s, F[2]);
                P.dtm = Pm[2];
            });
        }
        case SP_OP_SHTC_CONVECT_A: {
            NEED(5, 4, 3, 3, 1, 9, 1);
            NEED_CELLS();
            sp_wrote(s, F[3]);
            // order-dependent (ldc.jl:98 multiplies the running A_p): always the reference's visiting order
            return dispatch_kernel<OpShtcConvectA>(s, (int)Pm[0], Pm[1], flags | SP_FLAG_STRICT_ORDER, [&](auto& P) {
                set_v3(s, F[1], P.qp);
                P.type = sc(s, F[4]);
                P.rho = sc(s, F[2]);
                P.A = sc(s, F[3]);
                P.cap = s->cap;
                P.dtm = Pm[2];
                P.skip = Pm[3];
            });
        }
        case SP_OP_SHTC_RELAX_A: {
            NEED(1, 2, 9);
            sp_wrote(s, F[0]);
            UShtcRelaxA::Params P{sc(s, F[0]), s->cap, Pm[0], -3.0 / Pm[1]};
            return launch_unary<UShtcRelaxA>(s, P);
        }
        case SP_OP_SHTC_MOVE: {
            NEED(3, 1, 3, 3, 1);
            sp_wrote(s, F[0]);
            UShtcMove::Params P{wv3(s, F[0]), rv3(s, F[1]), sc(s, F[2]), Pm[0]};
            return launch_unary<UShtcMove>(s, P);
        }
        case SP_OP_BE_FIND_L: {
            NEED(5, 3, 3, 3, 1, 9, 9);
            NEED_CELLS();
            sp_wrote(s, F[3]);
            sp_wrote(s, F[4]);
            return dispatch_kernel<OpBeFindL>(s, (int)Pm[0], Pm[1], flags, [&](auto& P) {
                P.qp[0] = sc(s, F[2]);
                P.qp[1] = sc(s, F[1]);
                P.qp[2] = sc(s, F[1]) + s->cap;
                P.T = sc(s, F[3]);
                P.L = sc(s, F[4]);
                P.J = nullptr;
                P.Kf = nullptr;
                P.cap = s->cap;
                P.rho0 = Pm[2];
                P.h = Pm[1];
            });
        }
        case SP_OP_BE_UPDATE_A: {
            NEED(3, 1, 9, 9, 9);
            sp_wrote(s, F[0]);
            sp_wrote(s, F[2]);
            UBeUpdateA::Params P{sc(s, F[0]), sc(s, F[1]), sc(s, F[2]), s->cap, Pm[0]};
            return launch_unary<UBeUpdateA>(s, P);
        }
        case SP_OP_BE_FIND_J: {
            NEED(5, 3, 3, 1, 9, 1, 1);
            NEED_CELLS();
            sp_wrote(s, F[2]);
            sp_wrote(s, F[3]);
            sp_wrote(s, F[4]);
            return dispatch_kernel<OpBeFindJ>(s, (int)Pm[0], Pm[1], flags, [&](auto& P) {
                P.qp[0] = sc(s, F[1]);
                P.qp[1] = sc(s, F[1]);  // velocities are not read by find_J!: any valid plane
                P.qp[2] = sc(s, F[1]);
                P.T = sc(s, F[2]);
                P.L = nullptr;
                P.J = sc(s, F[3]);
                P.Kf = sc(s, F[4]);
                P.cap = s->cap;
                P.rho0 = Pm[2];
                P.h = Pm[1];
            });
        }
        case SP_OP_BE_FIND_T: {
            NEED(4, 3, 9, 9, 1, 1);
            sp_wrote(s, F[1]);
            sp_wrote(s, F[2]);
            UBeFindT::Params P{sc(s, F[0]), sc(s, F[1]), sc(s, F[2]), sc(s, F[3]), s->cap, Pm[0], Pm[1] * Pm[1], Pm[2] * Pm[2]};
            return launch_unary<UBeFindT>(s, P);
        }
        case SP_OP_BE_FIND_F: {
            NEED(5, 4, 3, 1, 9, 1, 3);
            NEED_CELLS();
            sp_wrote(s, F[4]);
            return dispatch_kernel<OpBeFindF>(s, (int)Pm[0], Pm[1], flags, [&](auto& P) {
                const double* T = sc(s, F[2]);
                const long long cap = s->cap;
                P.qp[0] = sc(s, F[1]);
                P.qp[1] = sc(s, F[3]);
                P.qp[2] = T; P.qp[3] = T + cap; P.qp[4] = T + 3 * cap; P.qp[5] = T + 4 * cap;
                P.f = wv3(s, F[4]);
                P.rho0 = Pm[2];
                P.cp2 = Pm[3] * Pm[3];
                P.h = Pm[1];
            });
        }
        case SP_OP_BE_RESET: {
            const int ncs[] = {3, 9, 9, 1, 1, 1, 1};
            int rc2 = sp_check_fields(s, F, nf, ncs, 7);
            if (rc2) return rc2;
            if (np != 0) return sp_fail(s, SP_ERR_INVALID, "wrong number of parameters for this operator");
            sp_zeroed(s, F[0]);
            sp_zeroed(s, F[1]);
            sp_zeroed(s, F[2]);
            sp_wrote(s, F[3]);
            sp_wrote(s, F[4]);
            UBeReset::Params P{wv3(s, F[0]), sc(s, F[1]), sc(s, F[2]), sc(s, F[3]), sc(s, F[4]), sc(s, F[5]), sc(s, F[6]), s->cap};
            return launch_unary<UBeReset>(s, P);
        }
        case SP_OP_BE_UPDATE_V: {
            NEED(3, 1, 3, 3, 1);
            sp_wrote(s, F[0]);
            UBeUpdateV::Params P{wv3(s, F[0]), rv3(s, F[1]), sc(s, F[2]), Pm[0]};
            return launch_unary<UBeUpdateV>(s, P);
        }
        case SP_OP_TW_FIND_L: {
            NEED(5, 3, 3, 3, 1, 9, 9);
            NEED_CELLS();
            sp_wrote(s, F[3]);
            sp_wrote(s, F[4]);
            return dispatch_kernel<OpTwFindL>(s, (int)Pm[0], Pm[1], flags, [&](auto& P) {
                P.qp[0] = sc(s, F[2]);
                set_v3(s, F[1], P.qp + 1);
                P.T = sc(s, F[3]);
                P.L = sc(s, F[4]);
                P.J = nullptr;
                P.Kf = nullptr;
                P.cap = s->cap;
                P.rho0 = Pm[2];
                P.h = Pm[1];
            });
        }
        case SP_OP_TW_UPDATE_A: {
            NEED(3, 1, 9, 9, 9);
            sp_wrote(s, F[0]);
            sp_wrote(s, F[2]);
            UTwUpdateA::Params P{sc(s, F[0]), sc(s, F[1]), sc(s, F[2]), s->cap, Pm[0]};
            return launch_unary<UTwUpdateA>(s, P);
        }
        case SP_OP_TW_FIND_J: {
            NEED(5, 3, 3, 1, 9, 1, 1);
            NEED_CELLS();
            sp_wrote(s, F[2]);
            sp_wrote(s, F[3]);
            sp_wrote(s, F[4]);
            return dispatch_kernel<OpTwFindJ>(s, (int)Pm[0], Pm[1], flags, [&](auto& P) {
                for (int k = 0; k < 4; k++) P.qp[k] = sc(s, F[1]);  // velocities are not read by find_J!: any valid plane
                P.T = sc(s, F[2]);
                P.L = nullptr;
                P.J = sc(s, F[3]);
                P.Kf = sc(s, F[4]);
                P.cap = s->cap;
                P.rho0 = Pm[2];
                P.h = Pm[1];
            });
        }
        case SP_OP_TW_FIND_T: {
            NEED(4, 3, 9, 9, 1, 1);
            sp_wrote(s, F[1]);
            sp_wrote(s, F[2]);
            UTwFindT::Params P{sc(s, F[0]), sc(s, F[1]), sc(s, F[2]), sc(s, F[3]), s->cap, Pm[0], Pm[1] * Pm[1], Pm[2] * Pm[2]};
            return launch_unary<UTwFindT>(s, P);
        }
        case SP_OP_TW_FIND_F: {
            NEED(5, 4, 3, 1, 9, 1, 3);
            NEED_CELLS();
            sp_wrote(s, F[4]);
            return dispatch_kernel<OpTwFindF>(s, (int)Pm[0], Pm[1], flags, [&](auto& P) {
                P.qp[0] = sc(s, F[1]);
                P.qp[1] = sc(s, F[3]);
                for (int c = 0; c < 9; c++) P.qp[2 + c] = sc(s, F[2]) + (size_t)c * s->cap;
                P.f = wv3(s, F[4]);
                P.rho0 = Pm[2];
                P.cp2 = Pm[3] * Pm[3];
                P.h = Pm[1];
            });
        }
        case SP_OP_TW_UPDATE_V: {
            NEED(4, 1, 3, 3, 3, 1);
            sp_wrote(s, F[1]);
            UTwUpdateV::Params P{sc(s, F[0]) + 2 * s->cap, wv3(s, F[1]), rv3(s, F[2]), sc(s, F[3]), Pm[0]};
            return launch_unary<UTwUpdateV>(s, P);
        }
        case SP_OP_TA_FIND_T: {
            NEED(4, 3, 9, 9, 1, 1);
            sp_wrote(s, F[1]);
            sp_wrote(s, F[2]);
            UTaFindT::Params P{sc(s, F[0]), sc(s, F[1]), sc(s, F[2]), sc(s, F[3]), s->cap, Pm[0], Pm[1] * Pm[1], Pm[2] * Pm[2]};
            return launch_unary<UTaFindT>(s, P);
        }
        case SP_OP_TA_FIND_F: {
            NEED(5, 3, 3, 1, 9, 1, 3);
            NEED_CELLS();
            sp_wrote(s, F[4]);
            return dispatch_kernel<OpTaFindF>(s, (int)Pm[0], Pm[1], flags, [&](auto& P) {
                P.qp[0] = sc(s, F[1]);
                P.qp[1] = sc(s, F[3]);
                for (int c = 0; c < 9; c++) P.qp[2 + c] = sc(s, F[2]) + (size_t)c * s->cap;
                P.f = wv3(s, F[4]);
                P.cpr2 = Pm[2];
                P.h = Pm[1];
            });
        }
        case SP_OP_TA_UPDATE_V: {
            NEED(5, 4, 3, 3, 3, 1, 1);
            sp_wrote(s, F[1]);
            UTaUpdateV::Params P{rv3(s, F[0]), wv3(s, F[1]), rv3(s, F[2]), sc(s, F[3]), sc(s, F[4]), Pm[0], Pm[1], Pm[2], Pm[3]};
            return launch_unary<UTaUpdateV>(s, P);
        }
        case SP_OP_TA_UPDATE_X: {
            NEED(4, 4, 3, 3, 3, 1);
            sp_wrote(s, F[0]);
            UTaUpdateX::Params P{wv3(s, F[0]), rv3(s, F[1]), rv3(s, F[2]), sc(s, F[3]), Pm[0], Pm[1], Pm[2], Pm[3]};
            return launch_unary<UTaUpdateX>(s, P);
        }
    }
    return sp_fail(s, SP_ERR_INVALID, "unknown operator id");
}

// ---- ISPH: coefficients of the pressure-Poisson operator in the layout of the cached neighbour lists (ELL).
// A is constant during a CG solve (positions are fixed), so A_ij = 2h^2 (m/rho) rDk(r_ij) is evaluated once per
// solve — what assemble_matrix(sys, projection_matrix) does in the reference (core.jl:196-225,
// collapse_dry_implicit.jl:154-163) — and every mat-vec of the CG is then one gather per entry (sp_isph.cu).
template <class K>
__global__ void __launch_bounds__(128) k_poisson_coeffs(SweepCtx c, const int* __restrict__ cnt, const int* __restrict__ ids,
                                                        const double* __restrict__ L, const double* __restrict__ lambda,
                                                        const double* __restrict__ type, double off_coef, double h2,
                                                        double C_free, SpKC kc, double* __restrict__ aval,
                                                        double* __restrict__ diag, int* __restrict__ overflow) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *c.alive) return;  // the dead tail [alive, n) takes no part (sp_internal.cuh)
    double Aii = h2 * L[i];
    if (type[i] == 0.0) Aii += C_free * fmax(lambda[i], 0.0);
    diag[i] = Aii;
    const int n_nb = cnt[i];
    if (n_nb > c.capk) {
        *overflow = 1;  // the caller falls back to the matrix-free operator
        return;
    }
    const double xi = c.x[i], yi = c.y[i], zi = c.z[i];
    const size_t base = ((size_t)(i >> 5) * c.capk << 5) + (i & 31);
    for (int k = 0; k < n_nb; k++) {
        const int j = ids[base + ((size_t)k << 5)];
        const double dx = __dsub_rn(xi, c.x[j]), dy = __dsub_rn(yi, c.y[j]), dz = __dsub_rn(zi, c.z[j]);
        aval[base + ((size_t)k << 5)] = off_coef * K::rD(kc, sp_sqrt_fast(sp_d2(dx, dy, dz)));
    }
}

// how many of the 9/27 stencil offsets are zero = how often the reference visits a particle's own cell (1, except on
// domains with fewer than three cells along an axis: the double-visit quirk of the linear key offsets)
int sp_self_visits(const sp_system* s) {
    int m = 0;
    for (int k = 0; k < s->n_key_diff; k++) m += s->key_diff[k] == 0;
    return m < 1 ? 1 : m;
}

// fields {x, L, lambda, type}; params {kernel, m, h, rho, C_free}; aval: cap*capk doubles, diag: n doubles
int sp_poisson_ell_build(sp_system* s, const int32_t* F, const double* Pm, double* aval, double* diag, int* d_overflow,
                         const int** ids_out, const int** cnt_out) {
    SweepCtx c;
    sp_sweep_ctx(s, c);
    int rc = sp_ensure_nbr_cache(s, c);
    if (rc) return rc;
    SpKC kc;
    const int kernel = (int)Pm[0];
    if (!sp_make_kc(kernel, Pm[2], &kc)) return sp_fail(s, SP_ERR_INVALID, "unknown SPH kernel id");
    const double mult = (double)sp_self_visits(s);  // the diagonal is the SUM over the visits of the own cell
    const double off_coef = 2.0 * (Pm[2] * Pm[2]) * Pm[1] / Pm[3], h2 = Pm[2] * Pm[2] * mult, C_free = Pm[4] * mult;
    const double *L = sc(s, F[1]), *lam = sc(s, F[2]), *ty = sc(s, F[3]);
    const unsigned nb = sp_blocks(s->n, 128);
    switch (kernel) {
        case SP_KERNEL_WENDLAND1:
        case SP_KERNEL_WENDLAND2:
        case SP_KERNEL_WENDLAND3:
            SP_LAUNCH(s, k_poisson_coeffs<KWendland>, nb, 128, 0, c, s->nbr_cnt, s->nbr_ids, L, lam, ty, off_coef, h2, C_free,
                      kc, aval, diag, d_overflow);
            break;
        case SP_KERNEL_SPLINE23:
            SP_LAUNCH(s, k_poisson_coeffs<KSpline23>, nb, 128, 0, c, s->nbr_cnt, s->nbr_ids, L, lam, ty, off_coef, h2, C_free,
                      kc, aval, diag, d_overflow);
            break;
        default:
            SP_LAUNCH(s, k_poisson_coeffs<KSpline24>, nb, 128, 0, c, s->nbr_cnt, s->nbr_ids, L, lam, ty, off_coef, h2, C_free,
                      kc, aval, diag, d_overflow);
    }
    *ids_out = s->nbr_ids;
    *cnt_out = s->nbr_cnt;
    return SP_OK;
}
// settle the cached lists (and their capacity) for the current positions; returns the capacity per target
int sp_nbr_prepare(sp_system* s, int* capk) {
    SweepCtx c;
    sp_sweep_ctx(s, c);
    int rc = sp_ensure_nbr_cache(s, c);
    if (rc) return rc;
    *capk = s->nbr_capk;
    return SP_OK;
}

// fused unary passes of the step programs (sp_program.cu)
// fields {v, Dv, x, type}; params {hdt, gx, gy, gz, dt_move}
int sp_kick_kick_move_impl(sp_system* s, const int32_t* F, const double* Pm) {
    sp_wrote(s, F[0]);
    sp_zeroed(s, F[1]);
    sp_wrote(s, F[2]);
    UKickKickMove::Params P{wv3(s, F[0]), wv3(s, F[1]), wv3(s, F[2]), sc(s, F[3]), Pm[0], Pm[1], Pm[2], Pm[3], Pm[4]};
    return launch_unary<UKickKickMove>(s, P);
}
// fields {rho, Drho, P}; params {dt, c2, rho0, P0}; also fills the scratch field _pr = P/rho^2
int sp_find_pressure_pr_impl(sp_system* s, const int32_t* F, const double* Pm) {
    int32_t fid;
    int rc = sp_add_field(s, "_pr", 1, &fid);
    if (rc) return rc;
    s->fields[fid].transient = true;
    sp_wrote(s, F[0]);
    sp_zeroed(s, F[1]);
    sp_wrote(s, F[2]);
    UFindPressurePr::Params P{sc(s, F[0]), sc(s, F[1]), sc(s, F[2]), s->fields[fid].d, Pm[0], Pm[1], Pm[2], Pm[3]};
    return launch_unary<UFindPressurePr>(s, P);
}

// fields {x, L, lambda, type, p_in, y_out}; params {kernel, m, h, rho, C_free}
int sp_poisson_apply_impl(sp_system* s, const int32_t* F, int32_t nf, const double* Pm, int32_t np) {
    const int32_t op = 0;
    (void)op;
    NEED(6, 5, 3, 1, 1, 1, 1, 1);
    NEED_CELLS();
    sp_wrote(s, F[5]);
    return dispatch_kernel<OpPoissonApply>(s, (int)Pm[0], Pm[2], 0, [&](auto& P) {
        P.L = sc(s, F[1]);
        P.lambda = sc(s, F[2]);
        P.type = sc(s, F[3]);
        P.qp[0] = sc(s, F[4]);
        P.y = sc(s, F[5]);
        P.off_coef = 2.0 * (Pm[2] * Pm[2]) * Pm[1] / Pm[3];
        // assemble_matrix does not skip p == q (core.jl:196-225): the diagonal entry is pushed once per visit of the own
        // cell and sparse() sums them — more than once only on domains with fewer than three cells along an axis
        const double mult = (double)sp_self_visits(s);
        P.h2 = Pm[2] * Pm[2] * mult;
        P.C_free = Pm[4] * mult;
    });
}
#undef NEED
#undef NEED_CELLS

// ------------------------------------------------------------------ neighbour-list export (parity view)
__global__ void k_nbr_count(SpGrid g, SweepCtx c, const int* ref, long long* counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *c.alive) return;  // the dead tail [alive, n) takes no part (sp_internal.cuh)
    int cnt = 0;
    sp_for_candidates<true>(g, c, c.x[i], c.y[i], c.z[i], [&](int j, double, double, double, double d2) {
        if (d2 > g.T2 || j == i) return;
        cnt++;
    });
    counts[ref[i]] = cnt;
}
__global__ void k_nbr_fill(SpGrid g, SweepCtx c, const int* ref, const long long* offsets, long long* ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *c.alive) return;  // the dead tail [alive, n) takes no part (sp_internal.cuh)
    long long o = offsets[ref[i]];
    sp_for_candidates<true>(g, c, c.x[i], c.y[i], c.z[i], [&](int j, double, double, double, double d2) {
        if (d2 > g.T2 || j == i) return;
        ids[o++] = (long long)ref[j] + 1;
    });
}

// the lists the default sweeps replay (cached), in their visiting order
__global__ void k_cache_count(SpGrid g, SweepCtx c, const int* cnt, const int* ref, long long* counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *c.alive) return;  // the dead tail [alive, n) takes no part (sp_internal.cuh)
    counts[ref[i]] = cnt[i];
}
__global__ void k_cache_fill(SpGrid g, SweepCtx c, const int* cnt, const int* lists, const int* ref,
                             const long long* offsets, long long* ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *c.alive) return;  // the dead tail [alive, n) takes no part (sp_internal.cuh)
    long long o = offsets[ref[i]];
    const int n_nb = cnt[i];
    if (n_nb <= c.capk) {
        const int* col = lists + ((size_t)(i >> 5) * c.capk << 5) + (i & 31);
        for (int k = 0; k < n_nb; k++) ids[o++] = (long long)ref[col[k << 5]] + 1;
    } else {
        sp_for_candidates<false>(g, c, c.x[i], c.y[i], c.z[i], [&](int j, double, double, double, double d2) {
            if (d2 > g.T2 || j == i) return;
            ids[o++] = (long long)ref[j] + 1;
        });
    }
}

// shared tail of the two list exports: counts (by reference index) -> offsets on the host, then fill
template <class Fill>
static int sp_export_lists(sp_system* s, long long* counts, int64_t* offsets, int64_t* ids, int64_t ids_cap, Fill&& fill) {
    const long long n = s->n;
    std::vector<long long> h(n + 1);
    SP_CUDA(s, cudaMemcpyAsync(h.data(), counts, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    long long run = 0;
    for (long long i = 0; i < n; i++) {
        long long cnt = h[i];
        h[i] = run;
        offsets[i] = run;
        run += cnt;
    }
    h[n] = run;
    offsets[n] = run;
    if (!ids) return SP_OK;
    if (ids_cap < run) return sp_fail(s, SP_ERR_INVALID, "ids buffer too small");
    if (run == 0) return SP_OK;
    long long *d_off = nullptr, *d_ids = nullptr;
    SP_CUDA(s, sp_dmalloc(&d_off, (size_t)(n + 1) * sizeof(long long)));
    cudaError_t e = sp_dmalloc(&d_ids, (size_t)run * sizeof(long long));
    if (e != cudaSuccess) {
        sp_dfree(s, d_off);
        return sp_fail_cuda(s, e, "cudaMalloc ids", __FILE__, __LINE__);
    }
    cudaMemcpyAsync(d_off, h.data(), (size_t)(n + 1) * sizeof(long long), cudaMemcpyHostToDevice, s->stream);
    fill(d_off, d_ids);
    s->launches++;
    cudaMemcpyAsync(ids, d_ids, (size_t)run * sizeof(long long), cudaMemcpyDeviceToHost, s->stream);
    e = cudaStreamSynchronize(s->stream);
    sp_dfree(s, d_off);
    sp_dfree(s, d_ids);
    if (e != cudaSuccess) return sp_fail_cuda(s, e, "neighbour list export", __FILE__, __LINE__);
    return SP_OK;
}

extern "C" {

int32_t sp_build_neighbour_lists(sp_system* s) {
    if (!s) return SP_ERR_INVALID;
    if (!s->have_cells) return sp_fail(s, SP_ERR_STATE, "no cell list: call sp_create_cell_list first");
    SP_CUDA(s, cudaSetDevice(s->device));
    int rc = sp_time_begin(s);
    if (rc) return rc;
    if (s->n > 0) {
        SweepCtx c;
        sp_sweep_ctx(s, c);
        if ((rc = sp_ensure_nbr_cache(s, c))) return rc;
    }
    return sp_time_end(s);
}

int32_t sp_neighbour_list_capacity(sp_system* s, int32_t* capk) {
    if (!s || !capk) return SP_ERR_INVALID;
    *capk = s->nbr_capk;
    return SP_OK;
}

int32_t sp_get_sweep_neighbour_lists(sp_system* s, int64_t* offsets, int64_t* ids, int64_t ids_cap) {
    if (!s || !offsets) return SP_ERR_INVALID;
    if (!s->have_cells) return sp_fail(s, SP_ERR_STATE, "no cell list: call sp_create_cell_list first");
    SP_CUDA(s, cudaSetDevice(s->device));
    if (int rcs = sp_settle(s)) return rcs;
    const long long n = s->n;
    offsets[0] = 0;
    if (n == 0) return SP_OK;
    SweepCtx c;
    sp_sweep_ctx(s, c);
    int rc = sp_ensure_nbr_cache(s, c);
    if (rc) return rc;
    if ((rc = sp_ensure_stage(s, n + 1))) return rc;
    long long* counts = (long long*)s->stage;
    SP_LAUNCH(s, k_cache_count, sp_blocks(n, 128), 128, 0, s->g, c, s->nbr_cnt, s->ref, counts);
    return sp_export_lists(s, counts, offsets, ids, ids_cap, [&](long long* d_off, long long* d_ids) {
        k_cache_fill<<<sp_blocks(n, 128), 128, 0, s->stream>>>(s->g, c, s->nbr_cnt, s->nbr_ids, s->ref, d_off, d_ids);
    });
}

int32_t sp_apply(sp_system* s, int32_t op, const int32_t* fields, int32_t nfields, const double* params, int32_t nparams,
                 int32_t flags) {
    if (!s) return SP_ERR_INVALID;
    SP_CUDA(s, cudaSetDevice(s->device));
    int rc = sp_time_begin(s);
    if (rc) return rc;
    if ((rc = sp_apply_impl(s, op, fields, nfields, params, nparams, flags))) return rc;
    return sp_time_end(s);
}

int32_t sp_poisson_apply(sp_system* s, const int32_t* fields, int32_t nfields, const double* params, int32_t nparams) {
    if (!s) return SP_ERR_INVALID;
    SP_CUDA(s, cudaSetDevice(s->device));
    int rc = sp_time_begin(s);
    if (rc) return rc;
    if ((rc = sp_poisson_apply_impl(s, fields, nfields, params, nparams))) return rc;
    return sp_time_end(s);
}

int32_t sp_get_neighbour_lists(sp_system* s, int64_t* offsets, int64_t* ids, int64_t ids_cap) {
    if (!s || !offsets) return SP_ERR_INVALID;
    if (!s->have_cells) return sp_fail(s, SP_ERR_STATE, "no cell list: call sp_create_cell_list first");
    SP_CUDA(s, cudaSetDevice(s->device));
    if (int rcs = sp_settle(s)) return rcs;
    const long long n = s->n;
    offsets[0] = 0;
    if (n == 0) return SP_OK;
    SweepCtx c;
    sp_sweep_ctx(s, c);
    int rc = sp_ensure_stage(s, n + 1);
    if (rc) return rc;
    long long* counts = (long long*)s->stage;
    SP_LAUNCH(s, k_nbr_count, sp_blocks(n, 128), 128, 0, s->g, c, s->ref, counts);
    return sp_export_lists(s, counts, offsets, ids, ids_cap, [&](long long* d_off, long long* d_ids) {
        k_nbr_fill<<<sp_blocks(n, 128), 128, 0, s->stream>>>(s->g, c, s->ref, d_off, d_ids);
    });
}

}  // extern "C"
