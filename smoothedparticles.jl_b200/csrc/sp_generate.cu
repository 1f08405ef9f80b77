// sp_generate.cu — generate_particles! on the device (reference: src/grids.jl:52-144, 253-258 lattice coverings;
// src/geometry.jl:15-258 shapes).  The reference pushes one heap object per lattice point in a serial double/triple
// loop; here every lattice point of the bounding index box is tested in parallel against the shape, the survivors
// are compacted by an exclusive scan — which keeps the reference's generation ORDER (first index outermost, last
// innermost) — and written straight into the position planes behind the existing particles.  Positions are the same
// doubles (i*dr etc., no FMA), membership is the same closed-interval / un-fused arithmetic, so the result is
// bit-identical to the host generator (smoothedparticles.jl_b200/geometry.py, tests/test_generate_gpu.py).
//
// A shape is a postfix program of sp_shape_node (children before parents, root last):
//   BOX lo[3] hi[3] | CIRCLE cx cy r^2 | BALL cx cy cz r^2 | HALFSPACE axis op bound | UNION a b | INTERSECTION a b |
//   DIFFERENCE a b | BOUNDARY_LAYER a  (not in a, but x + dx in a for one of the lattice offsets, geometry.jl:198-219)
#include <algorithm>

#include "sp_internal.cuh"

#define SP_SHAPE_MAX_NODES 32 /* results are kept in one 32-bit word */

// evaluate nodes 0..upto at one point.  LAYERS = false: the sub-programme below a boundary layer (no layer nodes in it,
// checked on the host); LAYERS = true: the whole programme, a layer node re-evaluates its child's sub-programme at the
// shifted points.  No recursion, the programme lives in global memory (read-only, broadcast to the whole warp).
template <bool LAYERS>
__device__ __forceinline__ bool shape_eval(const sp_shape_node* __restrict__ node, int upto, double x, double y, double z,
                                           const double* __restrict__ offsets, int n_off) {
    unsigned res = 0u;  // bit k = result of node k (SP_SHAPE_MAX_NODES <= 32)
    for (int k = 0; k <= upto; k++) {
        const int kind = node[k].kind, a = node[k].a, b = node[k].b;
        const double* p = node[k].p;
        bool r = false;
        switch (kind) {
            case SP_SHAPE_BOX:
                r = p[0] <= x && x <= p[3] && p[1] <= y && y <= p[4] && p[2] <= z && z <= p[5];
                break;
            case SP_SHAPE_CIRCLE: {
                const double u = __dsub_rn(x, p[0]), v = __dsub_rn(y, p[1]);
                r = __dadd_rn(__dmul_rn(u, u), __dmul_rn(v, v)) <= p[2];
                break;
            }
            case SP_SHAPE_BALL: {
                const double u = __dsub_rn(x, p[0]), v = __dsub_rn(y, p[1]), w = __dsub_rn(z, p[2]);
                r = __dadd_rn(__dadd_rn(__dmul_rn(u, u), __dmul_rn(v, v)), __dmul_rn(w, w)) <= p[3];
                break;
            }
            case SP_SHAPE_HALFSPACE: {
                const double v = a == 0 ? x : (a == 1 ? y : z);
                r = b == 0 ? v < p[0] : b == 1 ? v <= p[0] : b == 2 ? v > p[0] : v >= p[0];
                break;
            }
            case SP_SHAPE_UNION: r = ((res >> a) | (res >> b)) & 1u; break;
            case SP_SHAPE_INTERSECTION: r = ((res >> a) & (res >> b)) & 1u; break;
            case SP_SHAPE_DIFFERENCE: r = ((res >> a) & ~(res >> b)) & 1u; break;
            case SP_SHAPE_BOUNDARY_LAYER:
                if constexpr (LAYERS) {
                    if (!((res >> a) & 1u))
                        for (int o = 0; o < n_off && !r; o++)
                            r = shape_eval<false>(node, a, __dadd_rn(x, offsets[3 * o]), __dadd_rn(y, offsets[3 * o + 1]),
                                                  __dadd_rn(z, offsets[3 * o + 2]), offsets, n_off);
                }
                break;
        }
        res |= (r ? 1u : 0u) << k;
    }
    return (res >> upto) & 1u;
}

// lattice point t of the chunk -> coordinates, in the reference's loop order (i outermost, k innermost)
__device__ __forceinline__ void lattice_point(int grid, double dr, double ha, double hb, long long t, long long i0,
                                              long long j0, long long k0, long long nj, long long nk, double* x,
                                              double* y, double* z) {
    const long long k = k0 + t % nk;
    const long long j = j0 + (t / nk) % nj;
    const long long i = i0 + t / (nk * nj);
    if (grid == SP_GRID_HEXAGONAL) {  // grids.jl:80-86: x1 = (i + (j % 2)/2)*a, x2 = j*b  (% = C remainder)
        *x = __dmul_rn(__dadd_rn((double)i, (double)(j % 2) / 2.0), ha);
        *y = __dmul_rn((double)j, hb);
        *z = 0.0;
    } else {
        *x = __dmul_rn((double)i, dr);
        *y = __dmul_rn((double)j, dr);
        *z = grid == SP_GRID_CUBIC ? __dmul_rn((double)k, dr) : 0.0;
    }
}

__global__ void __launch_bounds__(256) k_gen_flags(const sp_shape_node* __restrict__ nodes, int n_nodes,
                                                   const double* __restrict__ offsets, int n_off, int grid,
                                                   double dr, double ha, double hb, long long i0, long long j0,
                                                   long long k0, long long nj, long long nk, long long count,
                                                   int* __restrict__ flag, int* __restrict__ pos) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= count) return;
    double x, y, z;
    lattice_point(grid, dr, ha, hb, t, i0, j0, k0, nj, nk, &x, &y, &z);
    const int f = shape_eval<true>(nodes, n_nodes - 1, x, y, z, offsets, n_off) ? 1 : 0;
    flag[t] = f;
    pos[t] = f;
}

__global__ void __launch_bounds__(256) k_gen_scatter(int grid, double dr, double ha, double hb, long long i0, long long j0,
                                                     long long k0, long long nj, long long nk, long long count,
                                                     const int* __restrict__ flag, const int* __restrict__ pos,
                                                     double* __restrict__ X, long long cap, long long base) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= count || !flag[t]) return;
    double x, y, z;
    lattice_point(grid, dr, ha, hb, t, i0, j0, k0, nj, nk, &x, &y, &z);
    const long long s = base + pos[t];
    X[s] = x;
    X[cap + s] = y;
    X[2 * cap + s] = z;
}

__global__ void k_gen_fill(double* f, long long cap, int ncomp, double value, long long from, long long to) {
    const long long s = from + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= to) return;
    for (int c = 0; c < ncomp; c++) f[(size_t)c * cap + s] = value;
}

// scratch arrays of one call: released on every exit path, the error returns of SP_CUDA / SP_LAUNCH included
namespace {
struct ScratchGuard {
    sp_system* s;
    void* p[4] = {nullptr, nullptr, nullptr, nullptr};
    explicit ScratchGuard(sp_system* sys) : s(sys) {}
    ~ScratchGuard() {
        bool any = false;
        for (void* q : p) any = any || q;
        if (!any) return;
        if (s->stream) cudaStreamSynchronize(s->stream);
        for (void* q : p)
            if (q) sp_dfree_impl(q);
        cudaGetLastError();
    }
};
}  // namespace

extern "C" int32_t sp_generate_particles(sp_system* s, int32_t grid, double dr, const sp_shape_node* nodes, int32_t n_nodes,
                                         const double* offsets, int32_t n_off, const int64_t irange[6],
                                         const int32_t* fill_fields, const double* fill_values, int32_t n_fill,
                                         int64_t* n_added) {
    if (!s || !nodes || !irange || n_nodes <= 0) return SP_ERR_INVALID;
    if (n_nodes > SP_SHAPE_MAX_NODES) return sp_fail(s, SP_ERR_INVALID, "shape program too long");
    if (grid != SP_GRID_SQUARE && grid != SP_GRID_HEXAGONAL && grid != SP_GRID_CUBIC)
        return sp_fail(s, SP_ERR_INVALID, "unknown grid kind");
    if (s->slab) return sp_fail(s, SP_ERR_STATE, "generate on the host side of a slab system (ownership is per rank)");
    SP_CUDA(s, cudaSetDevice(s->device));
    int n_layers = 0;
    for (int k = 0; k < n_nodes; k++) {
        const int kind = nodes[k].kind;
        const bool binary = kind == SP_SHAPE_UNION || kind == SP_SHAPE_INTERSECTION || kind == SP_SHAPE_DIFFERENCE;
        if (binary && (nodes[k].a < 0 || nodes[k].a >= k || nodes[k].b < 0 || nodes[k].b >= k))
            return sp_fail(s, SP_ERR_INVALID, "shape program: children must precede their parent");
        if (kind == SP_SHAPE_BOUNDARY_LAYER) {
            if (nodes[k].a < 0 || nodes[k].a >= k) return sp_fail(s, SP_ERR_INVALID, "shape program: bad layer child");
            for (int c = 0; c <= nodes[k].a; c++)
                if (nodes[c].kind == SP_SHAPE_BOUNDARY_LAYER)
                    return sp_fail(s, SP_ERR_INVALID, "nested boundary layers are not supported on the device");
            n_layers++;
        }
    }
    if (n_layers && (!offsets || n_off <= 0)) return sp_fail(s, SP_ERR_INVALID, "boundary layer without lattice offsets");
    const long long i0 = irange[0], i1 = irange[1], j0 = irange[2], j1 = irange[3];
    const long long k0 = grid == SP_GRID_CUBIC ? irange[4] : 0, k1 = grid == SP_GRID_CUBIC ? irange[5] : 0;
    if (n_added) *n_added = 0;
    if (i1 < i0 || j1 < j0 || k1 < k0) return SP_OK;
    const long long nj = j1 - j0 + 1, nk = k1 - k0 + 1;
    const double ha = pow(4.0 / 3.0, 0.25) * dr, hb = pow(3.0 / 4.0, 0.25) * dr;  // grids.jl:70-73
    int rc = sp_settle(s);
    if (rc) return rc;
    if ((rc = sp_time_begin(s))) return rc;
    ScratchGuard guard(s);
    sp_shape_node* d_nodes = nullptr;
    SP_CUDA(s, sp_dmalloc(&d_nodes, (size_t)n_nodes * sizeof(sp_shape_node)));
    guard.p[0] = d_nodes;
    SP_CUDA(s, cudaMemcpyAsync(d_nodes, nodes, (size_t)n_nodes * sizeof(sp_shape_node), cudaMemcpyHostToDevice, s->stream));
    double* d_off = nullptr;
    if (n_off > 0) {
        SP_CUDA(s, sp_dmalloc(&d_off, (size_t)3 * n_off * sizeof(double)));
        guard.p[1] = d_off;
        SP_CUDA(s, cudaMemcpyAsync(d_off, offsets, (size_t)3 * n_off * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    }
    // whole i-slabs per chunk, at most ~64 M lattice points at a time
    const long long per_i = nj * nk;
    const long long rows = std::max<long long>(1, (64LL << 20) / per_i);
    const long long chunk_cap = std::min(rows, i1 - i0 + 1) * per_i;
    int *flag = nullptr, *pos = nullptr;
    SP_CUDA(s, sp_dmalloc(&flag, (size_t)chunk_cap * sizeof(int)));
    guard.p[2] = flag;
    SP_CUDA(s, sp_dmalloc(&pos, (size_t)chunk_cap * sizeof(int)));
    guard.p[3] = pos;
    const int B = 256;
    long long total = 0;
    for (long long ia = i0; ia <= i1 && !rc; ia += rows) {
        const long long ib = std::min(ia + rows - 1, i1);
        const long long count = (ib - ia + 1) * per_i;
        SP_LAUNCH(s, k_gen_flags, sp_blocks(count, B), B, 0, d_nodes, n_nodes, d_off, n_off, grid, dr, ha, hb, ia, j0, k0, nj, nk, count, flag,
                  pos);
        if ((rc = sp_exclusive_scan_i32(s, pos, count))) break;
        int last[2];
        SP_CUDA(s, cudaMemcpyAsync(&last[0], pos + count - 1, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        SP_CUDA(s, cudaMemcpyAsync(&last[1], flag + count - 1, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        SP_CUDA(s, cudaStreamSynchronize(s->stream));
        const long long add = (long long)last[0] + last[1];
        if (add == 0) continue;
        const long long n_old = s->n;
        if ((rc = sp_resize(s, n_old + add))) break;  // zero-fills every field of the new particles
        sp_wrote(s, 0);
        SP_LAUNCH(s, k_gen_scatter, sp_blocks(count, B), B, 0, grid, dr, ha, hb, ia, j0, k0, nj, nk, count, flag, pos,
                  s->fields[0].d, s->cap, n_old);
        for (int f = 0; f < n_fill; f++) {
            const int fid = fill_fields[f];
            if (fid <= 0 || fid >= (int)s->fields.size()) {
                rc = sp_fail(s, SP_ERR_INVALID, "generate: bad fill field");
                break;
            }
            sp_wrote(s, fid);
            SP_LAUNCH(s, k_gen_fill, sp_blocks(add, B), B, 0, s->fields[fid].d, s->cap, s->fields[fid].ncomp, fill_values[f],
                      n_old, n_old + add);
        }
        total += add;
    }
    // (the scratch arrays are released by `guard` on every path)
    if (rc) return rc;
    if (n_added) *n_added = total;
    return sp_time_end(s);
}

// ------------------------------------------------------------------ inflow buffer (examples/cylinder.jl:145-156)
// flag[r] = 1 when the particle with reference index r leaves the buffer; indexed by REFERENCE index so that the
// exclusive scan numbers the new particles in the order the script's serial loop pushes them
__global__ void __launch_bounds__(256) k_respawn_flags(const double* __restrict__ x1, const double* __restrict__ type,
                                                       const int* __restrict__ ref, long long n, double from_type,
                                                       double x1_min, int* __restrict__ flag, int* __restrict__ pos) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int f = (type[s] == from_type && x1[s] >= x1_min) ? 1 : 0;
    flag[ref[s]] = f;
    pos[ref[s]] = f;
}
__global__ void __launch_bounds__(256) k_respawn_scatter(double* __restrict__ X, long long cap, double* __restrict__ type,
                                                         const int* __restrict__ ref, long long n_old, double to_type,
                                                         double shift, const int* __restrict__ flag,
                                                         const int* __restrict__ pos) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= n_old) return;
    const int r = ref[s];
    if (!flag[r]) return;
    type[s] = to_type;
    const long long d = n_old + pos[r];
    X[d] = __dsub_rn(X[s], __dmul_rn(shift, 1.0));  // p.x - bc_width*VECX
    X[cap + d] = __dsub_rn(X[cap + s], __dmul_rn(shift, 0.0));
    X[2 * cap + d] = __dsub_rn(X[2 * cap + s], __dmul_rn(shift, 0.0));
}

extern "C" int32_t sp_respawn(sp_system* s, int32_t type_field, double from_type, double to_type, double x1_min,
                              double shift, const int32_t* fill_fields, const double* fill_values, int32_t n_fill,
                              int64_t* n_added) {
    if (!s || n_fill < 0 || (n_fill > 0 && (!fill_fields || !fill_values))) return SP_ERR_INVALID;
    if (type_field <= 0 || type_field >= (int)s->fields.size() || s->fields[type_field].ncomp != 1)
        return sp_fail(s, SP_ERR_INVALID, "respawn: bad type field");
    for (int f = 0; f < n_fill; f++)
        if (fill_fields[f] <= 0 || fill_fields[f] >= (int)s->fields.size())
            return sp_fail(s, SP_ERR_INVALID, "respawn: bad fill field");
    if (s->slab) return sp_fail(s, SP_ERR_STATE, "respawn is not available on a slab system");
    if (n_added) *n_added = 0;
    SP_CUDA(s, cudaSetDevice(s->device));
    int rc = sp_settle(s);
    if (rc) return rc;
    if (s->n == 0) return SP_OK;
    if ((rc = sp_time_begin(s))) return rc;
    const long long n_old = s->n;
    const int B = 256;
    ScratchGuard guard(s);
    int *flag = nullptr, *pos = nullptr;
    SP_CUDA(s, sp_dmalloc(&flag, (size_t)n_old * sizeof(int)));
    guard.p[0] = flag;
    SP_CUDA(s, sp_dmalloc(&pos, (size_t)n_old * sizeof(int)));
    guard.p[1] = pos;
    SP_LAUNCH(s, k_respawn_flags, sp_blocks(n_old, B), B, 0, s->fields[0].d, s->fields[type_field].d, s->ref, n_old,
              from_type, x1_min, flag, pos);
    long long add = 0;
    if (!(rc = sp_exclusive_scan_i32(s, pos, n_old))) {
        int last[2];
        SP_CUDA(s, cudaMemcpyAsync(&last[0], pos + n_old - 1, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        SP_CUDA(s, cudaMemcpyAsync(&last[1], flag + n_old - 1, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        SP_CUDA(s, cudaStreamSynchronize(s->stream));
        add = (long long)last[0] + last[1];
    }
    if (!rc && add > 0 && !(rc = sp_resize(s, n_old + add))) {  // zero-fills the new tail, numbers it n_old, n_old+1, ...
        sp_wrote(s, 0);
        sp_wrote(s, type_field);
        SP_LAUNCH(s, k_respawn_scatter, sp_blocks(n_old, B), B, 0, s->fields[0].d, s->cap, s->fields[type_field].d, s->ref,
                  n_old, to_type, shift, flag, pos);
        for (int f = 0; f < n_fill; f++) {
            const int fid = fill_fields[f];
            sp_wrote(s, fid);
            SP_LAUNCH(s, k_gen_fill, sp_blocks(add, B), B, 0, s->fields[fid].d, s->cap, s->fields[fid].ncomp, fill_values[f],
                      n_old, n_old + add);
        }
        SP_LAUNCH(s, k_gen_fill, sp_blocks(add, B), B, 0, s->fields[type_field].d, s->cap, 1, from_type, n_old, n_old + add);
    }
    if (rc) return rc;
    if (n_added) *n_added = add;
    return sp_time_end(s);
}
