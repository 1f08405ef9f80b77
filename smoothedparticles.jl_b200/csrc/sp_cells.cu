// sp_cells.cu — create_cell_list! on the device (reference src/core.jl:51-90, src/structs.jl:97-106).
//
// Reference semantics reproduced bit-exactly:
//   * particles outside the closed domain box (or with NaN coordinates) are removed with the
//     swap-with-tail rule of core.jl:72-81 — victims in descending index order, the i-th victim slot
//     receives the CURRENT particles[end+1-i] — which fixes the post-removal numbering;
//   * key = find_key(x) with true division and floor;
//   * every cell lists its members in DESCENDING particle index (add_index!, core.jl:26-41).
// Device algorithm: one pass computes keys and a per-cell arrival offset with warp-aggregated atomics,
// an exclusive scan of the per-cell counts gives cell_start, a scatter groups slots by cell, a rank
// pass orders each cell by descending reference index (deterministic whatever the atomic order was),
// and one gather pass permutes every SoA plane so that a cell's particles are contiguous in HBM.
#include <algorithm>

#include "sp_internal.cuh"

// ------------------------------------------------------------------ exclusive scan (int32)
// 256 threads x 4 items; block sums scanned recursively.
#define SCAN_B 256
#define SCAN_ITEMS 4
#define SCAN_TILE (SCAN_B * SCAN_ITEMS)

// gate: when gate_counters is given, the scan belongs to the removal renumbering of a build and is skipped (every
// kernel of it returns at once) unless that build culled particles — see sp_new_victims below
struct ScanGate {
    const int* counters;
    long long n;
};
__device__ __forceinline__ bool scan_gated_off(const ScanGate& g) {
    return g.counters && g.counters[SP_CNT_TRASH] - (int)(g.n - g.counters[SP_CNT_ALIVE]) == 0;
}
__global__ void k_scan_tile(int* data, long long len, int* block_sums, ScanGate gate) {
    __shared__ int warp_tot[SCAN_B / 32];
    if (scan_gated_off(gate)) return;
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = (base + i < len) ? data[base + i] : 0;
        sum += v[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < SCAN_B / 32) ? warp_tot[lane] : 0;
        int wi = w;
#pragma unroll
        for (int d = 1; d < SCAN_B / 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < SCAN_B / 32) warp_tot[lane] = wi - w;  // exclusive warp offsets
        if (lane == SCAN_B / 32 - 1 && block_sums) block_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    int run = warp_tot[warp] + incl - sum;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < len) data[base + i] = run;
        run += v[i];
    }
}
__global__ void k_scan_add(int* data, long long len, const int* block_offsets, ScanGate gate) {
    if (scan_gated_off(gate)) return;
    const long long i = (long long)blockIdx.x * SCAN_TILE + threadIdx.x;
    const int off = block_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        long long j = i + (long long)k * SCAN_B;
        if (j < len) data[j] += off;
    }
}

// short arrays (the cell counts of the 2-D configs, ~6 k cells): the whole scan in ONE block and one launch — the small
// configs are bound by their launch count
#define SCAN_SMALL_MAX 32768
__global__ void __launch_bounds__(1024) k_scan_small(int* data, int len, ScanGate gate) {
    if (scan_gated_off(gate)) return;
    __shared__ int part[1024];
    const int t = threadIdx.x, T = blockDim.x;
    const int chunk = (len + T - 1) / T;
    const int b = min(t * chunk, len), e = min(b + chunk, len);
    int sum = 0;
    for (int i = b; i < e; i++) sum += data[i];
    part[t] = sum;
    __syncthreads();
    for (int d = 1; d < T; d <<= 1) {  // inclusive Hillis-Steele over the 1024 chunk sums
        const int v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int run = part[t] - sum;
    for (int i = b; i < e; i++) {
        const int v = data[i];
        data[i] = run;
        run += v;
    }
}

static int scan_rec(sp_system* s, int* data, long long len, int* tmp, long long tmp_len, ScanGate gate) {
    if (len <= SCAN_SMALL_MAX) {
        SP_LAUNCH(s, k_scan_small, 1, 1024, 0, data, (int)len, gate);
        return SP_OK;
    }
    const long long nb = (len + SCAN_TILE - 1) / SCAN_TILE;
    if (nb <= 1) {
        SP_LAUNCH(s, k_scan_tile, 1, SCAN_B, 0, data, len, (int*)nullptr, gate);
        return SP_OK;
    }
    if (nb > tmp_len) return sp_fail(s, SP_ERR_STATE, "scan scratch too small");
    SP_LAUNCH(s, k_scan_tile, (unsigned)nb, SCAN_B, 0, data, len, tmp, gate);
    int rc = scan_rec(s, tmp, nb, tmp + nb, tmp_len - nb, gate);
    if (rc) return rc;
    SP_LAUNCH(s, k_scan_add, (unsigned)nb, SCAN_B, 0, data, len, tmp, gate);
    return SP_OK;
}

static int sp_exclusive_scan_gated(sp_system* s, int* data, long long len, ScanGate gate);
int sp_exclusive_scan_i32(sp_system* s, int* data, long long len) {
    return sp_exclusive_scan_gated(s, data, len, ScanGate{nullptr, 0});
}
static int sp_exclusive_scan_gated(sp_system* s, int* data, long long len, ScanGate gate) {
    if (len <= 0) return SP_OK;
    long long need = len / SCAN_TILE + 2048;
    if (need > s->scan_tmp_len) {
        if (s->scan_tmp) SP_CUDA(s, sp_dfree(s, s->scan_tmp));
        s->scan_tmp = nullptr;
        SP_CUDA(s, sp_dmalloc(&s->scan_tmp, (size_t)need * sizeof(int)));
        s->scan_tmp_len = need;
    }
    return scan_rec(s, data, len, s->scan_tmp, s->scan_tmp_len, gate);
}

// ------------------------------------------------------------------ cell list kernels
// Pass 1: domain test (core.jl:64-69), key (structs.jl:97-106), per-cell count and arrival offset.
// Particles outside the domain get the trash key key_max+1 and are counted in counters[SP_CNT_TRASH]; so do the slots of
// the dead tail [alive, n) — particles culled by an earlier build stay culled wherever their coordinates drift.
__global__ void k_cull_key(SpGrid g, const double* __restrict__ x, long long cap, long long n, int* __restrict__ key,
                           int* __restrict__ off, int* __restrict__ cell_count, int* __restrict__ counters) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int alive = counters[SP_CNT_ALIVE];
    int k = -1 - lane;  // inactive lanes: unique negative keys, no atomic
    if (s < n) {
        k = (int)g.key_max + 1;
        if (s < alive) {
            double px = x[s], py = x[cap + s], pz = x[2 * cap + s];
            if (sp_inside(g, px, py, pz)) k = (int)sp_find_key(g, px, py, pz);
            else if (g.slab_axis >= 0 && px == px) {
                // slab system: inside the global box but outside this rank's window = the particle outran the exchange
                // (more than the ghost width in one step); counted, and reported by a later rebuild (sp_slab.cu)
                bool in_box = true;
                const double p[3] = {px, py, pz};
                for (int a = 0; a < 3; a++)
                    if (a != g.slab_axis || !g.slab_periodic) in_box = in_box && g.lo[a] <= p[a] && p[a] <= g.hi[a];
                if (in_box) atomicAdd(&counters[SP_CNT_LOST], 1);
            }
        }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, k);
    const int leader = __ffs(peers) - 1;
    const int rank = __popc(peers & ((1u << lane) - 1u));
    int base = 0;
    if (lane == leader && k >= 0) {
        base = atomicAdd(&cell_count[k], __popc(peers));
        if (k == (int)g.key_max + 1) atomicAdd(&counters[SP_CNT_TRASH], __popc(peers));
    }
    base = __shfl_sync(peers, base, leader);
    if (s < n) {
        key[s] = k;
        off[s] = base + rank;
    }
}

// ---- removal renumbering (rare path), all arrays indexed by reference index r.
// Everything is decided on the device: alive = counters[SP_CNT_ALIVE] particles entered the build, of which
// n_out = counters[SP_CNT_TRASH] - (n - alive) were culled by it; every kernel leaves at once when n_out == 0.
__device__ __forceinline__ int sp_new_victims(const int* counters, long long n) {
    return counters[SP_CNT_TRASH] - (int)(n - counters[SP_CNT_ALIVE]);
}
__global__ void k_mark_victims(const int* key, const int* ref, int trash, int* V, int* Vscan, long long n,
                               const int* counters) {
    if (sp_new_victims(counters, n) == 0) return;
    const int alive = counters[SP_CNT_ALIVE];
    // grid-stride: the launch is a fixed few blocks per SM, so the common "nothing was culled" exit costs one launch
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (long long)gridDim.x * blockDim.x) {
        if (s < alive) {
            int v = key[s] == trash;
            V[ref[s]] = v;  // the references of the alive slots are a permutation of [0, alive)
            Vscan[ref[s]] = v;
        } else {
            V[s] = 0;
            Vscan[s] = 0;
        }
    }
}
// For a hole r (victim with r < n_new): i0 = number of victims with a larger index; the particle that lands
// in r is the one at tail position t = N-1-i0, following the chain while t is itself a victim
// (its content was overwritten earlier by the same rule).  Verified against the literal swap-with-tail loop of
// core.jl:72-81 in tests/test_oracle_pins.py::test_device_chain_rule_equals_the_literal_swap_with_tail_loop (host restatement of this rule) and on the device by the removal tests of
// tests/test_parity_gpu.py.
__global__ void k_chain(const int* V, const int* Vexcl, long long n, const int* counters, int* newref) {
    const int n_out = sp_new_victims(counters, n);
    if (n_out == 0) return;
    const long long N = counters[SP_CNT_ALIVE];
    const long long n_new = N - n_out;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_new; r += (long long)gridDim.x * blockDim.x) {
        if (!V[r]) continue;
        long long t = N - 1 - (n_out - (Vexcl[r] + 1));
        while (V[t]) t = N - 1 - (n_out - (Vexcl[t] + 1));
        newref[t] = (int)r;
    }
}
__global__ void k_apply_newref(const int* key, int trash, int* ref, const int* newref, long long n, const int* counters) {
    const int n_out = sp_new_victims(counters, n);
    if (n_out == 0) return;
    const long long alive = counters[SP_CNT_ALIVE];
    const long long n_new = alive - n_out;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < alive; s += (long long)gridDim.x * blockDim.x)
        if (key[s] != trash && ref[s] >= n_new) ref[s] = newref[ref[s]];
}
// the same three steps for small systems in ONE launch of one CTA (the launch count is what bounds the small configs)
__global__ void __launch_bounds__(1024) k_renumber_small(const int* key, int* ref, int trash, int* V, int* Vexcl, int* newref,
                                                         long long n, const int* counters) {
    const int n_out = sp_new_victims(counters, n);
    if (n_out == 0) return;
    __shared__ int part[1024];
    const int N = counters[SP_CNT_ALIVE];
    const int n_new = N - n_out;
    const int T = blockDim.x, t = threadIdx.x;
    for (int s = t; s < N; s += T) V[ref[s]] = (key[s] == trash);
    __syncthreads();
    // exclusive scan of V[0..N) in contiguous chunks, one per thread
    const int chunk = (N + T - 1) / T;
    const int b = min(t * chunk, N), e = min(b + chunk, N);
    int sum = 0;
    for (int r = b; r < e; r++) sum += V[r];
    part[t] = sum;
    __syncthreads();
    if (t == 0) {
        int run = 0;
        for (int u = 0; u < T; u++) {
            const int v = part[u];
            part[u] = run;
            run += v;
        }
    }
    __syncthreads();
    int run = part[t];
    for (int r = b; r < e; r++) {
        Vexcl[r] = run;
        run += V[r];
    }
    __syncthreads();
    for (int r = t; r < n_new; r += T) {
        if (!V[r]) continue;
        int q = N - 1 - (n_out - (Vexcl[r] + 1));
        while (V[q]) q = N - 1 - (n_out - (Vexcl[q] + 1));
        newref[q] = r;
    }
    __syncthreads();
    for (int s = t; s < N; s += T)
        if (key[s] != trash && ref[s] >= n_new) ref[s] = newref[ref[s]];
}

__global__ void k_scatter(const int* __restrict__ key, const int* __restrict__ off, const int* __restrict__ cell_start,
                          int* __restrict__ member, long long n) {
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) member[cell_start[key[s]] + off[s]] = (int)s;
}

// Order every cell by descending reference index: perm[cell_start + rank] = slot.  Only the n - trash kept slots; the
// trash cell (last in the scatter order) is left alone.
// On a slab system (gid != nullptr) the order inside a cell is by descending GLOBAL id instead — the reference numbering
// has no meaning across ranks, and an order that every rank agrees on makes a rank's boundary layers and its
// neighbour's ghost layers identical slot sequences (sp_slab.cu).
__global__ void k_rank(const int* __restrict__ key, const int* __restrict__ ref, const int* __restrict__ cell_start,
                       const int* __restrict__ member, int* __restrict__ perm, long long n, const int* __restrict__ counters,
                       const double* __restrict__ gid) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n - counters[SP_CNT_TRASH]) return;
    const int s = member[t];
    const int k = key[s];
    const int b = cell_start[k], e = cell_start[k + 1];
    int rank = 0;
    if (gid) {
        const double mine = gid[s];
        for (int u = b; u < e; u++) {
            const int o = member[u];
            const double g = gid[o];
            rank += (g > mine) || (g == mine && ref[o] > ref[s]);  // equal ids cannot happen; keep it a total order anyway
        }
    } else {
        const int mine = ref[s];
        for (int u = b; u < e; u++) rank += (ref[member[u]] > mine);
    }
    perm[b + rank] = s;
}
// close the build: publish the new alive count and the removal statistics
__global__ void k_finish_build(int* counters, long long n) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int alive_old = counters[SP_CNT_ALIVE];
        const int alive_new = (int)n - counters[SP_CNT_TRASH];
        counters[SP_CNT_CULLED] = alive_old - alive_new;
        counters[SP_CNT_REMOVED] += alive_old - alive_new;
        counters[SP_CNT_ALIVE] = alive_new;
        counters[SP_CNT_TRASH] = 0;  // ready for the next build (saves it a memset)
    }
}
__global__ void k_set_count(int* counters, int n) {
    if (threadIdx.x == 0 && blockIdx.x == 0) counters[SP_CNT_ALIVE] = n;
}

#define PERM_PLANES 40
struct PlaneTable {
    const double* in[PERM_PLANES];
    double* out[PERM_PLANES];
    int count;
    // the launch that moves the three position planes (planes 0-2 of the first table) also writes the FP32 cell-unit
    // coordinates u = (x - lo)/h of the pre-filter for the new slot order (sp_sweep.cu: k_prefilter_coords) — one pass
    // over the positions and one launch less per step
    float* ucoord;  // nullptr: not this launch
    long long ucap;
    float U;
};
__global__ void k_permute(SpGrid g, PlaneTable tab, const int* __restrict__ perm, const int* __restrict__ ref_in,
                          int* __restrict__ ref_out, const int* __restrict__ key_in, int* __restrict__ key_out,
                          long long n, const int* __restrict__ counters) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n - counters[SP_CNT_TRASH]) return;  // the kept slots; the dead tail keeps whatever it holds
    const int p = perm[t];
    if (ref_out) {
        ref_out[t] = ref_in[p];
        key_out[t] = key_in[p];
    }
    // 12 planes per trip: all gathers are issued before the first store (memory-level parallelism per thread)
    for (int c0 = 0; c0 < tab.count; c0 += 12) {
        double v[12];
#pragma unroll
        for (int u = 0; u < 12; u++)
            if (c0 + u < tab.count) v[u] = tab.in[c0 + u][p];
#pragma unroll
        for (int u = 0; u < 12; u++)
            if (c0 + u < tab.count) tab.out[c0 + u][t] = v[u];
        if (c0 == 0 && tab.ucoord) {  // same arithmetic as k_prefilter_coords
            const double ih = 1.0 / g.h;
            float a = (float)((v[0] - g.lo[0]) * ih), b = (float)((v[1] - g.lo[1]) * ih), c = (float)((v[2] - g.lo[2]) * ih);
            if (!(fabsf(a) <= tab.U) || !(fabsf(b) <= tab.U) || !(fabsf(c) <= tab.U)) a = b = c = nanf("");
            tab.ucoord[t] = a;
            tab.ucoord[tab.ucap + t] = b;
            tab.ucoord[2 * tab.ucap + t] = c;
        }
    }
}

int sp_permute_all(sp_system* s, long long n_keep) {
    const int B = 256;
    PlaneTable tab;
    tab.count = 0;
    tab.ucoord = nullptr;
    tab.ucap = 0;
    tab.U = 0.f;
    bool first = true;
    // the position planes are planes 0-2 of the first launch (field 0 is x and is never known-zero in a system with
    // particles in a domain; if it were, the separate pass of sp_sweep.cu computes the coordinates as before)
    const bool with_u = s->ucoord && s->ucoord_cap == s->cap && !s->fields[0].known_zero && !s->fields[0].transient;
    auto flush = [&]() -> int {
        if (tab.count == 0 && !first) return SP_OK;
        if (first && with_u) {
            tab.ucoord = s->ucoord;
            tab.ucap = s->cap;
            tab.U = sp_prefilter_range(s);
        } else
            tab.ucoord = nullptr;
        SP_LAUNCH(s, k_permute, sp_blocks(n_keep, B), B, 0, s->g, tab, s->perm, s->ref, first ? s->ref_alt : (int*)nullptr,
                  s->key, s->key_alt, n_keep, s->counters);
        first = false;
        tab.count = 0;
        return SP_OK;
    };
    // transient fields need not survive; a field that is +0.0 everywhere is its own permutation (no traffic, no swap)
    for (SpField& f : s->fields)
        for (int c = 0; c < f.ncomp && !f.transient && !f.known_zero; c++) {
            tab.in[tab.count] = f.d + (size_t)c * s->cap;
            tab.out[tab.count] = f.alt + (size_t)c * s->cap;
            if (++tab.count == PERM_PLANES) {
                int rc = flush();
                if (rc) return rc;
            }
        }
    int rc = flush();
    if (rc) return rc;
    for (SpField& f : s->fields)
        if (!f.transient && !f.known_zero) std::swap(f.d, f.alt);
    std::swap(s->ref, s->ref_alt);
    std::swap(s->key, s->key_alt);
    if (with_u) s->ucoord_version = s->x_version;  // valid until the next position write
    return SP_OK;
}

// Adopt a build's alive count if it has arrived on the host — never waits.
static void sp_adopt_count(sp_system* s) {
    if (s->n_exact || !s->count_pending || s->capturing) return;
    if (cudaEventQuery(s->ev_count) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    // valid: only builds lower the alive count, and every host-side change of the count settles first
    s->count_pending = false;
    s->n = s->h_counters[SP_CNT_ALIVE];
    s->n_removed = s->h_counters[SP_CNT_REMOVED];  // cumulative on the device
    s->last_culled = s->h_counters[SP_CNT_CULLED];
    s->n_exact = true;
}

int sp_settle(sp_system* s) {
    // every entry point that hands counts or data to the host comes through here: none of them can be part of a graph
    if (s->capturing)
        return sp_fail(s, SP_ERR_STATE, "this call hands data or counts to the host and cannot be recorded into a step graph");
    if (s->n_exact) return SP_OK;
    SP_CUDA(s, cudaMemcpyAsync(s->h_counters, s->counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    s->count_pending = false;
    s->n = s->h_counters[SP_CNT_ALIVE];
    s->n_removed = s->h_counters[SP_CNT_REMOVED];  // cumulative on the device
    s->last_culled = s->h_counters[SP_CNT_CULLED];
    s->n_exact = true;
    return SP_OK;
}

int sp_publish_count(sp_system* s) {
    SP_LAUNCH(s, k_set_count, 1, 32, 0, s->counters, (int)s->n);
    s->n_exact = true;
    s->count_pending = false;
    return SP_OK;
}

// create_cell_list! without a device->host read-back: the number of culled particles stays on the device, the kept
// slots are sorted by (key ascending, ref descending) into [0, alive) and the culled ones form the dead tail [alive, n).
int sp_build_cells(sp_system* s) {
    const SpGrid& g = s->g;
    s->x_version++;  // the slot order changes
    const int B = 256;
    sp_adopt_count(s);
    const long long N = s->n;
    const long long K = g.key_max;
    SP_CUDA(s, cudaMemsetAsync(s->cell_start, 0, (size_t)(K + 3) * sizeof(int), s->stream));
    // (counters[SP_CNT_TRASH] is zero between builds: k_finish_build leaves it so)
    if (N == 0) {
        s->have_cells = true;
        return SP_OK;
    }
    int* off = s->perm;  // arrival offsets live in perm until the scatter has consumed them
    SP_LAUNCH(s, k_cull_key, sp_blocks(N, B), B, 0, g, s->fields[0].d, s->cap, N, s->key, off, s->cell_start, s->counters);
    int rc = sp_exclusive_scan_i32(s, s->cell_start, K + 3);
    if (rc) return rc;
    SP_LAUNCH(s, k_scatter, sp_blocks(N, B), B, 0, s->key, off, s->cell_start, s->tmp_slot, N);
    if (g.slab_axis < 0) {
        // swap-with-tail renumbering of the reference indices (core.jl:72-81); no-ops unless this build culled particles
        int* V = s->flags;
        int* Vscan = s->key_alt;
        int* newref = s->ref_alt;
        if (N <= 65536) {
            SP_LAUNCH(s, k_renumber_small, 1, 1024, 0, s->key, s->ref, (int)K + 1, V, Vscan, newref, N, s->counters);
        } else {
            const unsigned G = (unsigned)std::min<long long>(sp_blocks(N, B), 148 * 8);
            SP_LAUNCH(s, k_mark_victims, G, B, 0, s->key, s->ref, (int)K + 1, V, Vscan, N, s->counters);
            if ((rc = sp_exclusive_scan_gated(s, Vscan, N, ScanGate{s->counters, N}))) return rc;
            SP_LAUNCH(s, k_chain, G, B, 0, V, Vscan, N, s->counters, newref);
            SP_LAUNCH(s, k_apply_newref, G, B, 0, s->key, (int)K + 1, s->ref, newref, N, s->counters);
        }
    }
    SP_LAUNCH(s, k_rank, sp_blocks(N, B), B, 0, s->key, s->ref, s->cell_start, s->tmp_slot, s->perm, N, s->counters,
              sp_slab_gid(s));
    rc = sp_permute_all(s, N);
    if (rc) return rc;
    SP_LAUNCH(s, k_finish_build, 1, 32, 0, s->counters, N);
    s->n_exact = false;
    if (!s->capturing) {
        SP_CUDA(s, cudaMemcpyAsync(s->h_counters, s->counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        SP_CUDA(s, cudaEventRecord(s->ev_count, s->stream));
        s->count_pending = true;
    }
    s->identity_order = false;
    s->have_cells = true;
    return SP_OK;
}

extern "C" int32_t sp_create_cell_list(sp_system* s) {
    if (!s) return SP_ERR_INVALID;
    SP_CUDA(s, cudaSetDevice(s->device));
    if (s->slab) return sp_fail(s, SP_ERR_STATE, "slab system: use sp_slab_create_cell_list");
    int rc = sp_time_begin(s);
    if (rc) return rc;
    if ((rc = sp_build_cells(s))) return rc;
    return sp_time_end(s);
}
