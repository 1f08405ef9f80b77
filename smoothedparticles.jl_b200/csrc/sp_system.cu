// sp_system.cu — handle life cycle, SoA field store, upload/download in reference order.
// Replaces ParticleSystem's constructor and storage (reference src/structs.jl:44-91, 118-125).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "sp_internal.cuh"

thread_local std::string g_sp_create_error;

int sp_fail(sp_system* s, int code, const std::string& msg) {
    if (s) s->err = msg;
    else g_sp_create_error = msg;
    return code;
}
int sp_fail_cuda(sp_system* s, cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    cudaGetLastError();
    return sp_fail(s, SP_ERR_CUDA, buf);
}

int sp_time_begin(sp_system* s) {
    if (s->in_program || s->capturing) return SP_OK;  // the step program times itself as one call; a recording is not timed
    SP_CUDA(s, cudaEventRecord(s->ev0, s->stream));
    return SP_OK;
}
int sp_time_end(sp_system* s) {
    if (s->in_program || s->capturing) return SP_OK;
    SP_CUDA(s, cudaEventRecord(s->ev1, s->stream));
    return SP_OK;
}

// ------------------------------------------------------------------ caching device allocator
namespace {
struct PoolState {
    bool tried = false, use_pool = false;
    cudaStream_t stream = nullptr;
};
PoolState g_pool[64];
std::once_flag g_pool_once[64];
PoolState& pool_for_current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    PoolState& ps = g_pool[dev & 63];
    // handles are not thread-safe, but two threads may create their first handles on one device at the same time
    std::call_once(g_pool_once[dev & 63], [&ps, dev]() {
        ps.tried = true;
        int supported = 0;
        cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, dev);
        const char* off = getenv("SP_NO_MEMPOOL");
        if (supported && !(off && atoi(off))) {
            cudaMemPool_t mp;
            unsigned long long keep = ~0ULL;
            if (cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess &&
                cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess &&
                cudaStreamCreateWithFlags(&ps.stream, cudaStreamNonBlocking) == cudaSuccess)
                ps.use_pool = true;
        }
        cudaGetLastError();
    });
    return ps;
}
}  // namespace

cudaError_t sp_dmalloc_impl(void** p, size_t bytes) {
    PoolState& ps = pool_for_current_device();
    if (bytes == 0) bytes = 1;
    if (!ps.use_pool) return cudaMalloc(p, bytes);
    cudaError_t e = cudaMallocAsync(p, bytes, ps.stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(ps.stream);  // complete before any other stream touches the block
}

cudaError_t sp_dfree_impl(void* p) {
    if (!p) return cudaSuccess;
    PoolState& ps = pool_for_current_device();
    if (!ps.use_pool) return cudaFree(p);
    return cudaFreeAsync(p, ps.stream);
}

int sp_ensure_stage(sp_system* s, long long doubles) {
    if (doubles <= s->stage_len) return SP_OK;
    long long want = doubles + doubles / 4 + 1024;
    if (s->stage) SP_CUDA(s, sp_dfree(s, s->stage));
    s->stage = nullptr;
    s->stage_len = 0;
    SP_CUDA(s, sp_dmalloc(&s->stage, (size_t)want * sizeof(double)));
    s->stage_len = want;
    return SP_OK;
}

template <class T>
static int regrow(sp_system* s, T** p, long long old_cap, long long new_cap, int planes, long long n_keep) {
    T* q = nullptr;
    SP_CUDA(s, sp_dmalloc(&q, (size_t)new_cap * planes * sizeof(T)));
    SP_CUDA(s, cudaMemsetAsync(q, 0, (size_t)new_cap * planes * sizeof(T), s->stream));
    if (*p) {
        for (int c = 0; c < planes && n_keep > 0; c++)
            SP_CUDA(s, cudaMemcpyAsync(q + (size_t)c * new_cap, *p + (size_t)c * old_cap, (size_t)n_keep * sizeof(T),
                                       cudaMemcpyDeviceToDevice, s->stream));
        SP_CUDA(s, cudaStreamSynchronize(s->stream));
        SP_CUDA(s, sp_dfree(s, *p));
    }
    *p = q;
    return SP_OK;
}

int sp_ensure_capacity(sp_system* s, long long n) {
    if (n <= s->cap) return SP_OK;
    long long nc = n + n / (s->slab ? 6 : 16) + 8;  // head room: arrivals on slab systems; the tile kernel's aligned
                                    // bulk copies may read one slot past n
    nc = (nc + 127) / 128 * 128;
    int rc;
    for (SpField& f : s->fields) {
        if ((rc = regrow(s, &f.d, s->cap, nc, f.ncomp, s->n))) return rc;
        if ((rc = regrow(s, &f.alt, s->cap, nc, f.ncomp, 0))) return rc;
    }
    if ((rc = regrow(s, &s->ref, s->cap, nc, 1, s->n))) return rc;
    if ((rc = regrow(s, &s->ref_alt, s->cap, nc, 1, 0))) return rc;
    if ((rc = regrow(s, &s->key, s->cap, nc, 1, s->n))) return rc;
    if ((rc = regrow(s, &s->key_alt, s->cap, nc, 1, 0))) return rc;
    if ((rc = regrow(s, &s->perm, s->cap, nc, 1, 0))) return rc;
    if ((rc = regrow(s, &s->tmp_slot, s->cap, nc, 1, 0))) return rc;
    if ((rc = regrow(s, &s->flags, s->cap, nc, 1, 0))) return rc;
    s->cap = nc;
    return SP_OK;
}

int sp_check_fields(sp_system* s, const int32_t* fields, int nfields, const int* ncomps, int nexpected) {
    if (!fields || nfields != nexpected) return sp_fail(s, SP_ERR_INVALID, "wrong number of fields for this operator");
    for (int i = 0; i < nfields; i++) {
        if (fields[i] < 0 || fields[i] >= (int)s->fields.size()) return sp_fail(s, SP_ERR_INVALID, "bad field id");
        if (ncomps[i] > 0 && s->fields[fields[i]].ncomp != ncomps[i])
            return sp_fail(s, SP_ERR_INVALID,
                           "field '" + s->fields[fields[i]].name + "' has the wrong number of components");
    }
    return SP_OK;
}

// ------------------------------------------------------------------ kernels
__global__ void k_iota_ref(int* ref, long long from, long long to) {
    long long i = from + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < to) ref[i] = (int)i;
}

// field planes <- staged host data (reference order)
__global__ void k_upload(double* d, long long cap, const double* stage, const int* ref, long long n, int ncomp,
                         int layout) {
    long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= n) return;
    long long r = ref[s];
    for (int c = 0; c < ncomp; c++)
        d[(size_t)c * cap + s] = (layout == SP_LAYOUT_AOS) ? stage[r * ncomp + c] : stage[(size_t)c * n + r];
}
__global__ void k_download(const double* d, long long cap, double* stage, const int* ref, long long n, int ncomp,
                           int layout) {
    long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= n) return;
    long long r = ref[s];
    for (int c = 0; c < ncomp; c++) {
        double v = d[(size_t)c * cap + s];
        if (layout == SP_LAYOUT_AOS) stage[r * ncomp + c] = v;
        else stage[(size_t)c * n + r] = v;
    }
}
__global__ void k_keys_by_ref(const int* key, const int* ref, long long* out, long long n) {
    long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s < n) out[ref[s]] = key[s];
}
__global__ void k_inverse_ref(const int* ref, int* perm, long long n) {
    long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s < n) perm[ref[s]] = (int)s;
}
__global__ void k_gather_plane(const double* in, double* out, const int* perm, long long n) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < n) out[t] = in[perm[t]];
}

// Put the slots back into reference order (slot == reference index).
static int restore_reference_order(sp_system* s) {
    if (s->identity_order || s->n == 0) {
        s->identity_order = true;
        return SP_OK;
    }
    const int B = 256;
    SP_LAUNCH(s, k_inverse_ref, sp_blocks(s->n, B), B, 0, s->ref, s->perm, s->n);
    for (SpField& f : s->fields) {
        for (int c = 0; c < f.ncomp; c++)
            SP_LAUNCH(s, k_gather_plane, sp_blocks(s->n, B), B, 0, f.d + (size_t)c * s->cap,
                      f.alt + (size_t)c * s->cap, s->perm, s->n);
        std::swap(f.d, f.alt);
    }
    SP_LAUNCH(s, k_iota_ref, sp_blocks(s->n, B), B, 0, s->ref, 0LL, s->n);
    s->identity_order = true;
    s->have_cells = false;
    s->x_version++;
    return SP_OK;
}

// ------------------------------------------------------------------ C ABI
extern "C" {

int32_t sp_version(void) { return SP_ABI_VERSION; }

const char* sp_last_error(const sp_system* sys) { return sys ? sys->err.c_str() : g_sp_create_error.c_str(); }

int32_t sp_device_count(int32_t* count) {
    if (!count) return SP_ERR_INVALID;
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return sp_fail(nullptr, SP_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
    }
    *count = c;
    return SP_OK;
}

int32_t sp_create(sp_system** out, const double lo[3], const double hi[3], double h, int32_t device) {
    if (!out || !lo || !hi) return sp_fail(nullptr, SP_ERR_INVALID, "null argument");
    *out = nullptr;
    // structs.jl:59  @assert(h > 0.0)
    if (!(h > 0.0)) return sp_fail(nullptr, SP_ERR_INVALID, "invalid ParticleSystem declaration! (h must be a positive float)");
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        return sp_fail(nullptr, SP_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= cnt) return sp_fail(nullptr, SP_ERR_INVALID, "bad device ordinal");
    sp_system* s = new sp_system();
    s->device = device;
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        int rc = sp_fail_cuda(nullptr, e, "cudaSetDevice", __FILE__, __LINE__);
        delete s;
        return rc;
    }
    SpGrid& g = s->g;
    g.h = h;
    for (int a = 0; a < 3; a++) {
        g.lo[a] = lo[a];
        g.hi[a] = hi[a];
        g.phase[a] = (long long)std::floor(lo[a] / h);                    // structs.jl:66
        g.lim[a] = (long long)std::floor(hi[a] / h) - g.phase[a] + 1;    // structs.jl:67
        if (g.lim[a] < 1) {
            delete s;
            return sp_fail(nullptr, SP_ERR_INVALID, "empty domain box");
        }
    }
    g.key_max = g.lim[0] * g.lim[1] * g.lim[2];  // structs.jl:68
    if (g.key_max >= (1LL << 31) - 8) {
        delete s;
        return sp_fail(nullptr, SP_ERR_INVALID, "key_max exceeds the 31-bit cell index of this build");
    }
    g.dim = (g.lim[2] == 1) ? 2 : 3;  // structs.jl:70
    g.slab_axis = -1;
    g.slab_periodic = 0;
    s->n_key_diff = 0;
    if (g.dim == 2) {
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++) s->key_diff[s->n_key_diff++] = di + g.lim[0] * dj;  // :73-75
    } else {
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++)
                for (int dk = -1; dk <= 1; dk++)
                    s->key_diff[s->n_key_diff++] = di + g.lim[0] * (dj + g.lim[1] * dk);  // :79-81
    }
    // (r > h) <=> (d2 > T2) with r = sqrt_rn(d2): sqrt_rn is monotone, so find the largest t with sqrt(t) <= h.
    double t = h * h;
    while (std::sqrt(t) <= h) t = std::nextafter(t, INFINITY);
    while (std::sqrt(t) > h) t = std::nextafter(t, 0.0);
    g.T2 = t;

#define CREATE_TRY(call)                                                         \
    do {                                                                         \
        cudaError_t _e = (call);                                                 \
        if (_e != cudaSuccess) {                                                 \
            int rc = sp_fail_cuda(nullptr, _e, #call, __FILE__, __LINE__);      \
            sp_destroy(s);                                                       \
            return rc;                                                           \
        }                                                                        \
    } while (0)
    CREATE_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaEventCreate(&s->ev0));
    CREATE_TRY(cudaEventCreate(&s->ev1));
    CREATE_TRY(cudaEventCreate(&s->tev0));
    CREATE_TRY(cudaEventCreate(&s->tev1));
    CREATE_TRY(cudaEventCreateWithFlags(&s->ev_count, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&s->ev_nbr, cudaEventDisableTiming));
    CREATE_TRY(sp_dmalloc(&s->cell_start, (size_t)(g.key_max + 3) * sizeof(int)));
    CREATE_TRY(sp_dmalloc(&s->cell_fill, (size_t)(g.key_max + 3) * sizeof(int)));
    CREATE_TRY(cudaMemset(s->cell_start, 0, (size_t)(g.key_max + 3) * sizeof(int)));
    CREATE_TRY(sp_dmalloc(&s->counters, 64 * sizeof(int)));
    CREATE_TRY(cudaMemset(s->counters, 0, 64 * sizeof(int)));
    CREATE_TRY(cudaHostAlloc(&s->h_counters, 64 * sizeof(int), cudaHostAllocDefault));
    s->scan_tmp_len = (g.key_max + 3) / 1024 + 1024;
    CREATE_TRY(sp_dmalloc(&s->scan_tmp, (size_t)s->scan_tmp_len * sizeof(int)));
#undef CREATE_TRY
    SpField fx;
    fx.name = "x";
    fx.ncomp = 3;
    s->fields.push_back(fx);
    *out = s;
    return SP_OK;
}

int32_t sp_destroy(sp_system* s) {
    if (!s) return SP_OK;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    sp_program_free(s);
    sp_slab_free(s);
    for (SpField& f : s->fields) {
        sp_dfree(s, f.d);
        sp_dfree(s, f.alt);
    }
    sp_dfree(s, s->ref);
    sp_dfree(s, s->ref_alt);
    sp_dfree(s, s->key);
    sp_dfree(s, s->key_alt);
    sp_dfree(s, s->cell_start);
    sp_dfree(s, s->cell_fill);
    sp_dfree(s, s->perm);
    sp_dfree(s, s->tmp_slot);
    sp_dfree(s, s->flags);
    sp_dfree(s, s->scan_tmp);
    sp_dfree(s, s->counters);
    sp_dfree(s, s->stage);
    sp_dfree(s, s->dscal);
    sp_dfree(s, s->ucoord);
    sp_dfree(s, s->nbr_ids);
    sp_dfree(s, s->ell_val);
    sp_dfree(s, s->nbr_cnt);
    if (s->h_scal) cudaFreeHost(s->h_scal);
    if (s->h_counters) cudaFreeHost(s->h_counters);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->tev0) cudaEventDestroy(s->tev0);
    if (s->tev1) cudaEventDestroy(s->tev1);
    if (s->ev_count) cudaEventDestroy(s->ev_count);
    if (s->ev_nbr) cudaEventDestroy(s->ev_nbr);
    if (s->stream) cudaStreamDestroy(s->stream);
    cudaGetLastError();
    delete s;
    return SP_OK;
}

int32_t sp_key_params(const sp_system* s, int64_t key_phase[3], int64_t key_lim[3], int64_t* key_max,
                      int32_t* n_key_diff, int64_t key_diff[27]) {
    if (!s) return SP_ERR_INVALID;
    for (int a = 0; a < 3; a++) {
        if (key_phase) key_phase[a] = s->g.phase[a];
        if (key_lim) key_lim[a] = s->g.lim[a];
    }
    if (key_max) *key_max = s->g.key_max;
    if (n_key_diff) *n_key_diff = s->n_key_diff;
    if (key_diff)
        for (int i = 0; i < s->n_key_diff; i++) key_diff[i] = s->key_diff[i];
    return SP_OK;
}

int32_t sp_add_field(sp_system* s, const char* name, int32_t ncomp, int32_t* fid) {
    if (!s || !name || !fid) return SP_ERR_INVALID;
    if (ncomp != 1 && ncomp != 3 && ncomp != 9) return sp_fail(s, SP_ERR_INVALID, "ncomp must be 1, 3 or 9");
    SP_CUDA(s, cudaSetDevice(s->device));
    for (size_t i = 0; i < s->fields.size(); i++)
        if (s->fields[i].name == name) {
            if (s->fields[i].ncomp != ncomp) return sp_fail(s, SP_ERR_INVALID, "field exists with another ncomp");
            *fid = (int32_t)i;
            return SP_OK;
        }
    if (s->fields.size() >= SP_MAX_FIELDS) return sp_fail(s, SP_ERR_INVALID, "too many fields");
    SpField f;
    f.name = name;
    f.ncomp = ncomp;
    if (s->cap > 0) {
        size_t bytes = (size_t)s->cap * ncomp * sizeof(double);
        SP_CUDA(s, sp_dmalloc(&f.d, bytes));
        SP_CUDA(s, sp_dmalloc(&f.alt, bytes));
        SP_CUDA(s, cudaMemsetAsync(f.d, 0, bytes, s->stream));
    }
    s->fields.push_back(f);
    *fid = (int32_t)s->fields.size() - 1;
    return SP_OK;
}

int32_t sp_find_field(const sp_system* s, const char* name, int32_t* fid) {
    if (!s || !name || !fid) return SP_ERR_INVALID;
    for (size_t i = 0; i < s->fields.size(); i++)
        if (s->fields[i].name == name) {
            *fid = (int32_t)i;
            return SP_OK;
        }
    return SP_ERR_INVALID;
}

int32_t sp_resize(sp_system* s, int64_t n) {
    if (!s || n < 0) return SP_ERR_INVALID;
    if (n >= (1LL << 31) - 256) return sp_fail(s, SP_ERR_INVALID, "particle count exceeds 31 bits");
    SP_CUDA(s, cudaSetDevice(s->device));
    int rc = sp_settle(s);
    if (rc) return rc;
    if (n == s->n) return SP_OK;
    sp_slab_host_touched(s);
    s->x_version++;
    if (n < s->n) {
        rc = restore_reference_order(s);
        if (rc) return rc;
        s->n = n;
        s->have_cells = false;
        return sp_publish_count(s);
    }
    rc = sp_ensure_capacity(s, n);
    if (rc) return rc;
    for (SpField& f : s->fields)
        for (int c = 0; c < f.ncomp; c++)
            SP_CUDA(s, cudaMemsetAsync(f.d + (size_t)c * s->cap + s->n, 0, (size_t)(n - s->n) * sizeof(double), s->stream));
    SP_LAUNCH(s, k_iota_ref, sp_blocks(n - s->n, 256), 256, 0, s->ref, (long long)s->n, (long long)n);
    s->n = n;
    s->have_cells = false;
    return sp_publish_count(s);
}

int32_t sp_num_particles(sp_system* s, int64_t* n) {
    if (!s || !n) return SP_ERR_INVALID;
    SP_CUDA(s, cudaSetDevice(s->device));
    int rc = sp_settle(s);
    if (rc) return rc;
    *n = s->n;
    return SP_OK;
}

int32_t sp_num_removed(sp_system* s, int64_t* n_removed) {
    if (!s || !n_removed) return SP_ERR_INVALID;
    SP_CUDA(s, cudaSetDevice(s->device));
    int rc = sp_settle(s);
    if (rc) return rc;
    *n_removed = s->n_removed;
    return SP_OK;
}

int32_t sp_upload(sp_system* s, int32_t fid, const double* host, int64_t n, int32_t layout) {
    if (!s || !host) return SP_ERR_INVALID;
    if (fid < 0 || fid >= (int)s->fields.size()) return sp_fail(s, SP_ERR_INVALID, "bad field id");
    SP_CUDA(s, cudaSetDevice(s->device));
    if (int rcs = sp_settle(s)) return rcs;
    if (n != s->n) return sp_fail(s, SP_ERR_INVALID, "upload: n differs from the particle count");
    if (layout != SP_LAYOUT_AOS && layout != SP_LAYOUT_SOA) return sp_fail(s, SP_ERR_INVALID, "bad layout");
    if (n == 0) return SP_OK;
    SP_CUDA(s, cudaSetDevice(s->device));
    SpField& f = s->fields[fid];
    if (fid == 0) sp_slab_host_touched(s);
    sp_wrote(s, fid);
    int rc = sp_time_begin(s);
    if (rc) return rc;
    if (s->identity_order && layout == SP_LAYOUT_SOA) {
        for (int c = 0; c < f.ncomp; c++)
            SP_CUDA(s, cudaMemcpyAsync(f.d + (size_t)c * s->cap, host + (size_t)c * n, (size_t)n * sizeof(double),
                                       cudaMemcpyHostToDevice, s->stream));
    } else {
        if ((rc = sp_ensure_stage(s, n * f.ncomp))) return rc;
        SP_CUDA(s, cudaMemcpyAsync(s->stage, host, (size_t)n * f.ncomp * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        SP_LAUNCH(s, k_upload, sp_blocks(n, 256), 256, 0, f.d, s->cap, s->stage, s->ref, (long long)n, f.ncomp, layout);
    }
    if ((rc = sp_time_end(s))) return rc;
    // the host buffer is only borrowed for the duration of the call
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    return SP_OK;
}

int32_t sp_download(sp_system* s, int32_t fid, double* host, int64_t n, int32_t layout) {
    if (!s || !host) return SP_ERR_INVALID;
    if (fid < 0 || fid >= (int)s->fields.size()) return sp_fail(s, SP_ERR_INVALID, "bad field id");
    SP_CUDA(s, cudaSetDevice(s->device));
    if (int rcs = sp_settle(s)) return rcs;
    if (n != s->n) return sp_fail(s, SP_ERR_INVALID, "download: n differs from the particle count");
    if (layout != SP_LAYOUT_AOS && layout != SP_LAYOUT_SOA) return sp_fail(s, SP_ERR_INVALID, "bad layout");
    if (n == 0) return SP_OK;
    SP_CUDA(s, cudaSetDevice(s->device));
    SpField& f = s->fields[fid];
    int rc = sp_time_begin(s);
    if (rc) return rc;
    if (s->identity_order && layout == SP_LAYOUT_SOA) {
        for (int c = 0; c < f.ncomp; c++)
            SP_CUDA(s, cudaMemcpyAsync(host + (size_t)c * n, f.d + (size_t)c * s->cap, (size_t)n * sizeof(double),
                                       cudaMemcpyDeviceToHost, s->stream));
    } else {
        if ((rc = sp_ensure_stage(s, n * f.ncomp))) return rc;
        SP_LAUNCH(s, k_download, sp_blocks(n, 256), 256, 0, f.d, s->cap, s->stage, s->ref, (long long)n, f.ncomp, layout);
        SP_CUDA(s, cudaMemcpyAsync(host, s->stage, (size_t)n * f.ncomp * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    }
    if ((rc = sp_time_end(s))) return rc;
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    return SP_OK;
}

int32_t sp_synchronize(sp_system* s) {
    if (!s) return SP_ERR_INVALID;
    SP_NOT_WHILE_RECORDING(s);
    SP_CUDA(s, cudaSetDevice(s->device));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    return SP_OK;
}

int32_t sp_last_call_ms(sp_system* s, float* ms) {
    if (!s || !ms) return SP_ERR_INVALID;
    SP_NOT_WHILE_RECORDING(s);
    SP_CUDA(s, cudaSetDevice(s->device));
    SP_CUDA(s, cudaEventSynchronize(s->ev1));
    SP_CUDA(s, cudaEventElapsedTime(ms, s->ev0, s->ev1));
    return SP_OK;
}

int32_t sp_timer_start(sp_system* s) {
    if (!s) return SP_ERR_INVALID;
    SP_NOT_WHILE_RECORDING(s);
    SP_CUDA(s, cudaSetDevice(s->device));
    SP_CUDA(s, cudaEventRecord(s->tev0, s->stream));
    return SP_OK;
}
int32_t sp_timer_stop(sp_system* s, float* ms) {
    if (!s || !ms) return SP_ERR_INVALID;
    SP_NOT_WHILE_RECORDING(s);
    SP_CUDA(s, cudaSetDevice(s->device));
    SP_CUDA(s, cudaEventRecord(s->tev1, s->stream));
    SP_CUDA(s, cudaEventSynchronize(s->tev1));
    SP_CUDA(s, cudaEventElapsedTime(ms, s->tev0, s->tev1));
    return SP_OK;
}

int32_t sp_launch_count(sp_system* s, int64_t* launches) {
    if (!s || !launches) return SP_ERR_INVALID;
    *launches = s->launches;
    return SP_OK;
}

int32_t sp_get_cell_keys(sp_system* s, int64_t* keys, int64_t n) {
    if (!s || !keys) return SP_ERR_INVALID;
    if (!s->have_cells) return sp_fail(s, SP_ERR_STATE, "no cell list: call sp_create_cell_list first");
    SP_CUDA(s, cudaSetDevice(s->device));
    if (int rcs = sp_settle(s)) return rcs;
    if (n != s->n) return sp_fail(s, SP_ERR_INVALID, "n differs from the particle count");
    if (n == 0) return SP_OK;
    SP_CUDA(s, cudaSetDevice(s->device));
    int rc = sp_ensure_stage(s, n);
    if (rc) return rc;
    SP_LAUNCH(s, k_keys_by_ref, sp_blocks(n, 256), 256, 0, s->key, s->ref, (long long*)s->stage, (long long)n);
    SP_CUDA(s, cudaMemcpyAsync(keys, s->stage, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, s->stream));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    return SP_OK;
}

int32_t sp_get_cell_list(sp_system* s, int64_t* offsets, int64_t* members) {
    if (!s || !offsets || !members) return SP_ERR_INVALID;
    if (!s->have_cells) return sp_fail(s, SP_ERR_STATE, "no cell list: call sp_create_cell_list first");
    SP_CUDA(s, cudaSetDevice(s->device));
    if (int rcs = sp_settle(s)) return rcs;
    const long long K = s->g.key_max;
    std::vector<int> cs(K + 1), rf(s->n);
    SP_CUDA(s, cudaMemcpyAsync(cs.data(), s->cell_start + 1, (size_t)(K + 1) * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    if (s->n)
        SP_CUDA(s, cudaMemcpyAsync(rf.data(), s->ref, (size_t)s->n * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    for (long long k = 0; k <= K; k++) offsets[k] = cs[k];
    for (long long t = 0; t < s->n; t++) members[t] = (int64_t)rf[t] + 1;
    return SP_OK;
}

}  // extern "C"
