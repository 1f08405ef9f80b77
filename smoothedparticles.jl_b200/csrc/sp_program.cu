// sp_program.cu — fused step programs: the time loops of the examples issued from inside the library,
// so a Julia/Python host pays one FFI crossing per batch of steps instead of 7-9 per step.
// Same kernels, same order, same arithmetic as the per-call path.
#include <cstdint>

#include "sp_internal.cuh"

int sp_apply_impl(sp_system* s, int32_t op, const int32_t* F, int32_t nf, const double* Pm, int32_t np, int32_t flags);
int sp_build_cells(sp_system* s);
int sp_kick_kick_move_impl(sp_system* s, const int32_t* F, const double* Pm);
int sp_find_pressure_pr_impl(sp_system* s, const int32_t* F, const double* Pm);

// fields {x, v, Dv, rho, Drho, P, type}; params {kernel, m, h, two_nu, dt, c2, rho0, mu, gx, gy, gz}
//
// CUDA graphs.  Since the cell-list build keeps its counts on the device (sp_cells.cu) a time step is a fixed sequence of
// launches with no device->host read-back, so a UNIT of two steps (two builds: every ping-pong pair of planes is back
// where it started) is captured once into a CUDA graph and replayed: the small configs are bound by the launch
// overhead of their ~25 kernels per step, not by the kernels.  The graph is replayed only while everything its
// launches depend on is unchanged — plane pointers, slot bound, list capacity, operator parameters — which is
// checked through a signature before every replay; otherwise the unit is captured again (or run eagerly).
// Particles that leave the domain inside a replayed unit are culled on the device exactly as in the eager path (dead
// tail); only the launch width stays at the bound it was captured with.
// SP_GRAPH=0 disables the graphs.
static int run_step(sp_system* s, int32_t program, const int32_t* F, const double* P, bool first, bool last);

static unsigned long long mix(unsigned long long h, unsigned long long v) {
    h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
    return h;
}
static unsigned long long state_signature(sp_system* s) {
    unsigned long long h = 1469598103934665603ULL;
    h = mix(h, (unsigned long long)s->n);
    h = mix(h, (unsigned long long)s->cap);
    for (const SpField& f : s->fields) {
        h = mix(h, (unsigned long long)(uintptr_t)f.d);
        h = mix(h, (unsigned long long)(uintptr_t)f.alt);
        h = mix(h, f.known_zero ? 2 : 1);
    }
    h = mix(h, (unsigned long long)(uintptr_t)s->ref);
    h = mix(h, (unsigned long long)(uintptr_t)s->key);
    h = mix(h, (unsigned long long)(uintptr_t)s->nbr_ids);
    h = mix(h, (unsigned long long)(uintptr_t)s->nbr_cnt);
    h = mix(h, (unsigned long long)s->nbr_capk);
    h = mix(h, (unsigned long long)(uintptr_t)s->ucoord);
    h = mix(h, (unsigned long long)(uintptr_t)s->stage);
    h = mix(h, (unsigned long long)(uintptr_t)s->scan_tmp);
    h = mix(h, s->pair_aux.valid ? 2 : 1);
    return h;
}
static bool graphs_enabled() {
    static const bool on = !(getenv("SP_GRAPH") && atoi(getenv("SP_GRAPH")) == 0);
    return on;
}
static void drop_graph(sp_system* s) {
    if (s->graph.exec) cudaGraphExecDestroy(s->graph.exec);
    s->graph.exec = nullptr;
    s->graph.sig = 0;
}
void sp_program_free(sp_system* s) {
    drop_graph(s);
    for (auto& g : s->user_graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    s->user_graphs.clear();
}

// ------------------------------------------------------------------ host-recorded graphs
extern "C" int32_t sp_graph_begin(sp_system* s) {
    if (!s) return SP_ERR_INVALID;
    if (s->capturing) return sp_fail(s, SP_ERR_STATE, "a recording is already in progress");
    if (s->slab) return sp_fail(s, SP_ERR_STATE, "step graphs are not available on a slab system");
    SP_CUDA(s, cudaSetDevice(s->device));
    s->rec_sig = state_signature(s);
    s->rec_launches = s->launches;
    SP_CUDA(s, cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeRelaxed));
    s->capturing = true;
    s->recording = true;
    return SP_OK;
}

extern "C" int32_t sp_graph_end(sp_system* s, int32_t* graph_id) {
    if (!s || !graph_id) return SP_ERR_INVALID;
    *graph_id = -1;
    if (!s->recording) return sp_fail(s, SP_ERR_STATE, "sp_graph_end without sp_graph_begin");
    SP_CUDA(s, cudaSetDevice(s->device));
    s->capturing = false;
    s->recording = false;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
    if (e != cudaSuccess || !graph) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        return sp_fail(s, SP_ERR_CUDA, std::string("recording failed: ") + cudaGetErrorString(e) +
                                            " — the recorded calls were NOT executed; the system's state is undefined");
    }
    const bool periodic = state_signature(s) == s->rec_sig;
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess || !exec) {
        cudaGetLastError();
        return sp_fail(s, SP_ERR_CUDA, "cudaGraphInstantiate failed — the recorded calls were NOT executed");
    }
    // the body runs once now: begin/end behaves like the calls it encloses
    SP_CUDA(s, cudaGraphLaunch(exec, s->stream));
    s->n_exact = false;
    s->count_pending = false;
    if (!periodic) {
        cudaGraphExecDestroy(exec);
        return sp_fail(s, SP_ERR_STATE,
                       "the recorded body was executed once but cannot be replayed: it does not leave the ping-pong buffers "
                       "where it found them (record an even number of cell-list builds, e.g. two time steps)");
    }
    sp_system::StepGraph g;
    g.exec = exec;
    g.sig = s->rec_sig;
    g.launches = s->launches - s->rec_launches;
    int id = -1;
    for (size_t i = 0; i < s->user_graphs.size(); i++)
        if (!s->user_graphs[i].exec) id = (int)i;
    if (id < 0) {
        s->user_graphs.push_back(g);
        id = (int)s->user_graphs.size() - 1;
    } else
        s->user_graphs[id] = g;
    *graph_id = id;
    return SP_OK;
}

extern "C" int32_t sp_graph_launch(sp_system* s, int32_t graph_id, int64_t times) {
    if (!s || times < 0) return SP_ERR_INVALID;
    if (graph_id < 0 || graph_id >= (int)s->user_graphs.size() || !s->user_graphs[graph_id].exec)
        return sp_fail(s, SP_ERR_INVALID, "unknown graph id");
    if (s->capturing) return sp_fail(s, SP_ERR_STATE, "sp_graph_launch inside a recording");
    SP_CUDA(s, cudaSetDevice(s->device));
    sp_system::StepGraph& g = s->user_graphs[graph_id];
    if (g.sig != state_signature(s))
        return sp_fail(s, SP_ERR_STATE, "the system changed since this graph was recorded (field storage, particle bound or "
                                        "list capacity): record it again");
    int rc = sp_time_begin(s);
    if (rc) return rc;
    for (int64_t k = 0; k < times; k++) {
        SP_CUDA(s, cudaGraphLaunch(g.exec, s->stream));
        s->launches += g.launches;
    }
    if (times > 0) {
        s->n_exact = false;
        s->count_pending = false;
    }
    return sp_time_end(s);
}

extern "C" int32_t sp_graph_destroy(sp_system* s, int32_t graph_id) {
    if (!s) return SP_ERR_INVALID;
    if (graph_id < 0 || graph_id >= (int)s->user_graphs.size() || !s->user_graphs[graph_id].exec)
        return sp_fail(s, SP_ERR_INVALID, "unknown graph id");
    SP_CUDA(s, cudaSetDevice(s->device));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    cudaGraphExecDestroy(s->user_graphs[graph_id].exec);
    s->user_graphs[graph_id].exec = nullptr;
    return SP_OK;
}

// capture `unit` middle steps; on success the graph is instantiated and NOT yet launched (the capture executes nothing)
static int capture_unit(sp_system* s, int32_t program, const int32_t* F, const double* P, int unit, bool* ok) {
    *ok = false;
    drop_graph(s);
    const unsigned long long sig0 = state_signature(s);
    const long long launches0 = s->launches;
    cudaGraph_t graph = nullptr;
    SP_CUDA(s, cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeRelaxed));
    s->capturing = true;
    int rc = SP_OK;
    for (int k = 0; k < unit && !rc; k++) rc = run_step(s, program, F, P, false, false);
    s->capturing = false;
    cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
    if (rc) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (e != cudaSuccess || !graph) {
        // the host bookkeeping of the unit has run but none of its launches: the state cannot be trusted any more
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        return sp_fail(s, SP_ERR_CUDA, std::string("step-graph capture failed: ") + cudaGetErrorString(e) +
                                            " (set SP_GRAPH=0 to run the step programs without CUDA graphs)");
    }
    // the captured launches did not execute, but the host-side bookkeeping of the unit did: versions, ping-pong swaps.
    // The unit is replayable iff that bookkeeping is back where it started.
    const bool periodic = state_signature(s) == sig0;
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess || !exec) {
        cudaGetLastError();
        return sp_fail(s, SP_ERR_CUDA, "cudaGraphInstantiate failed after a capture whose launches were skipped");
    }
    s->graph.exec = exec;
    s->graph.sig = periodic ? sig0 : 0;
    s->graph.program = program;
    s->graph.unit = unit;
    s->graph.launches = s->launches - launches0;
    for (int i = 0; i < 11; i++) s->graph.params[i] = P[i];
    for (int i = 0; i < 7; i++) s->graph.fields[i] = F[i];
    *ok = true;
    return SP_OK;
}

static bool graph_matches(sp_system* s, int32_t program, const int32_t* F, const double* P) {
    if (!s->graph.exec || !s->graph.sig || s->graph.program != program) return false;
    for (int i = 0; i < 11; i++)
        if (s->graph.params[i] != P[i]) return false;
    for (int i = 0; i < 7; i++)
        if (s->graph.fields[i] != F[i]) return false;
    return s->graph.sig == state_signature(s);
}

static int run_program(sp_system* s, int32_t program, const int32_t* F, int32_t nf, const double* P, int32_t np,
                       int64_t nsteps);

extern "C" int32_t sp_run_program(sp_system* s, int32_t program, const int32_t* F, int32_t nf, const double* P,
                                  int32_t np, int64_t nsteps) {
    if (!s) return SP_ERR_INVALID;
    SP_CUDA(s, cudaSetDevice(s->device));
    int rc = sp_time_begin(s);
    if (rc) return rc;
    s->in_program = true;
    rc = run_program(s, program, F, nf, P, np, nsteps);
    s->in_program = false;
    if (rc) return rc;
    return sp_time_end(s);
}

// One time step.  `first`: the step opens a run (3-D: the run's first move! is its own kernel); `last`: it closes one
// (3-D: the two accelerate! are not fused with the next step's move!).
// On a slab system (sp_slab.cu) the cell-list build is the slab rebuild (one exchange: migration + two ghost layers per
// side + local build).
static int run_step(sp_system* s, int32_t program, const int32_t* F, const double* P, bool first, bool last) {
    int rc;
    const bool slab = s->slab != nullptr;
    const int32_t x = F[0], v = F[1], Dv = F[2], rho = F[3], Drho = F[4], Pr = F[5], ty = F[6];
    const double kernel = P[0], m = P[1], h = P[2], two_nu = P[3], dt = P[4], c2 = P[5], rho0 = P[6], mu = P[7];
    const int32_t f_bom[4] = {x, v, rho, Drho}, f_fp[3] = {rho, Drho, Pr}, f_if[6] = {x, v, Pr, rho, Dv, ty},
                  f_mv[4] = {x, v, Dv, ty}, f_ac[3] = {v, Dv, ty};
    const double p_bom[4] = {kernel, m, h, two_nu}, p_fp[4] = {dt, c2, rho0, 0.0}, p_if[5] = {kernel, m, h, mu, rho0},
                 p_ac[4] = {0.5 * dt, P[8], P[9], P[10]};
#define STEP(call) \
    if ((rc = (call))) return rc;
    if (program == SP_PROGRAM_WCSPH_3D) {  // examples/collapse3d.jl:136-150
        // same statements in the same order as the per-call loop; the unary passes that touch the same
        // fields back to back are issued as one kernel each (bit-identical, tested against the per-call path):
        //   accelerate!, accelerate! of step k with move! of step k+1;  find_pressure! with the P/rho^2 pass
        const double p_mv[1] = {dt};
        const int32_t f_kkm[4] = {v, Dv, x, ty};
        const double p_kkm[5] = {0.5 * dt, P[8], P[9], P[10], dt};
        if (first) STEP(sp_apply_impl(s, SP_OP_MOVE, f_mv, 4, p_mv, 1, 0));
        {
            // P is dead here: find_pressure! rewrites it for every particle before anything reads it (collapse3d.jl:140-141),
            // so the build need not carry it along (one plane less to permute, and to send on a slab system)
            const bool was = s->fields[Pr].transient;
            s->fields[Pr].transient = true;
            rc = slab ? sp_slab_create_cell_list(s) : sp_build_cells(s);
            s->fields[Pr].transient = was;
            if (rc) return rc;
        }
        STEP(sp_apply_impl(s, SP_OP_BALANCE_OF_MASS, f_bom, 4, p_bom, 4, 0));
        // (slab systems: the two ghost layers per side make the inner one integrate its own density — same neighbours,
        // same order as on its owner — so rho and P of every ghost a force sum reads are already right: no refresh)
        STEP(sp_find_pressure_pr_impl(s, f_fp, p_fp));
        STEP(sp_apply_impl(s, SP_OP_INTERNAL_FORCE, f_if, 6, p_if, 5, SP_FLAG_INTERNAL_PR_READY));
        if (!last) {
            STEP(sp_kick_kick_move_impl(s, f_kkm, p_kkm));
        } else {
            STEP(sp_apply_impl(s, SP_OP_ACCELERATE, f_ac, 3, p_ac, 4, 0));
            STEP(sp_apply_impl(s, SP_OP_ACCELERATE, f_ac, 3, p_ac, 4, 0));
        }
    } else if (program == SP_PROGRAM_WCSPH_2D) {  // examples/collapse_dry.jl:203-211
        const double p_mv[1] = {0.5 * dt};
        STEP(sp_apply_impl(s, SP_OP_ACCELERATE, f_ac, 3, p_ac, 4, 0));
        STEP(sp_apply_impl(s, SP_OP_MOVE, f_mv, 4, p_mv, 1, 0));
        STEP(sp_build_cells(s));
        STEP(sp_apply_impl(s, SP_OP_BALANCE_OF_MASS, f_bom, 4, p_bom, 4, 0));
        STEP(sp_apply_impl(s, SP_OP_FIND_PRESSURE, f_fp, 3, p_fp, 4, 0));
        STEP(sp_apply_impl(s, SP_OP_MOVE, f_mv, 4, p_mv, 1, 0));
        STEP(sp_build_cells(s));
        STEP(sp_apply_impl(s, SP_OP_INTERNAL_FORCE, f_if, 6, p_if, 5, 0));
        STEP(sp_apply_impl(s, SP_OP_ACCELERATE, f_ac, 3, p_ac, 4, 0));
    } else
        return sp_fail(s, SP_ERR_INVALID, "unknown program id");
#undef STEP
    return SP_OK;
}

static int run_program(sp_system* s, int32_t program, const int32_t* F, int32_t nf, const double* P, int32_t np,
                       int64_t nsteps) {
    const int nc[] = {3, 3, 3, 1, 1, 1, 1};
    int rc = sp_check_fields(s, F, nf, nc, 7);
    if (rc) return rc;
    if (np != 11 || !P) return sp_fail(s, SP_ERR_INVALID, "wrong number of parameters for this program");
    if (program != SP_PROGRAM_WCSPH_3D && program != SP_PROGRAM_WCSPH_2D)
        return sp_fail(s, SP_ERR_INVALID, "unknown program id");
    const bool slab = s->slab != nullptr;
    if (slab && program != SP_PROGRAM_WCSPH_3D)
        return sp_fail(s, SP_ERR_STATE, "only the 3-D WCSPH program runs on slab systems");
    const int UNIT = 2;
    // steps [0, nsteps): step 0 opens the run, step nsteps-1 closes it; the steps in between are all alike and are
    // what a graph unit holds.  Two eager steps come first (lazy allocations, list capacity), so a graph pays off
    // from about 8 steps on.
    const bool use_graph = graphs_enabled() && !slab && nsteps >= 8 && !s->capturing;
    int64_t k = 0;
    auto eager = [&](int64_t upto) -> int {
        for (; k < upto; k++)
            if ((rc = run_step(s, program, F, P, k == 0, k == nsteps - 1))) return rc;
        return SP_OK;
    };
    if (!use_graph) return eager(nsteps);
    if ((rc = eager(2))) return rc;
    // middle steps [2, nsteps-1)
    int tries = 0;
    bool shifted = false;
    while (k + UNIT <= nsteps - 1) {
        if (!graph_matches(s, program, F, P)) {
            if (s->graph.exec && s->graph.sig && !shifted && k + 1 + UNIT <= nsteps - 1) {
                // a graph from an earlier call may be one step out of phase with the ping-pong planes: one eager step
                // is cheaper than a capture
                shifted = true;
                if ((rc = eager(k + 1))) return rc;
                continue;
            }
            if (tries >= 2) break;  // not periodic in this state: run the rest eagerly
            tries++;
            bool ok = false;
            if ((rc = capture_unit(s, program, F, P, UNIT, &ok))) return rc;
            if (!ok) break;
            // the capture advanced the host bookkeeping (versions, ping-pong swaps) by one unit without executing
            // anything: launch the captured unit once for that.  If the bookkeeping did not close on itself (a field
            // changed its known-zero state, ...) the next round captures again from the new state.
            SP_CUDA(s, cudaGraphLaunch(s->graph.exec, s->stream));
            s->n_exact = false;
            s->count_pending = false;  // a count fetched before this launch says nothing about the state after it
            k += UNIT;
            continue;
        }
        SP_CUDA(s, cudaGraphLaunch(s->graph.exec, s->stream));
        s->launches += s->graph.launches;
        s->n_exact = false;
        s->count_pending = false;
        k += UNIT;
    }
    return eager(nsteps);  // at least the closing step: its build also fetches the counts and the longest-list report
}
