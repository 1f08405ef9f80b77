# SmoothedParticlesB200.jl — thin Julia shim over libsp_b200.so (include/sp_b200.h).
#
# Keeps the reference's surface — ParticleSystem / create_cell_list! / apply! / ParticleField /
# assemble_vector / sum (src/SmoothedParticles.jl:10-72) — and forwards every call with `ccall`.
# All logic lives in the shared library; this file only marshals arguments.  It could not be executed in the
# build environment (no Julia runtime there); the same entry points are exercised through Python ctypes.
#
# Differences a user sees (unavoidable once particles live in HBM):
#   * the particle struct is declared as a list of fields  (:v => 3, :rho => 1, ...)  instead of a mutable struct;
#   * `sys.particles[i].x` becomes bulk `download(sys, :x)` / `upload!(sys, :x, array)` in reference order;
#   * the closures passed to apply! are registered operators (Operators.balance_of_mass(...), ...).
module SmoothedParticlesB200

export ParticleSystem, record!, replay!, destroy_graph!, create_cell_list!, build_neighbour_lists!, apply!, respawn!, upload!, download, add_particles!, ParticleField,
       assemble_vector, assemble_matrix, Operators, poisson_cg!, reduce_energy_wcsph, sum_at_points, run_program!, front, cfl_time_step, positions, sp_reduce,
       kernel_eval, wendland1, Dwendland1, rDwendland1, wendland2, Dwendland2, rDwendland2, wendland3, Dwendland3, rDwendland3,
       DDwendland3, spline23, Dspline23, rDspline23, spline24, Dspline24, rDspline24, synchronize, num_removed, key_params,
       poisson_apply!, cell_keys, cell_list, neighbour_lists, slab_unique_id, slab_init!, slab_range, slab_create_cell_list!,
       slab_halo_refresh!, slab_num_owned, slab_allreduce!, generate_particles!, HalfSpace

const LIB = get(ENV, "SP_B200_LIB", joinpath(@__DIR__, "..", "smoothedparticles.jl_b200", "libsp_b200.so"))

const SP_LAYOUT_AOS = Int32(0)
const SP_FLAG_SELF = Int32(1)
const SP_ERR_NOT_CONVERGED = Int32(6)
const KERNELS = Dict(:wendland1 => 1.0, :wendland2 => 2.0, :wendland3 => 3.0, :spline23 => 4.0, :spline24 => 5.0)

struct SpError <: Exception
    code::Int32
    msg::String
end

function check(code::Int32, handle::Ptr{Cvoid} = C_NULL)
    code == 0 && return
    msg = unsafe_string(ccall((:sp_last_error, LIB), Cstring, (Ptr{Cvoid},), handle))
    throw(SpError(code, msg))     # the reference throws / @asserts; the ABI never throws across ccall
end

mutable struct ParticleSystem
    handle::Ptr{Cvoid}
    h::Float64
    fields::Dict{Symbol,Tuple{Int32,Int32}}   # name => (field id, ncomp)
    # ParticleSystem(T, domain, h), src/structs.jl:57-91: `fields` replaces T, `lo`/`hi` = boundarybox(domain)
    function ParticleSystem(fields::Vector{Pair{Symbol,Int}}, lo::NTuple{3,Float64}, hi::NTuple{3,Float64},
                            h::Float64; device::Integer = 0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:sp_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Ref{NTuple{3,Float64}}, Ref{NTuple{3,Float64}}, Float64, Int32),
                    out, Ref(lo), Ref(hi), h, Int32(device)))
        sys = new(out[], h, Dict{Symbol,Tuple{Int32,Int32}}(:x => (Int32(0), Int32(3))))
        for (name, nc) in fields
            fid = Ref{Int32}(0)
            check(ccall((:sp_add_field, LIB), Int32, (Ptr{Cvoid}, Cstring, Int32, Ref{Int32}), sys.handle, String(name),
                        Int32(nc), fid), sys.handle)
            sys.fields[name] = (fid[], Int32(nc))
        end
        finalizer(s -> ccall((:sp_destroy, LIB), Int32, (Ptr{Cvoid},), s.handle), sys)
        return sys
    end
end

# ParticleSystem(fields, domain, h) with `domain` = boundarybox(shape) of the reference's geometry module (any object
# with the six Box fields, src/geometry.jl:15-22)
ParticleSystem(fields::Vector{Pair{Symbol,Int}}, box, h::Float64; device::Integer = 0) =
    ParticleSystem(fields, (box.x1_min, box.x2_min, box.x3_min), (box.x1_max, box.x2_max, box.x3_max), h; device = device)

# 3 x n Float64 matrix (the AoS layout of upload!) from a Vector of RealVector / SVector{3,Float64}
positions(xs::AbstractVector) = Float64[x[i] for i in 1:3, x in xs]

Base.length(sys::ParticleSystem) = begin
    n = Ref{Int64}(0)
    check(ccall((:sp_num_particles, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), sys.handle, n), sys.handle)
    Int(n[])
end

# host arrays are ncomp x n column-major (a reinterpreted Vector{SVector{3,Float64}}) == SP_LAYOUT_AOS
function upload!(sys::ParticleSystem, name::Symbol, a::AbstractArray{Float64})
    fid, _ = sys.fields[name]
    check(ccall((:sp_upload, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64, Int32), sys.handle, fid, a,
                length(sys), SP_LAYOUT_AOS), sys.handle)
end
function download(sys::ParticleSystem, name::Symbol)
    fid, nc = sys.fields[name]
    a = nc == 1 ? Vector{Float64}(undef, length(sys)) : Matrix{Float64}(undef, nc, length(sys))
    check(ccall((:sp_download, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64, Int32), sys.handle, fid, a,
                length(sys), SP_LAYOUT_AOS), sys.handle)
    return a
end
# generate_particles! / push!, src/grids.jl:253-258: append at the end of the reference order
function add_particles!(sys::ParticleSystem; kwargs...)
    x = kwargs[:x]
    n_old, n_new = length(sys), size(x, 2)
    old = Dict(k => download(sys, k) for k in keys(kwargs) if n_old > 0)
    check(ccall((:sp_resize, LIB), Int32, (Ptr{Cvoid}, Int64), sys.handle, n_old + n_new), sys.handle)
    for (k, v) in kwargs
        upload!(sys, k, n_old > 0 ? (ndims(v) == 1 ? vcat(old[k], v) : hcat(old[k], v)) : v)
    end
end

# create_cell_list!(sys), src/core.jl:51-90
create_cell_list!(sys::ParticleSystem) =
    check(ccall((:sp_create_cell_list, LIB), Int32, (Ptr{Cvoid},), sys.handle), sys.handle)

# Optional: evaluate the op-independent half of apply_binary! (src/core.jl:94-110) now; every binary apply! replays
# the cached neighbour lists until positions change (the first apply! after a move does this lazily).
build_neighbour_lists!(sys::ParticleSystem) =
    check(ccall((:sp_build_neighbour_lists, LIB), Int32, (Ptr{Cvoid},), sys.handle), sys.handle)

struct Operator
    id::Int32
    fields::Vector{Symbol}
    params::Vector{Float64}
end

# apply!(sys, action!; self = false), src/core.jl:151-161
function apply!(sys::ParticleSystem, op::Operator; self::Bool = false)
    F = Int32[sys.fields[f][1] for f in op.fields]
    check(ccall((:sp_apply, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Float64}, Int32, Int32), sys.handle,
                op.id, F, length(F), op.params, length(op.params), self ? SP_FLAG_SELF : Int32(0)), sys.handle)
end

# the closures of the shipped examples as registered operators (ids and parameter order: include/sp_b200.h)
module Operators
import ..Operator, ..KERNELS
balance_of_mass(kernel, m, h, nu; x = :x, v = :v, rho = :rho, Drho = :Drho) =
    Operator(1, [x, v, rho, Drho], [KERNELS[kernel], m, h, 2 * nu])                  # collapse_dry.jl:112-115
find_pressure(dt, c, rho0; P0 = 0.0, rho = :rho, Drho = :Drho, P = :P) =
    Operator(2, [rho, Drho, P], [dt, c^2, rho0, P0])                                 # collapse_dry.jl:123-127
internal_force(kernel, m, h, mu, rho0; x = :x, v = :v, P = :P, rho = :rho, Dv = :Dv, type = :type) =
    Operator(3, [x, v, P, rho, Dv, type], [KERNELS[kernel], m, h, mu, rho0])         # collapse_dry.jl:135-141
internal_force_cavity(m, h, Re, vlid; ylid = 1.0, lid = 2.0, x = :x, v = :v, P = :P, rho = :rho, Dv = :Dv, type = :type) =
    Operator(4, [x, v, P, rho, Dv, type], [m, h, Float64(Re), vlid, ylid, lid])      # cavity_flow.jl:102-114
move(dtm; x = :x, v = :v, Dv = :Dv, type = :type) = Operator(5, [x, v, Dv, type], [dtm])                 # :148-153
accelerate(hdt, g = (0.0, 0.0, 0.0); v = :v, Dv = :Dv, type = :type) = Operator(6, [v, Dv, type], [hdt, g...])  # :155-159
density_sum(kernel, m, h; x = :x, out = :rho) = Operator(7, [x, out], [KERNELS[kernel], m, h])           # test_collision_2d.jl:63-69
pressure_from_rho(c; rho = :rho, rho0 = :rho0, P = :P) = Operator(8, [rho, rho0, P], [c^2])
internal_force_sym(kernel, m, h, rho0; x = :x, P = :P, a = :a) = Operator(9, [x, P, a], [KERNELS[kernel], m, h, rho0])
fill(field, value = 0.0) = Operator(10, [field], [value])
advect(dt; x = :x, v = :v) = Operator(11, [x, v], [dt])
kick(hdt; v = :v, a = :a) = Operator(12, [v, a], [hdt])
isph_initialize(dt, g; x = :x, v = :v, div = :div, L = :L, lambda = :lambda, type = :type) =
    Operator(20, [x, v, div, L, lambda, type], [dt, g...])                           # collapse_dry_implicit.jl:118-126
isph_viscous_force(kernel, m, h, mu, rho; x = :x, v = :v, Dv = :Dv) = Operator(21, [x, v, Dv], [KERNELS[kernel], m, h, mu, rho])
isph_div_L_lambda(kernel, m, h, rho, dim; x = :x, v = :v, div = :div, L = :L, lambda = :lambda) =
    Operator(22, [x, v, div, L, lambda], [KERNELS[kernel], m, h, rho, Float64(dim)])
isph_projection_vector(h, dt; div = :div, b = :b) = Operator(23, [div, b], [h, dt])
isph_internal_force(kernel, m, h, rho; x = :x, P = :P, Dv = :Dv) = Operator(25, [x, P, Dv], [KERNELS[kernel], m, h, rho])
isph_accelerate(dt; v = :v, Dv = :Dv, type = :type) = Operator(26, [v, Dv, type], [dt])
# examples/static_container.jl:102-119
sc_balance_of_mass(kernel, m, h, dt; x = :x, v = :v, rho = :rho) = Operator(30, [x, v, rho], [KERNELS[kernel], m, h, dt])
sc_internal_force(kernel, m, h, mu, c, rho0; x = :x, v = :v, rho = :rho, a = :a, type = :type) =
    Operator(31, [x, v, rho, a, type], [KERNELS[kernel], m, h, mu, c^2, rho0])
move_all(dtm; x = :x, v = :v, a = :a) = Operator(32, [x, v, a], [dtm])
# examples/drop.jl:76-113
find_normal(kernel, vol, h; x = :x, n = :n) = Operator(33, [x, n], [KERNELS[kernel], 2 * vol * vol, h])
normalize(s0; n = :n) = Operator(34, [n], [s0])
internal_force_tension(m, h, mu, rho0, beta, s0; x = :x, v = :v, P = :P, n = :n, a = :a) =
    Operator(35, [x, v, P, n, a], [m, h, mu, rho0, beta, s0])
# examples/collapse_symplectic.jl:98-153, examples/Kepler_vortex.jl:139-193 (rev_add: examples/utils/FixPA.jl)
density_sum_fluid(kernel, m, h; x = :x, out = :rho, type = :type) = Operator(40, [x, out, type], [KERNELS[kernel], m, h])
internal_force_lj(kernel, m, h, dr_wall, E_wall, eps; rho0 = 0.0, wall_type = 1.0, x = :x, P = :P, rho = :rho, a = :a,
                  type = :type) =
    Operator(41, [x, P, rho, a, type], [KERNELS[kernel], m, h, rho0, wall_type, dr_wall, E_wall, eps])
move_rev(dt; x = :x, v = :v, type = :type) = Operator(42, [x, v, type], [dt])
accelerate_rev(hdt, g = (0.0, 0.0, 0.0); v = :v, a = :a, type = :type) = Operator(43, [v, a, type], [hdt, g...])
accelerate_rev_central(hdt, GM; x = :x, v = :v, a = :a, type = :type) = Operator(44, [x, v, a, type], [hdt, GM])
lj_potential(h, m, E_wall, dr_wall, eps; wall_type = 1.0, x = :x, out = :U, type = :type) =
    Operator(45, [x, out, type], [h, m * E_wall, wall_type, dr_wall, eps])
# examples/cylinder.jl:91-143
cyl_balance_of_mass(kernel, h, nu; x = :x, v = :v, rho = :rho, Drho = :Drho, m = :m, type = :type) =
    Operator(50, [x, v, rho, Drho, m, type], [KERNELS[kernel], h, 2 * nu])
cyl_find_pressure(dt, c, rho0, x1_min; x = :x, rho = :rho, Drho = :Drho, P = :P) =
    Operator(51, [x, rho, Drho, P], [dt, c^2, rho0, x1_min])
cyl_internal_force(kernel, h, mu; x = :x, v = :v, P = :P, rho = :rho, a = :a, m = :m) =
    Operator(52, [x, v, P, rho, a, m], [KERNELS[kernel], h, mu, 0.01 * h * h])
move_types(dt, type_a, type_b; x = :x, v = :v, a = :a, type = :type) = Operator(53, [x, v, a, type], [dt, type_a, type_b])
cyl_accelerate(hdt, cyl1, U_max; x = :x, v = :v, a = :a, type = :type) = Operator(54, [x, v, a, type], [hdt, cyl1, 0.3 * U_max^2])
set_inflow_speed(t, t_acc, U_max, chan_w; inflow_type = 1.0, x = :x, v = :v, type = :type) =
    Operator(55, [x, v, type], [inflow_type, min(1.0, t / t_acc), U_max, chan_w])
# examples/rod.jl:128-188 (A, H, B are 9-component RealMatrix fields)
rod_find_A(kernel, h; x = :x, X = :X, A = :A, H = :H) = Operator(60, [x, X, A, H], [KERNELS[kernel], h])
rod_find_B(m, c_l, c_s; A = :A, H = :H, B = :B) = Operator(61, [A, H, B], [m, c_l, c_s])
rod_find_f(kernel, h, m, vol, nu; x = :x, v = :v, X = :X, A = :A, B = :B, f = :f) =
    Operator(62, [x, v, X, A, B, f], [KERNELS[kernel], h, 2 * m * vol, nu])
rod_pull(X1_min, fy; X = :X, f = :f) = Operator(63, [X, f], [X1_min, fy])
rod_update_v(hdt, m, X1_clamp; v = :v, f = :f, X = :X) = Operator(64, [v, f, X], [hdt, m, X1_clamp])
rod_update_x(dt; x = :x, v = :v, A = :A, H = :H, f = :f, e = :e) = Operator(65, [x, v, A, H, f, e], [dt])
rod_find_e(h; x = :x, X = :X, A = :A, e = :e) = Operator(66, [x, X, A, e], [h])
# examples/SHTC/ldc.jl:90-133 (A and stress are 9-component RealMatrix fields; parity on B200: tests/test_shtc_gpu.py)
shtc_find_stress(c_l, c_s, rho0, acf; A = :A, rho = :rho, stress = :stress) = Operator(70, [A, rho, stress], [c_l, c_s, rho0 / (1.0 + acf)])
shtc_update_v(kernel, h, dt, m; x = :x, v = :v, rho = :rho, stress = :stress, type = :type) =
    Operator(71, [x, v, rho, stress, type], [KERNELS[kernel], h, dt * m])
shtc_update_rho(kernel, h, dt, m; x = :x, v = :v, rho = :rho, type = :type) = Operator(72, [x, v, rho, type], [KERNELS[kernel], h, dt * m])
shtc_convect_A(kernel, h, dt, m, skip_type; x = :x, v = :v, rho = :rho, A = :A, type = :type) =
    Operator(73, [x, v, rho, A, type], [KERNELS[kernel], h, dt * m, skip_type])
shtc_relax_A(dt, tau; A = :A) = Operator(74, [A], [dt, tau])
shtc_move(dt; x = :x, v = :v, type = :type) = Operator(75, [x, v, type], [dt])
# examples/SHTC/beryllium.jl:132-184 (update_x! is advect; parity on B200: tests/test_shtc_gpu.py)
be_find_L(kernel, h, rho0; x = :x, v = :v, m = :m, T = :T, L = :L) = Operator(80, [x, v, m, T, L], [KERNELS[kernel], h, rho0])
be_update_A(hdt; A = :A, T = :T, L = :L) = Operator(81, [A, T, L], [hdt])
be_find_J(kernel, h, rho0; x = :x, m = :m, T = :T, J = :J, K = :K) = Operator(82, [x, m, T, J, K], [KERNELS[kernel], h, rho0])
be_find_T(rho0, c_0, c_s; A = :A, T = :T, P = :P, J = :J) = Operator(83, [A, T, P, J], [rho0, c_0, c_s])
be_find_f(kernel, h, rho0, c_p; x = :x, m = :m, T = :T, K = :K, f = :f) = Operator(84, [x, m, T, K, f], [KERNELS[kernel], h, rho0, c_p])
be_reset(; f = :f, L = :L, T = :T, J = :J, K = :K, J0 = :J0, K0 = :K0) = Operator(85, [f, L, T, J, K, J0, K0], Float64[])
be_update_v(hdt; v = :v, f = :f, m = :m) = Operator(86, [v, f, m], [hdt])
# examples/SHTC/twist3d.jl:125-172 (reset! is be_reset, update_x! is advect; parity on B200: tests/test_shtc_gpu.py)
tw_find_L(kernel, h, rho0; x = :x, v = :v, m = :m, T = :T, L = :L) = Operator(90, [x, v, m, T, L], [KERNELS[kernel], h, rho0])
tw_update_A(hdt; A = :A, T = :T, L = :L) = Operator(91, [A, T, L], [hdt])
tw_find_J(kernel, h, rho0; x = :x, m = :m, T = :T, J = :J, K = :K) = Operator(92, [x, m, T, J, K], [KERNELS[kernel], h, rho0])
tw_find_T(rho0, c_0, c_s; A = :A, T = :T, P = :P, J = :J) = Operator(93, [A, T, P, J], [rho0, c_0, c_s])
tw_find_f(kernel, h, rho0, c_p; x = :x, m = :m, T = :T, K = :K, f = :f) = Operator(94, [x, m, T, K, f], [KERNELS[kernel], h, rho0, c_p])
tw_update_v(hdt; x = :x, v = :v, f = :f, m = :m) = Operator(95, [x, v, f, m], [hdt])
# examples/SHTC/taco.jl:108-162: find_L!, update_A!, reset!, find_rho! (apply! with self = true) are be_find_L / be_update_A /
# be_reset / be_find_J with rho0 = 1.0 and the taco field names; relax_A! is shtc_relax_A (parity on B200: tests/test_shtc_gpu.py)
ta_find_T(rho0, c_0, c_s; A = :A, T = :T, P = :P, rho = :rho) = Operator(100, [A, T, P, rho], [rho0, c_0, c_s])
ta_find_f(kernel, h, c_p, rho0; x = :x, m = :m, T = :T, lambda = :lambda, f = :f) =
    Operator(101, [x, m, T, lambda, f], [KERNELS[kernel], h, (c_p / rho0)^2])
ta_update_v(hdt, R1, R2, omega; x = :x, v = :v, f = :f, m = :m, type = :type) = Operator(102, [x, v, f, m, type], [hdt, R1, R2, omega])
ta_update_x(hdt, omega, t, outer_type; x = :x, v = :v, x0 = :x0, type = :type) =
    Operator(103, [x, v, x0, type], [hdt, cos(omega * t), sin(omega * t), outer_type])
end # module Operators

# add_new_particles!, examples/cylinder.jl:145-156: particles of `from_type` with x[1] >= x1_min become `to_type`, a new
# `from_type` particle appears `shift` upstream of each (reference order kept); `constants` = what the constructor sets
function respawn!(sys::ParticleSystem, type_field::Symbol, from_type, to_type, x1_min, shift; constants...)
    ff = Int32[sys.fields[k][1] for (k, _) in constants]
    fv = Float64[v for (_, v) in constants]
    added = Ref{Int64}(0)
    check(ccall((:sp_respawn, LIB), Int32,
                (Ptr{Cvoid}, Int32, Float64, Float64, Float64, Float64, Ptr{Int32}, Ptr{Float64}, Int32, Ref{Int64}),
                sys.handle, sys.fields[type_field][1], Float64(from_type), Float64(to_type), Float64(x1_min), Float64(shift),
                ff, fv, length(ff), added), sys.handle)
    return added[]
end

# assemble_vector(sys, func), src/core.jl:175-182: the unary operator writes its last bound field
function assemble_vector(sys::ParticleSystem, op::Operator)
    apply!(sys, op)
    return download(sys, op.fields[end])
end

# P .= cg(assemble_matrix(sys, projection_matrix), b), collapse_dry_implicit.jl:223-227, matrix-free
function poisson_cg!(sys::ParticleSystem, kernel, m, h, rho, C_free; x = :x, L = :L, lambda = :lambda, type = :type,
                     b = :b, P = :P, reltol = sqrt(eps()), abstol = 0.0, maxiter = 0)
    F = Int32[sys.fields[f][1] for f in (x, L, lambda, type, b, P)]
    prm = Float64[KERNELS[kernel], m, h, rho, C_free]
    iters = Ref{Int64}(0); resid = Ref{Float64}(0.0)
    rc = ccall((:sp_poisson_cg, LIB), Int32,
               (Ptr{Cvoid}, Ptr{Int32}, Int32, Ptr{Float64}, Int32, Float64, Float64, Int64, Ref{Int64}, Ref{Float64}),
               sys.handle, F, 6, prm, 5, reltol, abstol, maxiter, iters, resid)
    # IterativeSolvers.cg returns its last iterate at maxiter without an error: SP_ERR_NOT_CONVERGED is a soft status,
    # P holds that iterate; the third return value says whether the tolerance was met
    rc == SP_ERR_NOT_CONVERGED || check(rc, sys.handle)
    return iters[], resid[], rc != SP_ERR_NOT_CONVERGED
end

# generic diagnostics reduction (SP_RED_* of include/sp_b200.h): returns the first `nout` results
function sp_reduce(sys::ParticleSystem, red::Integer, fields::Vector{Symbol}, params::Vector{Float64} = Float64[]; nout::Integer = 1)
    F = Int32[sys.fields[f][1] for f in fields]
    out = zeros(3)
    prm = isempty(params) ? zeros(1) : params
    check(ccall((:sp_reduce, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Float64}, Int32, Ptr{Float64}),
                sys.handle, Int32(red), F, length(F), prm, length(params), out), sys.handle)
    return out[1:nout]
end
# get_globals, collapse_dry.jl:173-187: (X, H) of the dam-break front
front(sys::ParticleSystem, width, height, h, xmax; x = :x, type = :type) =
    sp_reduce(sys, 2, [x, type], Float64[width, height, h, xmax]; nout = 2)
# adaptive time step dt = cfl*h/(c + max|v|) (no counterpart in the reference; all-reduced on slab systems)
cfl_time_step(sys::ParticleSystem, cfl, h, c; v = :v) = cfl * h / (c + sp_reduce(sys, 8, [v])[1])

# the whole time loop of examples/collapse3d.jl:136-150 (program = 1) or collapse_dry.jl:203-211 (program = 2) issued
# from inside the library: one ccall for `nsteps` steps, same arithmetic as the call-by-call loop
function run_program!(sys::ParticleSystem, program::Integer, kernel, m, h, nu, dt, c, rho0, mu, g, nsteps::Integer;
                      x = :x, v = :v, Dv = :Dv, rho = :rho, Drho = :Drho, P = :P, type = :type)
    F = Int32[sys.fields[f][1] for f in (x, v, Dv, rho, Drho, P, type)]
    prm = Float64[KERNELS[kernel], m, h, 2 * nu, dt, c^2, rho0, mu, g...]
    check(ccall((:sp_run_program, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Float64}, Int32, Int64),
                sys.handle, Int32(program), F, length(F), prm, length(prm), Int64(nsteps)), sys.handle)
end

# A loop body recorded once and replayed as one CUDA graph launch (sp_graph_* of include/sp_b200.h):
#     g = record!(sys, 2) do; step!(sys); end      # runs the body twice as one unit, returns the graph
#     replay!(sys, g, 500)                          # 1000 more time steps, 500 graph launches
# The body may only contain apply! / create_cell_list! style calls (nothing that hands data to the host), and the unit
# must contain an even number of cell-list builds.
function record!(body::Function, sys::ParticleSystem, repeat::Integer = 2)
    check(ccall((:sp_graph_begin, LIB), Int32, (Ptr{Cvoid},), sys.handle), sys.handle)
    gid = Ref{Int32}(-1)
    try
        for _ in 1:repeat
            body()
        end
    catch
        ccall((:sp_graph_end, LIB), Int32, (Ptr{Cvoid}, Ref{Int32}), sys.handle, gid)
        rethrow()
    end
    check(ccall((:sp_graph_end, LIB), Int32, (Ptr{Cvoid}, Ref{Int32}), sys.handle, gid), sys.handle)
    return gid[]
end
replay!(sys::ParticleSystem, graph::Integer, times::Integer = 1) =
    check(ccall((:sp_graph_launch, LIB), Int32, (Ptr{Cvoid}, Int32, Int64), sys.handle, Int32(graph), Int64(times)), sys.handle)
destroy_graph!(sys::ParticleSystem, graph::Integer) =
    check(ccall((:sp_graph_destroy, LIB), Int32, (Ptr{Cvoid}, Int32), sys.handle, Int32(graph)), sys.handle)

# A = assemble_matrix(sys, projection_matrix), src/core.jl:196-225, for hosts that want the matrix itself:
# returns (I, J, V) — feed them to SparseArrays.sparse(I, J, V, N, N) exactly as the reference does
function assemble_matrix(sys::ParticleSystem, kernel, m, h, rho, C_free; x = :x, L = :L, lambda = :lambda, type = :type)
    F = Int32[sys.fields[f][1] for f in (x, L, lambda, type)]
    prm = Float64[KERNELS[kernel], m, h, rho, C_free]
    nnz = Ref{Int64}(0)
    check(ccall((:sp_assemble_matrix, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Int32}, Int32, Ptr{Float64}, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Ref{Int64}),
                sys.handle, F, 4, prm, 5, C_NULL, C_NULL, C_NULL, 0, nnz), sys.handle)
    I = Vector{Int64}(undef, nnz[]); J = Vector{Int64}(undef, nnz[]); V = Vector{Float64}(undef, nnz[])
    nnz[] > 0 && check(ccall((:sp_assemble_matrix, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Int32}, Int32, Ptr{Float64}, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Ref{Int64}),
                sys.handle, F, 4, prm, 5, I, J, V, nnz[], nnz), sys.handle)
    return I, J, V
end

# sum(energy, sys.particles), collapse_dry.jl:166-171
function reduce_energy_wcsph(sys::ParticleSystem, m, c, rho0, g; x = :x, v = :v, rho = :rho)
    F = Int32[sys.fields[f][1] for f in (x, v, rho)]
    out = zeros(3)
    check(ccall((:sp_reduce, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Float64}, Int32, Ptr{Float64}),
                sys.handle, 1, F, 3, Float64[m, c, rho0, g...], 6, out), sys.handle)
    return out[1]
end

# SmoothedParticles.sum(sys, f, x), src/core.jl:240-260, for many points at once (3 x m matrix)
function sum_at_points(sys::ParticleSystem, sum_op::Integer, fields::Vector{Symbol}, params::Vector{Float64},
                       pts::Matrix{Float64})
    F = Int32[sys.fields[f][1] for f in fields]
    out = Vector{Float64}(undef, size(pts, 2))
    check(ccall((:sp_sum_at_points, LIB), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Int64, Ptr{Float64}),
                sys.handle, Int32(sum_op), F, length(F), params, length(params), pts, size(pts, 2), out), sys.handle)
    return out
end

# ---- generate_particles!(sys, grid, shape, constructor) on the device (src/grids.jl:253-258, sp_generate_particles)
# `grid` and `shape` are the reference package's own objects (SmoothedParticles.Grid(...), Rectangle, Circle, Ball,
# Boolean +,-,*, BoundaryLayer, Specification): the shape tree is walked by type NAME (no dependency on the package) into
# the postfix programme of include/sp_b200.h.  A Specification's predicate must be a HalfSpace (below) — a Julia
# closure cannot run inside a CUDA kernel; use the host path (SP.covering + add_particles!) for anything else.
struct HalfSpace          # x[axis] op bound, op in (:<, :<=, :>, :>=): the form of every Specification lambda in the examples
    axis::Int
    op::Symbol
    bound::Float64
end
(hs::HalfSpace)(x) = getfield(Base, hs.op)(x[hs.axis], hs.bound)   # so that the reference's own is_inside accepts it too

struct ShapeNode          # sp_shape_node
    kind::Int32
    a::Int32
    b::Int32
    p::NTuple{8,Float64}
end
pad8(v...) = ntuple(i -> i <= length(v) ? Float64(v[i]) : 0.0, 8)

function compile_shape!(nodes::Vector{ShapeNode}, offsets::Vector{Float64}, s)
    T = nameof(typeof(s))
    emit(kind, a, b, p) = (push!(nodes, ShapeNode(Int32(kind), Int32(a), Int32(b), p)); length(nodes) - 1)
    if T == :Box
        return emit(1, 0, 0, pad8(s.x1_min, s.x2_min, s.x3_min, s.x1_max, s.x2_max, s.x3_max))
    elseif T == :Circle
        return emit(2, 0, 0, pad8(s.x1, s.x2, s.r * s.r))
    elseif T == :Ball
        return emit(3, 0, 0, pad8(s.x1, s.x2, s.x3, s.r * s.r))
    elseif T in (:BooleanUnion, :BooleanIntersection, :BooleanDifference)
        a = compile_shape!(nodes, offsets, s.s1)
        b = compile_shape!(nodes, offsets, s.s2)
        return emit(T == :BooleanUnion ? 4 : T == :BooleanIntersection ? 5 : 6, a, b, pad8())
    elseif T == :Specification
        s.f isa HalfSpace || error("only HalfSpace predicates can run on the device (a Julia closure cannot)")
        a = compile_shape!(nodes, offsets, s.s)
        b = emit(7, s.f.axis - 1, findfirst(==(s.f.op), (:<, :<=, :>, :>=)) - 1, pad8(s.f.bound))
        return emit(5, b, a, pad8())               # f(x) && is_inside(x, s), geometry.jl:183-185
    elseif T == :BoundaryLayer
        isempty(offsets) || error("one BoundaryLayer per shape on the device")
        a = compile_shape!(nodes, offsets, s.s)
        for dx in s.dxs, c in 1:3
            push!(offsets, dx[c])
        end
        return emit(8, a, 0, pad8())
    end
    error("shape $T is not supported by the device generator")
end

# `bbox` = SmoothedParticles.boundarybox(shape); `constants` = the fields the constructor sets to a constant
function generate_particles!(sys::ParticleSystem, grid, shape, bbox; constants...)
    G = nameof(typeof(grid))
    kind, a, b = G == :Squaregrid ? (1, grid.dr, grid.dr) : G == :Hexagrid ? (2, grid.a, grid.b) :
                 G == :CubicGrid ? (3, grid.dr, grid.dr) : error("grid $G is not supported by the device generator")
    fl(v, d) = Int64(floor(v / d)); cl(v, d) = Int64(ceil(v / d))
    irange = Int64[fl(bbox.x1_min, a) - (kind == 2 ? 1 : 0), cl(bbox.x1_max, a), fl(bbox.x2_min, b), cl(bbox.x2_max, b),
                   kind == 3 ? fl(bbox.x3_min, grid.dr) : 0, kind == 3 ? cl(bbox.x3_max, grid.dr) : 0]   # grids.jl:53-56,76-79,129-134
    nodes = ShapeNode[]; offsets = Float64[]
    compile_shape!(nodes, offsets, shape)
    ff = Int32[sys.fields[k][1] for (k, _) in constants]
    fv = Float64[v for (_, v) in constants]
    added = Ref{Int64}(0)
    check(ccall((:sp_generate_particles, LIB), Int32,
                (Ptr{Cvoid}, Int32, Float64, Ptr{ShapeNode}, Int32, Ptr{Float64}, Int32, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}, Int32,
                 Ref{Int64}),
                sys.handle, Int32(kind), grid.dr, nodes, length(nodes), isempty(offsets) ? zeros(3) : offsets,
                length(offsets) ÷ 3, irange, isempty(ff) ? Int32[0] : ff, isempty(fv) ? [0.0] : fv, length(ff), added), sys.handle)
    return added[]
end

# ---- kernel functions under the reference's names (src/kernels.jl), evaluated by the library: wendland2(h, r), ...
function kernel_eval(kernel::Symbol, kfun::Integer, h::Float64, r::Vector{Float64}; device::Integer = 0)
    out = similar(r)
    check(ccall((:sp_kernel_eval, LIB), Int32, (Int32, Int32, Float64, Ptr{Float64}, Ptr{Float64}, Int64, Int32),
                Int32(KERNELS[kernel]), Int32(kfun), h, r, out, length(r), Int32(device)))
    return out
end
for (k, name) in ((:wendland1, "wendland1"), (:wendland2, "wendland2"), (:wendland3, "wendland3"), (:spline23, "spline23"),
                  (:spline24, "spline24"))
    @eval $(Symbol(name))(h::Float64, r::Float64) = kernel_eval($(QuoteNode(k)), 0, h, [r])[1]
    @eval $(Symbol("D" * name))(h::Float64, r::Float64) = kernel_eval($(QuoteNode(k)), 1, h, [r])[1]
    @eval $(Symbol("rD" * name))(h::Float64, r::Float64) = kernel_eval($(QuoteNode(k)), 2, h, [r])[1]
end
DDwendland3(h::Float64, r::Float64) = kernel_eval(:wendland3, 3, h, [r])[1]

# ---- small entry points
version() = ccall((:sp_version, LIB), Int32, ())
function device_count()
    n = Ref{Int32}(0)
    check(ccall((:sp_device_count, LIB), Int32, (Ref{Int32},), n))
    return Int(n[])
end
synchronize(sys::ParticleSystem) = check(ccall((:sp_synchronize, LIB), Int32, (Ptr{Cvoid},), sys.handle), sys.handle)
function num_removed(sys::ParticleSystem)       # particles dropped by create_cell_list! so far (src/core.jl:63-81)
    n = Ref{Int64}(0)
    check(ccall((:sp_num_removed, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), sys.handle, n), sys.handle)
    return Int(n[])
end
function key_params(sys::ParticleSystem)        # key_phase, key_lim, key_max, key_diff of src/structs.jl:63-82
    phase = zeros(Int64, 3); lim = zeros(Int64, 3); kmax = Ref{Int64}(0); nd = Ref{Int32}(0); diff = zeros(Int64, 27)
    check(ccall((:sp_key_params, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ref{Int64}, Ref{Int32}, Ptr{Int64}),
                sys.handle, phase, lim, kmax, nd, diff), sys.handle)
    return phase, lim, kmax[], diff[1:nd[]]
end
# (A p) of the pressure matrix without assembling it, collapse_dry_implicit.jl:154-163
function poisson_apply!(sys::ParticleSystem, kernel, m, h, rho, C_free, p_in::Symbol, y_out::Symbol;
                        x = :x, L = :L, lambda = :lambda, type = :type)
    F = Int32[sys.fields[f][1] for f in (x, L, lambda, type, p_in, y_out)]
    prm = Float64[KERNELS[kernel], m, h, rho, C_free]
    check(ccall((:sp_poisson_apply, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int32, Ptr{Float64}, Int32), sys.handle, F, 6, prm, 5),
          sys.handle)
end

# ---- parity / debug views, all in the reference's numbering (1-based)
function cell_keys(sys::ParticleSystem)         # find_key(sys, particles[i].x) as stored by the last create_cell_list!
    keys = Vector{Int64}(undef, length(sys))
    check(ccall((:sp_get_cell_keys, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Int64), sys.handle, keys, length(keys)), sys.handle)
    return keys
end
function cell_list(sys::ParticleSystem)         # CSR view of sys.cell_list: members of cell k in descending index
    _, _, kmax, _ = key_params(sys)
    offsets = Vector{Int64}(undef, kmax + 1); members = Vector{Int64}(undef, max(length(sys), 1))
    check(ccall((:sp_get_cell_list, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), sys.handle, offsets, members), sys.handle)
    return offsets, members[1:length(sys)]
end
function neighbour_lists(sys::ParticleSystem)   # the q of every action!(p, q, r) call of apply_binary!, in visiting order
    n = length(sys)
    offsets = Vector{Int64}(undef, n + 1)
    check(ccall((:sp_get_neighbour_lists, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Int64), sys.handle, offsets,
                C_NULL, 0), sys.handle)
    ids = Vector{Int64}(undef, max(offsets[end], 1))
    check(ccall((:sp_get_neighbour_lists, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Int64), sys.handle, offsets,
                ids, length(ids)), sys.handle)
    return offsets, ids[1:offsets[end]]
end

# ---- slab decomposition, one Julia process per GPU (INTEGRATION.md section 4)
function slab_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:sp_slab_unique_id, LIB), Int32, (Ptr{UInt8},), id))
    return id                                     # rank 0 sends these 128 bytes to the other ranks (MPI.jl / sockets)
end
# cuts (optional): rank r owns the cell layers cuts[r+1] .. cuts[r+2]-1 (0-based layer numbers; count-balanced slabs)
function slab_init!(sys::ParticleSystem, id::Vector{UInt8}, rank::Integer, nranks::Integer; periodic::Bool = false,
                    cuts::Union{Nothing, Vector{Int64}} = nothing)
    if cuts === nothing
        check(ccall((:sp_slab_init, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32, Int32), sys.handle, id, Int32(rank),
                    Int32(nranks), Int32(periodic)), sys.handle)
    else
        check(ccall((:sp_slab_init_cuts, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32, Int32, Ptr{Int64}), sys.handle, id,
                    Int32(rank), Int32(nranks), Int32(periodic), cuts), sys.handle)
    end
end
function slab_range(sys::ParticleSystem)          # the cell layers [cell_lo, cell_hi) and coordinates this rank owns
    clo = Ref{Int64}(0); chi = Ref{Int64}(0); xlo = Ref{Float64}(0.0); xhi = Ref{Float64}(0.0); axis = Ref{Int32}(0)
    check(ccall((:sp_slab_range, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Float64}, Ref{Float64}, Ref{Int32}),
                sys.handle, clo, chi, xlo, xhi, axis), sys.handle)
    return clo[], chi[], xlo[], xhi[], Int(axis[]) + 1
end
slab_create_cell_list!(sys::ParticleSystem) =     # migration + ghost layers + the local create_cell_list!
    check(ccall((:sp_slab_create_cell_list, LIB), Int32, (Ptr{Cvoid},), sys.handle), sys.handle)
function slab_halo_refresh!(sys::ParticleSystem, names::Symbol...)   # ghost copies of fields that change between rebuilds (a CG search vector); the WCSPH loops need none
    F = Int32[sys.fields[f][1] for f in names]
    check(ccall((:sp_slab_halo_refresh, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int32), sys.handle, F, length(F)), sys.handle)
end
function slab_num_owned(sys::ParticleSystem)
    n = Ref{Int64}(0)
    check(ccall((:sp_slab_num_owned, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), sys.handle, n), sys.handle)
    return Int(n[])
end
function slab_allreduce!(sys::ParticleSystem, values::Vector{Float64}; max::Bool = false)
    check(ccall((:sp_slab_allreduce, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32), sys.handle, values, length(values),
                Int32(max)), sys.handle)
    return values
end

# ParticleField(sys, :var), src/structs.jl:118-125
struct ParticleField <: AbstractVector{Float64}
    sys::ParticleSystem
    name::Symbol
end
Base.size(f::ParticleField) = (length(f.sys),)
Base.getindex(f::ParticleField, i::Int) = download(f.sys, f.name)[i]     # bulk use: collect(f)
Base.collect(f::ParticleField) = download(f.sys, f.name)
Base.copyto!(f::ParticleField, v::AbstractVector{Float64}) = (upload!(f.sys, f.name, collect(v)); f)

end # module
