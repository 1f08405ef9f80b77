# examples/cavity_flow.jl (2-D lid-driven cavity, BASELINE configs[3]) on the B200 engine.  NOT EXECUTED in the build
# environment (no Julia runtime); configs.cavity_flow() issues the same calls through ctypes and is parity-tested.
module cavity_flow_b200

import SmoothedParticles as SP
include(joinpath(@__DIR__, "..", "SmoothedParticlesB200.jl"))
using .SmoothedParticlesB200
const Ops = SmoothedParticlesB200.Operators

const N = 100                 # constants of the original, :28-47
const Re = 100
const llid = 1.0
const rho0 = 1.0
const vlid = 1.0
const dr = llid / N
const h = 3.0 * dr
const m = rho0 * dr^2
const c = 20 * vlid
const P0 = 5.0
const wwall = h
const dt = 0.1 * h / c
const t_end = 0.4
const FLUID, WALL, LID = 0.0, 1.0, 2.0

const find_pressure = Ops.find_pressure(dt, c, rho0; P0 = P0)                 # :96-100
const balance_of_mass = Ops.balance_of_mass(:wendland2, m, h, 0.0)           # :92-94
const internal_force = Ops.internal_force_cavity(m, h, Re, vlid; ylid = llid, lid = LID)   # :102-114
const move = Ops.move(0.5 * dt)                                              # :117-122
const accelerate = Ops.accelerate(0.5 * dt)                                  # :124-128

function make_system()                                                        # :73-86
    grid = SP.Grid(dr, :hexagonal)
    box = SP.Rectangle(0.0, 0.0, llid, llid)
    wall = SP.BoundaryLayer(box, grid, wwall)
    sys = ParticleSystem([:v => 3, :Dv => 3, :rho => 1, :Drho => 1, :P => 1, :type => 1], SP.boundarybox(box + wall), h)
    lid = SP.Specification(wall, x -> x[2] > llid)
    wall = SP.Specification(wall, x -> x[2] <= llid)
    xf, xl, xw = SP.covering(grid, box), SP.covering(grid, lid), SP.covering(grid, wall)
    n = length(xf) + length(xl) + length(xw)
    add_particles!(sys; x = positions(vcat(xf, xl, xw)), rho = fill(rho0, n),
                   type = vcat(fill(FLUID, length(xf)), fill(LID, length(xl)), fill(WALL, length(xw))))
    create_cell_list!(sys)
    apply!(sys, find_pressure)
    apply!(sys, internal_force)
    return sys
end

# compute_fluxes, :162-180: SmoothedParticles.sum(sys, f, x) for all sample points at once (SP_SUM_MASS_W = 1,
# SP_SUM_MASS_F_W = 2 of include/sp_b200.h)
function compute_fluxes(sys::ParticleSystem, res = 100)
    s = collect(range(0.0, 1.0, length = res))
    ycl = Float64[i == 1 ? 0.5 : (i == 2 ? s[k] : 0.0) for i in 1:3, k in 1:res]   # points (0.5, s, 0)
    xcl = Float64[i == 1 ? s[k] : (i == 2 ? 0.5 : 0.0) for i in 1:3, k in 1:res]   # points (s, 0.5, 0)
    gamma(pts) = sum_at_points(sys, 1, [:x, :type], Float64[2.0, m, h, FLUID], pts)
    flux(pts, comp) = sum_at_points(sys, 2, [:x, :type, :v], Float64[2.0, m, h, FLUID, comp], pts)
    v1 = flux(ycl, 0.0) ./ gamma(ycl)
    v2 = flux(xcl, 1.0) ./ gamma(xcl)
    return s, v1, v2
end

function main(; nsteps = Int64(round(t_end / dt)))
    sys = make_system()
    for k in 0:nsteps                                                         # :137-151
        apply!(sys, accelerate)
        apply!(sys, move)
        create_cell_list!(sys)
        apply!(sys, balance_of_mass)
        apply!(sys, find_pressure)
        apply!(sys, move)
        create_cell_list!(sys)
        apply!(sys, internal_force)
        apply!(sys, accelerate)
    end
    create_cell_list!(sys)
    return compute_fluxes(sys)
end

end # module
