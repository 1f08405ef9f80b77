# examples/collapse_dry.jl of SmoothedParticles.jl on the B200 engine: the host stays in Julia, geometry and lattice
# come from the reference package (set-up, host side), the hot path — create_cell_list! and every apply! — runs in
# libsp_b200.so through the shim.  The closures of the original script (:112-159) are the registered operators of the
# same names.  NOT EXECUTED in the build environment (no Julia runtime there); the same calls in the same order are
# what configs.collapse_dry() issues through ctypes, which is parity-tested against the CPU oracle.
module collapse_dry_b200

import SmoothedParticles as SP                 # geometry, grids, VTK output: unchanged host-side code
include(joinpath(@__DIR__, "..", "SmoothedParticlesB200.jl"))
using .SmoothedParticlesB200
const Ops = SmoothedParticlesB200.Operators

# constants as in the original (:44-66)
const dr = 1.5e-2
const h = 3.0 * dr
const rho0 = 1000.0
const m = rho0 * dr^2
const c = 50.0
const g = (0.0, -7.0, 0.0)
const mu = 8.4e-4
const nu = 1.0e-6
const water_column_width, water_column_height = 1.0, 2.0
const box_height, box_width = 3.0, 4.0
const wall_width = 2.5 * dr
const dt = 0.1 * h / c
const t_end = 4.0
const dt_frame = max(dt, t_end / 200)
const FLUID, WALL = 0.0, 1.0

function make_system()
    grid = SP.Grid(dr, :hexagonal)
    box = SP.Rectangle(0.0, 0.0, box_width, box_height)
    fluid = SP.Rectangle(0.0, 0.0, water_column_width, water_column_height)
    walls = SP.BoundaryLayer(box, grid, wall_width)
    walls = SP.Specification(walls, x -> (x[2] < box_height))
    # the particle struct of the original (:78-86) becomes a field list; x is implicit
    sys = ParticleSystem([:v => 3, :Dv => 3, :rho => 1, :Drho => 1, :P => 1, :type => 1], SP.boundarybox(box + walls), h)
    xf, xw = SP.covering(grid, fluid), SP.covering(grid, walls)    # generate_particles! = covering + push! (grids.jl:253-258)
    x = positions(vcat(xf, xw))
    type = vcat(fill(FLUID, length(xf)), fill(WALL, length(xw)))
    P = rho0 * g[2] .* (x[2, :] .- water_column_height)            # hydrostatic pressure (:98)
    add_particles!(sys; x = x, type = type, P = P, rho = rho0 .+ P ./ c^2)
    return sys
end

function main(; nsteps = round(Int64, t_end / dt), fused = false)
    sys = make_system()
    balance_of_mass = Ops.balance_of_mass(:wendland2, m, h, nu)
    find_pressure = Ops.find_pressure(dt, c, rho0)
    internal_force = Ops.internal_force(:wendland2, m, h, mu, rho0)
    move = Ops.move(0.5 * dt)
    accelerate = Ops.accelerate(0.5 * dt, g)
    create_cell_list!(sys)
    apply!(sys, internal_force)
    if fused                                        # the same loop issued from inside the library, one ccall
        run_program!(sys, 2, :wendland2, m, h, nu, dt, c, rho0, mu, g, nsteps)
        return sys
    end
    for k in 0:nsteps                               # the loop of the original (:202-211), call for call
        apply!(sys, accelerate)
        apply!(sys, move)
        create_cell_list!(sys)
        apply!(sys, balance_of_mass)
        apply!(sys, find_pressure)
        apply!(sys, move)
        create_cell_list!(sys)
        apply!(sys, internal_force)
        apply!(sys, accelerate)
        if k % round(Int64, dt_frame / dt) == 0     # diagnostics: device reductions instead of loops over sys.particles
            E = reduce_energy_wcsph(sys, m, c, rho0, g)
            X, H = front(sys, water_column_width, water_column_height, h, 2.0)
            println("t = ", k * dt, "  N = ", length(sys), "  E = ", E, "  X = ", X, "  H = ", H)
        end
    end
    return sys
end

end # module
