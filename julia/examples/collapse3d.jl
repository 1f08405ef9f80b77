# examples/collapse3d.jl (3-D dam break; BASELINE configs[1] when dr is scaled down to ~9.04e-4 for 10 M particles) on the
# B200 engine.  Set-up on the host with the reference's geometry module, hot path in libsp_b200.so.  The shipped script
# does not run as is (`import` instead of `using`, undefined `rho` in internal_force!, :101); the operator used here is
# the formula of examples/collapse_dry.jl:135-141 with rDwendland3 (DESIGN.md §1).  NOT EXECUTED in the build
# environment (no Julia runtime); configs.collapse3d() issues the same calls through ctypes.
module collapse3d_b200

import SmoothedParticles as SP
include(joinpath(@__DIR__, "..", "SmoothedParticlesB200.jl"))
using .SmoothedParticlesB200
const Ops = SmoothedParticlesB200.Operators

function main(; dr = 5.0e-3, nsteps = nothing, fused = true)
    h = 2.0 * dr
    rho0 = 1000.0
    m = rho0 * dr^3
    c = 50.0
    g = (0.0, 0.0, -9.8)
    mu = 8.4e-4
    nu = 1.0e-4
    water_column_width, water_column_height = 0.142, 0.293
    box_height, box_width, box_depth = 0.35, 0.584, 0.15
    wall_width = 2.5 * dr
    dt = 0.1 * h / c
    t_end = 0.5
    nsteps === nothing && (nsteps = Int64(round(t_end / dt)))
    FLUID, WALL = 0.0, 1.0

    grid = SP.Grid(dr, :cubic)                                           # make_system, :74-85
    box = SP.Box(0.0, 0.0, 0.0, box_width, box_height, box_depth)
    fluid = SP.Box(0.0, 0.0, 0.0, water_column_width, water_column_height, box_depth)
    walls = SP.BoundaryLayer(box, grid, wall_width)
    walls = SP.Specification(walls, x -> (x[2] < box_height))
    sys = ParticleSystem([:v => 3, :Dv => 3, :P => 1, :rho => 1, :Drho => 1, :type => 1], SP.boundarybox(walls), h)
    xf, xw = SP.covering(grid, fluid), SP.covering(grid, walls)
    n = length(xf) + length(xw)
    add_particles!(sys; x = positions(vcat(xf, xw)), rho = fill(rho0, n),
                   type = vcat(fill(FLUID, length(xf)), fill(WALL, length(xw))))
    println("# of parts = ", length(sys))

    if fused          # :134-151 issued from inside the library: one ccall, kick+kick+move and find_pressure+P/rho^2 fused
        run_program!(sys, 1, :wendland3, m, h, nu, dt, c, rho0, mu, g, nsteps)
    else
        balance_of_mass = Ops.balance_of_mass(:wendland3, m, h, nu)
        find_pressure = Ops.find_pressure(dt, c, rho0)
        internal_force = Ops.internal_force(:wendland3, m, h, mu, rho0)
        move = Ops.move(dt)
        accelerate = Ops.accelerate(0.5 * dt, g)
        for k in 0:nsteps - 1
            apply!(sys, move)
            create_cell_list!(sys)
            apply!(sys, balance_of_mass)
            apply!(sys, find_pressure)
            apply!(sys, internal_force)
            apply!(sys, accelerate)
            apply!(sys, accelerate)
        end
    end
    println("E = ", reduce_energy_wcsph(sys, m, c, rho0, g))
    return sys
end

end # module
