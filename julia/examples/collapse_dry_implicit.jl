# examples/collapse_dry_implicit.jl (2-D ISPH dam break, BASELINE configs[2]) on the B200 engine.  The reference
# assembles the pressure matrix serially (assemble_matrix, src/core.jl:196-225) and hands it to IterativeSolvers.cg;
# here `P .= cg(A, b)` is one call, poisson_cg!, which evaluates the same operator matrix-free / as ELL coefficients on
# the cached neighbour lists and runs the same un-preconditioned CG (x0 = 0, reltol = sqrt(eps), maxiter = N) in one
# persistent kernel.  NOT EXECUTED in the build environment (no Julia runtime); configs.collapse_dry_implicit() issues
# the same calls through ctypes and is parity-tested.
module collapse_dry_implicit_b200

import SmoothedParticles as SP
include(joinpath(@__DIR__, "..", "SmoothedParticlesB200.jl"))
using .SmoothedParticlesB200
const Ops = SmoothedParticlesB200.Operators

const dim = 2                       # constants of the original, :52-74
const dr = 1.0e-2
const h = 2.8 * dr
const rho = 1000.0
const g = (0.0, -9.8, 0.0)
const mu = 8.4e-4
const m = dr^dim * rho
const C_free = 10.0
const v_char = 5.0
const water_column_width, water_column_height = 1.0, 2.0
const box_height, box_width = 3.0, 4.0
const nlayers = 3.5
const dt = 0.1 * h / v_char
const t_end = 2.0
const FLUID, WALL, DUMMY = 0.0, 1.0, 2.0

function make_system()              # :101-114
    grid = SP.Grid(dr, :hexagonal)
    box = SP.Rectangle(0.0, 0.0, box_width, box_height)
    fluid = SP.Rectangle(0.0, 0.0, water_column_width, water_column_height)
    walls = SP.Specification(SP.BoundaryLayer(box, grid, 1.2 * dr), x -> (x[2] < box_height))
    dummy = SP.Specification(SP.BoundaryLayer(box, grid, nlayers * dr) - walls, x -> (x[2] < box_height))
    sys = ParticleSystem([:v => 3, :Dv => 3, :P => 1, :div => 1, :L => 1, :lambda => 1, :type => 1, :b => 1],
                         SP.boundarybox(fluid + dummy + walls), h)
    xf, xw, xd = SP.covering(grid, fluid), SP.covering(grid, walls), SP.covering(grid, dummy)
    add_particles!(sys; x = positions(vcat(xf, xw, xd)),
                   type = vcat(fill(FLUID, length(xf)), fill(WALL, length(xw)), fill(DUMMY, length(xd))))
    create_cell_list!(sys)
    return sys
end

function main(; nsteps = Int64(round(t_end / dt)))
    sys = make_system()
    initialize = Ops.isph_initialize(dt, g)                               # :118-126
    viscous_force = Ops.isph_viscous_force(:spline23, m, h, mu, rho)      # :128-130
    div_L_lambda = Ops.isph_div_L_lambda(:spline23, m, h, rho, dim)       # :147-152
    projection_vector = Ops.isph_projection_vector(h, dt)                 # :165-167 (writes the field :b)
    internal_force = Ops.isph_internal_force(:spline23, m, h, rho)        # :132-134
    accelerate = Ops.isph_accelerate(dt)                                  # :136-141
    for k in 0:nsteps                                                     # :206-233
        apply!(sys, initialize)
        create_cell_list!(sys)
        apply!(sys, viscous_force)
        apply!(sys, div_L_lambda)
        apply!(sys, projection_vector)                                    # b = assemble_vector(sys, projection_vector)
        iters, resid = poisson_cg!(sys, :spline23, m, h, rho, C_free)     # P .= cg(assemble_matrix(sys, projection_matrix), b)
        apply!(sys, internal_force)
        apply!(sys, accelerate)
        k % 100 == 0 && println("t = ", k * dt, "  CG iterations = ", iters, "  |r| = ", resid)
    end
    return sys
end

end # module
