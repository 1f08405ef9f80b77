"""B200-native neighbour search and pair sweeps behind the ParticleSystem / create_cell_list! / apply!
surface of SmoothedParticles.jl.  Importing this package does not load the CUDA library; constructing a
ParticleSystem does, and fails loudly when it is not built (no CPU fallback)."""
from . import abi, geometry, io, operators  # noqa: F401
from .abi import SpError  # noqa: F401
from .geometry import (Ball, BoundaryLayer, Box, Circle, CubicGrid, Hexagrid, Rectangle, Specification,  # noqa: F401
                       Squaregrid, covering, generate_positions, make_grid)
from .system import (KERNEL_FUNCTIONS, ParticleField, ParticleSystem, apply, apply_binary, apply_unary,  # noqa: F401
                     assemble_matrix, assemble_vector, cfl_time_step, create_cell_list, kernel_eval)
globals().update(KERNEL_FUNCTIONS)  # wendland2(h, r), rDwendland3(h, r), ... (src/SmoothedParticles.jl:23-27)
from . import slab  # noqa: F401,E402
from .slab import SlabSystem  # noqa: F401,E402

K = abi.K
__version__ = "0.1.0"
