import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smoothedparticles_jl_b200 as sp
from smoothedparticles_jl_b200 import ParticleSystem, configs
K = sp.K
dr = float(sys.argv[1]) if len(sys.argv) > 1 else 9.04e-4
block = int(sys.argv[2]) if len(sys.argv) > 2 else 20
nblocks = int(sys.argv[3]) if len(sys.argv) > 3 else 10
case = configs.collapse3d(dr)
c = case.consts
pe = (c["m"], c["c"], c["rho0"], *c["g"])
dev = case.make(ParticleSystem)
for b in range(nblocks):
    dev.run_program(case.program, case.program_fields, case.program_params, block)
    E = dev.reduce(K["SP_RED_ENERGY_WCSPH"], ("x", "v", "rho"), pe)[0]
    v = dev.get("v"); rho = dev.get("rho")
    print((b + 1) * block, len(dev), "E=%.6e" % E, "vmax=%.3e" % np.max(np.abs(v)), "rho[min,max]=%.4f %.4f" % (rho.min(), rho.max()),
          "capk", dev.neighbour_list_capacity, flush=True)
