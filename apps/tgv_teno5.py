#!/usr/bin/env python
"""Taylor-Green vortex, 3-D compressible Navier-Stokes, TENO5 characteristic fluxes (Roe average, LLF split),
StoreSome 4th-order central viscous terms, low-storage RK3, periodic box  --  BASELINE.json config 3 / 5.

A user problem script written against the unchanged OpenSBLI problem-definition API.  The equations follow the
convective form of the reference's compressible_TCF_TENO app (Conservative(...) with scheme 'Teno') without the
forcing term, and the constant-viscosity stress tensor / heat flux and initial condition of the reference's
taylor_green_vortex app.  The back end is selected by ONE line (`Backend(alg)`):
    OSBLI_BACKEND=opsc  -> reference OPSC text generation (used by oracle/gen_ref.py)
    OSBLI_BACKEND=b200  -> opensbli_b200.B200 (plan + hand-written sm_100a kernels)
"""
import os
import copy
from opensbli import *
from opensbli.utilities.helperfunctions import substitute_simulation_parameters

ndim = 3
TENO = "**{'scheme':'Teno'}"          # routes a Conservative() term to the shock-capturing scheme


def conservative(flux):
    return "Conservative(%s, x_j, %s)" % (flux, TENO)


# governing equations: convective fluxes in conservative form for the TENO scheme, viscous terms for the central scheme
governing = {
    'rho': "Eq(Der(rho,t), -%s)" % conservative("rhou_j"),
    'rhou': "Eq(Der(rhou_i,t), -%s + Der(tau_i_j,x_j))" % conservative("rhou_i*u_j + KD(_i,_j)*p"),
    'rhoE': "Eq(Der(rhoE,t), -%s + Der(q_j,x_j) + Der(u_i*tau_i_j,x_j))" % conservative("(p+rhoE)*u_j"),
}
closures = ["Eq(tau_i_j, (1.0/Re)*(Der(u_i,x_j)+ Der(u_j,x_i)- (2/3)* KD(_i,_j)* Der(u_k,x_k)))",      # Newtonian stress, constant viscosity
            "Eq(q_j, (1.0/((gama-1)*Minf*Minf*Pr*Re))*Der(T,x_j))"]                                    # Fourier heat flux
parameters = ["Re", "Pr", "gama", "Minf"]
relations = ["Eq(u_i, rhou_i/rho)",
             "Eq(p, (gama-1)*(rhoE - rho*(1/2)*(KD(_i,_j)*u_i*u_j)))",
             "Eq(a, (gama*p/rho)**0.5)",
             "Eq(T, p*gama*Minf*Minf/(rho))"]

expander = EinsteinEquation()
simulation_eq, constituent = SimulationEquations(), ConstituentRelations()
for key in ('rho', 'rhou', 'rhoE'):
    simulation_eq.add_equations(expander.expand(governing[key], ndim, "x", closures, parameters))
for rel in relations:
    constituent.add_equations(expander.expand(rel, ndim, "x", closures, parameters))

block = SimulationBlock(ndim, block_number=0)
namespace = {"block": block, "GridVariable": GridVariable, "DataObject": DataObject}

# Taylor-Green vortex: coordinates, velocity and pressure as kernel-local variables, then the conserved arrays
statements = ["Eq(GridVariable(x%d), block.deltas[%d]*block.grid_indexes[%d])" % (d, d, d) for d in range(ndim)]
statements += ["Eq(GridVariable(u0),sin(x0)*cos(x1)*cos(x2))",
               "Eq(GridVariable(u1),-cos(x0)*sin(x1)*cos(x2))",
               "Eq(GridVariable(u2), 0.0)",
               "Eq(GridVariable(p), 1.0/(gama*Minf*Minf)+ (1.0/16.0) * (cos(2.0*x0)+cos(2.0*x1))*(2.0 + cos(2.0*x2)))",
               "Eq(GridVariable(r), gama*Minf*Minf*p)",
               "Eq(DataObject(rho), r)"]
statements += ["Eq(DataObject(rhou%d), r*u%d)" % (d, d) for d in range(ndim)]
statements += ["Eq(DataObject(rhoE), p/(gama-1) + 0.5* r *(u0**2+ u1**2 + u2**2))"]
initial = GridBasedInitialisation()
initial.add_equations([parse_expr(s, local_dict=namespace) for s in statements])

# schemes: TENO5 (Roe average, LLF splitting) for the convective fluxes, StoreSome central-4 for the viscous terms, RK3-LS
schemes = {}
for scheme in (LLFTeno(5, averaging=RoeAverage([0, 1])), StoreSome(4, 'u0 u1 u2 T'), RungeKuttaLS(3)):
    schemes[scheme.name] = scheme

block.set_block_boundaries([PeriodicBC(d, side) for d in range(ndim) for side in (0, 1)])

output = iohdf5(save_every=100000, **{'iotype': "Write"})
output.add_arrays(simulation_eq.time_advance_arrays)
block.setio(copy.deepcopy(output))
block.set_equations([copy.deepcopy(constituent), copy.deepcopy(simulation_eq), initial])
block.set_discretisation_schemes(schemes)
block.discretise()

alg = TraditionalAlgorithmRK(block)
SimulationDataType.set_datatype(Double)

# ---- the one-line back-end switch -------------------------------------------------------------------------
if os.environ.get('OSBLI_BACKEND', 'opsc') == 'b200':
    from opensbli_b200 import B200 as Backend
else:
    Backend = OPSC
Backend(alg)

settings = [('Re', '1600.0'), ('gama', '1.4'), ('Minf', '0.1'), ('Pr', '0.71'), ('dt', '0.003385*64/block0np0'), ('niter', '100'),
            ('block0np0', '64'), ('block0np1', '64'), ('block0np2', '64'),
            ('Delta0block0', '2*M_PI/block0np0'), ('Delta1block0', '2*M_PI/block0np1'), ('Delta2block0', '2*M_PI/block0np2'),
            ('eps', '1.0e-16'), ('TENO_CT', '1.0e-6')]
substitute_simulation_parameters([name for name, _ in settings], [value for _, value in settings])
