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
sc1 = "**{\'scheme\':\'Teno\'}"
mass = "Eq(Der(rho,t), - Conservative(rhou_j,x_j,%s))" % sc1
momentum = "Eq(Der(rhou_i,t) , -Conservative(rhou_i*u_j + KD(_i,_j)*p,x_j , %s) + Der(tau_i_j,x_j))" % sc1
energy = "Eq(Der(rhoE,t), - Conservative((p+rhoE)*u_j,x_j, %s) + Der(q_j,x_j) + Der(u_i*tau_i_j ,x_j))" % sc1
stress_tensor = "Eq(tau_i_j, (1.0/Re)*(Der(u_i,x_j)+ Der(u_j,x_i)- (2/3)* KD(_i,_j)* Der(u_k,x_k)))"
heat_flux = "Eq(q_j, (1.0/((gama-1)*Minf*Minf*Pr*Re))*Der(T,x_j))"
substitutions = [stress_tensor, heat_flux]
constants = ["Re", "Pr", "gama", "Minf"]
coordinate_symbol = "x"

velocity = "Eq(u_i, rhou_i/rho)"
pressure = "Eq(p, (gama-1)*(rhoE - rho*(1/2)*(KD(_i,_j)*u_i*u_j)))"
speed_of_sound = "Eq(a, (gama*p/rho)**0.5)"
temperature = "Eq(T, p*gama*Minf*Minf/(rho))"

einstein_eq = EinsteinEquation()
simulation_eq = SimulationEquations()
for eqn in (mass, momentum, energy):
    simulation_eq.add_equations(einstein_eq.expand(eqn, ndim, coordinate_symbol, substitutions, constants))
constituent = ConstituentRelations()
for eqn in (velocity, pressure, speed_of_sound, temperature):
    constituent.add_equations(einstein_eq.expand(eqn, ndim, coordinate_symbol, substitutions, constants))

block = SimulationBlock(ndim, block_number=0)
local_dict = {"block": block, "GridVariable": GridVariable, "DataObject": DataObject}

x0 = "Eq(GridVariable(x0), block.deltas[0]*block.grid_indexes[0])"
x1 = "Eq(GridVariable(x1), block.deltas[1]*block.grid_indexes[1])"
x2 = "Eq(GridVariable(x2), block.deltas[2]*block.grid_indexes[2])"
u0 = "Eq(GridVariable(u0),sin(x0)*cos(x1)*cos(x2))"
u1 = "Eq(GridVariable(u1),-cos(x0)*sin(x1)*cos(x2))"
u2 = "Eq(GridVariable(u2), 0.0)"
p = "Eq(GridVariable(p), 1.0/(gama*Minf*Minf)+ (1.0/16.0) * (cos(2.0*x0)+cos(2.0*x1))*(2.0 + cos(2.0*x2)))"
r = "Eq(GridVariable(r), gama*Minf*Minf*p)"
rho = "Eq(DataObject(rho), r)"
rhou0 = "Eq(DataObject(rhou0), r*u0)"
rhou1 = "Eq(DataObject(rhou1), r*u1)"
rhou2 = "Eq(DataObject(rhou2), r*u2)"
rhoE = "Eq(DataObject(rhoE), p/(gama-1) + 0.5* r *(u0**2+ u1**2 + u2**2))"
eqns = [x0, x1, x2, u0, u1, u2, p, r, rho, rhou0, rhou1, rhou2, rhoE]
initial = GridBasedInitialisation()
initial.add_equations([parse_expr(eq, local_dict=local_dict) for eq in eqns])

schemes = {}
LLF = LLFTeno(5, averaging=RoeAverage([0, 1]))
schemes[LLF.name] = LLF
cent = StoreSome(4, 'u0 u1 u2 T')
schemes[cent.name] = cent
rk = RungeKuttaLS(3)
schemes[rk.name] = rk

boundaries = []
for direction in range(ndim):
    boundaries += [PeriodicBC(direction, 0)]
    boundaries += [PeriodicBC(direction, 1)]
block.set_block_boundaries(boundaries)

h5 = iohdf5(save_every=100000, **{'iotype': "Write"})
h5.add_arrays(simulation_eq.time_advance_arrays)
block.setio(copy.deepcopy(h5))
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

constants = ['Re', 'gama', 'Minf', 'Pr', 'dt', 'niter', 'block0np0', 'block0np1', 'block0np2',
             'Delta0block0', 'Delta1block0', 'Delta2block0', 'eps', 'TENO_CT']
values = ['1600.0', '1.4', '0.1', '0.71', '0.003385*64/block0np0', '100', '64', '64', '64',
          '2*M_PI/block0np0', '2*M_PI/block0np1', '2*M_PI/block0np2', '1.0e-16', '1.0e-6']
substitute_simulation_parameters(constants, values)
