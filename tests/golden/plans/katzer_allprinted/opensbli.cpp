// OpenSBLI B200 back end: simulation parameters (filled in by substitute_simulation_parameters)
// run with:  python -m opensbli_b200.run
int main(int argc, char **argv)
{
block0np0 = 500;
block0np1 = 250;
Delta0block0 = 400.0/(block0np0-1);
Delta1block0 = 115.0/(block0np1-1);
eps = 1e-15;
niter = 250000;
dt = 0.04;
Minf = 2.0;
Twall = 1.67619431;
gama = 1.4;
RefT = 288.0;
epsilon = 9.9999999999999998e-13;
SuthT = 110.4;
Re = 950.0;
Pr = 0.72;
inv_0 = 1.0/Delta0block0;
inv_1 = 1.0/Delta1block0;
gamma_m1 = gama - 1;
teno_a1 = 10.5;
teno_a2 = 4.5;
inv_2 = pow(Delta1block0, -2);
inv_3 = pow(Delta0block0, -2);
Lx1 = 115.0;
by = 5.0;
int iter=0;

}
