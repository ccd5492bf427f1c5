// OpenSBLI B200 back end: simulation parameters (filled in by substitute_simulation_parameters)
// run with:  python -m opensbli_b200.run
int main(int argc, char **argv)
{
block0np0 = 64;
block0np1 = 64;
block0np2 = 64;
Delta0block0 = 2*M_PI/block0np0;
Delta1block0 = 2*M_PI/block0np1;
Delta2block0 = 2*M_PI/block0np2;
eps = 1.0e-16;
TENO_CT = 1.0e-6;
niter = 100;
dt = 0.003385*64/block0np0;
gama = 1.4;
Minf = 0.1;
Re = 1600.0;
Pr = 0.71;
gamma_m1 = gama - 1;
inv_0 = 1.0/Delta0block0;
inv_1 = 1.0/Delta1block0;
inv_2 = 1.0/Delta2block0;
inv_3 = pow(Delta2block0, -2);
inv_4 = pow(Delta0block0, -2);
inv_5 = pow(Delta1block0, -2);
int iter=0;

}
