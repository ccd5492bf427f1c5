// OpenSBLI B200 back end: simulation parameters (filled in by substitute_simulation_parameters)
// run with:  python -m opensbli_b200.run
int main(int argc, char **argv)
{
block0np0 = 256;
block0np1 = 256;
block0np2 = 256;
Delta0block0 = 11.0/block0np0;
Delta1block0 = 2.0/(block0np1-1);
Delta2block0 = 4.0/block0np2;
gama = 1.4;
Minf = 0.01;
Twall = 1.0;
c2 = 0;
Re = 180.0;
Pr = 0.72;
c0 = -1;
c1 = 0;
lx0 = 11.0;
lx2 = 4.0;
niter = 100000;
dt = 0.00001;
inv_0 = 1.0/Delta0block0;
inv_1 = 1.0/Delta1block0;
inv_2 = 1.0/Delta2block0;
inv_3 = pow(Delta2block0, -2);
inv_4 = pow(Delta0block0, -2);
inv_5 = pow(Delta1block0, -2);
int iter=0;

}
