// OpenSBLI B200 back end: simulation parameters (filled in by substitute_simulation_parameters)
// run with:  python -m opensbli_b200.run
int main(int argc, char **argv)
{
block0np0 = 200;
Delta0block0 = 1.0/(block0np0-1);
niter = ceil(0.2/0.0002);
dt = 0.0002;
gama = 1.4;
gamma_m1 = gama - 1;
int iter=0;

}
