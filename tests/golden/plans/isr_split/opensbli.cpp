// OpenSBLI B200 back end: simulation parameters (filled in by substitute_simulation_parameters)
// run with:  python -m opensbli_b200.run
int main(int argc, char **argv)
{
niter = 10000;
dt = 0.1;
block0np0 = 457;
block0np1 = 255;
Delta0block0 = 350.0/(block0np0-1);
Delta1block0 = 115.0/(block0np1-1);
gama = 1.4;
Minf = 2.0;
gamma_m1 = gama - 1;
int split_range_100[] = {Input, Input, Input, Input};
int split_halo_range_100[] = {Input, Input, Input, Input};
int split_range_101[] = {Input, Input, Input, Input};
int split_halo_range_101[] = {Input, Input, Input, Input};
int iter=0;

}
