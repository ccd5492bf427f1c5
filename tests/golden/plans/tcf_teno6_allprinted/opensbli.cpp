// OpenSBLI B200 back end: simulation parameters (filled in by substitute_simulation_parameters)
// run with:  python -m opensbli_b200.run
int main(int argc, char **argv)
{
block0np0 = 129;
block0np1 = 129;
block0np2 = 129;
Delta0block0 = 4.0*M_PI/block0np0;
Delta1block0 = 2.0/(block0np1-1);
Delta2block0 = (4.0*M_PI/3.0)/block0np2;
Minf = 0.0955;
Twall = 1.0;
gama = 1.4;
c2 = 0;
Re = 190.71;
Pr = 0.7;
c0 = -1;
c1 = 0;
lx0 = 4.0*M_PI;
lx2 = (4.0*M_PI/3.0);
stretch = 1.7;
eps = 1e-15;
TENO_CT = 1e-7;
niter = 250000;
dt = 0.0002;
gamma_m1 = gama - 1;
inv_0 = 1.0/Delta0block0;
inv_1 = 1.0/Delta2block0;
inv_2 = 1.0/Delta1block0;
inv_3 = pow(Delta1block0, -2);
inv_4 = pow(Delta2block0, -2);
inv_5 = pow(Delta0block0, -2);
int iter=0;

if(fmod(iter+1, 250) == 0){
        ops_printf("Iteration is %d\n", iter+1); 
}
}
