// OpenSBLI B200 back end: simulation parameters (filled in by substitute_simulation_parameters)
// run with:  python -m opensbli_b200.run
int main(int argc, char **argv)
{
block0np0 = 500;
block0np1 = 250;
Delta0block0 = 1.0/(block0np0-1);
Delta1block0 = 0.5/(block0np1-1);
eps = 1e-15;
TENO_CT = 1e-6;
niter = 200000;
dt = 0.000005;
gama = 1.4;
Minf = 1.0;
Re = 200.0;
Pr = 0.73;
mu = 1.0;
gamma_m1 = gama - 1;
inv_0 = 1.0/Delta0block0;
inv_1 = 1.0/Delta1block0;
inv_2 = pow(Delta0block0, -2);
inv_3 = pow(Delta1block0, -2);
int iter=0;

if(fmod(iter+1, 250) == 0){
        ops_printf("Iteration is %d\n", iter+1); 
        ops_NaNcheck(rho_B0);
}
}
