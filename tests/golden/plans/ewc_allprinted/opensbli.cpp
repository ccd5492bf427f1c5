// OpenSBLI B200 back end: simulation parameters (filled in by substitute_simulation_parameters)
// run with:  python -m opensbli_b200.run
int main(int argc, char **argv)
{
block0np0 = 128;
block0np1 = 128;
Delta0block0 = 2.0/(block0np0);
Delta1block0 = 2.0/(block0np1);
niter = 5000;
dt = 0.0005;
gama = 1.4;
gamma_m1 = gama - 1;
inv_0 = 1.0/Delta1block0;
inv_1 = 1.0/Delta0block0;
int iter=0;

if(fmod(iter+1, 100) == 0){
        ops_printf("Iteration is %d\n", iter+1); 
        ops_NaNcheck(rho_B0);
}
}
