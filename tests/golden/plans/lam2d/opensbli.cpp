// OpenSBLI B200 back end: simulation parameters (filled in by substitute_simulation_parameters)
// run with:  python -m opensbli_b200.run
int main(int argc, char **argv)
{
block0np0 = 16;
block0np1 = 64;
Delta0block0 = 2.0*M_PI/block0np0;
Delta1block0 = 2.0/(block0np1-1);
Minf = 0.1;
Twall = 1.0;
gama = 1.4;
RefT = 273.0;
SuthT = 110.4;
Re = 90.0;
Pr = 0.72;
c0 = -1;
c1 = 0;
niter = 5000000;
dt = 0.0002;
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
