// OpenSBLI B200 back end: simulation parameters (filled in by substitute_simulation_parameters)
// run with:  python -m opensbli_b200.run
int main(int argc, char **argv)
{
restart_iteration_no = 0;
block0np0 = 500;
block0np1 = 200;
block0np2 = 100;
Delta0block0 = 375.0/(block0np0-1);
Delta1block0 = 140.0/(block0np1-1);
Delta2block0 = 27.32/(block0np2);
eps = 1e-30;
niter = 300000;
dt = 0.025;
Twall = 1.3809973268575328;
gama = 1.4;
Re = 750.0;
Pr = 0.72;
Minf = 1.5;
bta = 0.23;
epsilon = 9.9999999999999998e-13;
A = 2.5e-3;
omega = 0.1011;
yF = 4.0;
RefT = 202.17;
xF = 20.0;
SuthT = 110.4;
inv_0 = 1.0/Delta0block0;
inv_1 = 1.0/Delta2block0;
inv_2 = 1.0/Delta1block0;
gamma_m1 = gama - 1;
teno_a1 = 9.5;
teno_a2 = 3.5;
inv_3 = pow(Delta1block0, -2);
inv_4 = pow(Delta2block0, -2);
inv_5 = pow(Delta0block0, -2);
Lx1 = 140.0;
by = 5.0;
int iter=0;

if(fmod(iter+1, 250) == 0){
        ops_printf("Iteration is %d\n", iter+1); 
}
}
