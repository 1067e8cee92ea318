/*
 * main_sim_car_b200.c -- the reference's stand-alone C example, against libadmpc_b200.so.
 *
 * Same call sequence as c_generated_code/main_sim_car.c:76-228 of the reference (create capsule -> create with
 * discretization -> x0 bounds -> parameters per stage -> initial guess -> solve -> time_tot -> read x/u ->
 * kkt_norm_inf / sqp_iter -> print_stats -> free), with the libacados field setters/getters replaced by the flat
 * sim_car_acados_set/get/get_stat calls of include/admpc.h.
 *
 *   gcc -O2 -I include examples/main_sim_car_b200.c -o examples/main_sim_car_b200 \
 *       -L ad_mpc_b200 -ladmpc_b200 -Wl,-rpath,$PWD/ad_mpc_b200 -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "admpc.h"

#define N 20

int main(void)
{
    sim_car_solver_capsule *capsule = sim_car_acados_create_capsule();
    int status = sim_car_acados_create_with_discretization(capsule, N, NULL);
    if (status) {
        printf("sim_car_acados_create() returned status %d (%s). Exiting.\n", status, admpc_last_error());
        return 1;
    }
    /* circular reference, R = 50 m, 8 m/s; the car starts 0.3 m off the path */
    const double R = 50.0, v = 8.0, dt = 0.05;
    double x0[7] = {R + 0.3, 0.0, M_PI / 2, 7.5, 0.0, 0.0, 0.0};
    sim_car_acados_set(capsule, 0, "lbx", x0, 7);
    sim_car_acados_set(capsule, 0, "ubx", x0, 7);
    double p[1] = {0.0};
    for (int k = 0; k <= N; k++) {
        const double th = k * v * dt / R;
        double yref[9] = {R * cos(th), R * sin(th), th + M_PI / 2, v, 0, 0, 0, 0, 0};
        sim_car_acados_set(capsule, k, "yref", yref, k < N ? 9 : 7);
        sim_car_acados_update_params(capsule, k, p, 1);
        sim_car_acados_set(capsule, k, "x", yref, 7);            /* initial guess: the reference itself */
        if (k < N) { double u0[2] = {0.0, 0.0}; sim_car_acados_set(capsule, k, "u", u0, 2); }
    }
    double min_time = 1e12, kkt = 0, elapsed = 0;
    int sqp_iter = 0, qp_iter = 0;
    for (int rep = 0; rep < 5; rep++) {                           /* a few RTI steps on the carried iterate */
        status = sim_car_acados_solve(capsule);
        sim_car_acados_get_stat(capsule, "time_tot", &elapsed);
        if (elapsed < min_time) min_time = elapsed;
        if (status) break;
    }
    printf("\n--- utraj ---\n");
    for (int k = 0; k < N; k++) { double u[2]; sim_car_acados_get(capsule, k, "u", u, 2); printf("%2d  % .6e % .6e\n", k, u[0], u[1]); }
    double xN[7];
    sim_car_acados_get(capsule, N, "x", xN, 7);
    printf("\nx_N = [% .4f % .4f % .4f % .4f % .4f % .4f % .4f]\n", xN[0], xN[1], xN[2], xN[3], xN[4], xN[5], xN[6]);
    if (status == 0) printf("sim_car_acados_solve(): SUCCESS!\n");
    else printf("sim_car_acados_solve() failed with status %d.\n", status);
    sim_car_acados_get_stat(capsule, "kkt_norm_inf", &kkt);
    sim_car_acados_get_stat(capsule, "sqp_iter", &sqp_iter);
    sim_car_acados_get_stat(capsule, "qp_iter", &qp_iter);
    sim_car_acados_print_stats(capsule);
    printf("\nSolver info:\n SQP iterations %2d\n QP iterations %2d\n minimum time for 1 solve %f [ms]\n KKT %e\n",
           sqp_iter, qp_iter, min_time * 1000, kkt);
    const double rad = sqrt(xN[0] * xN[0] + xN[1] * xN[1]);
    status |= sim_car_acados_free(capsule);
    status |= sim_car_acados_free_capsule(capsule);
    if (status == 0 && fabs(rad - R) < 0.5 && kkt < 1e-6) { printf("EXAMPLE_OK\n"); return 0; }
    return 2;
}
