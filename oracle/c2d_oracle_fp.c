/* placeholder: Fokker-Planck part of the oracle (filled in below) */
#include "c2d_oracle.h"
double orc_gammln(double xx) { return 0; }
double orc_McDonald(double nu, double z) { return 0; }
double orc_gamma_bar(double Theta) { return 0; }
void orc_IC_loss(orc_state *s) {}
void orc_tridag(int n, const double *a, const double *b, const double *c, const double *f_old, double *f_new) {}
void orc_fp_calc(orc_state *s, int zone) {}
