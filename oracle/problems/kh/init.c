/* User file for the UNMODIFIED reference (test infrastructure, compiled by oracle/build_ref.py):
 * HD Kelvin-Helmholtz shear layer written for this repository (the reference tree has a KH
 * set-up for RMHD only).  tanh velocity profile of half-width A_KH across x2, uniform
 * pressure, density jump DRHO, single-mode x2-velocity seed; periodic in x1 (and x3),
 * x2 boundaries as chosen in pluto.ini.  Deterministic. */
#include "pluto.h"

void Init (double *v, double x1, double x2, double x3)
{
  double a = g_inputParam[A_KH];
  double s = tanh(x2/a);

  g_gamma = 1.4;
  v[RHO] = 1.0 + 0.5*g_inputParam[DRHO]*(1.0 + s);
  v[VX1] = 0.5*g_inputParam[MACH]*sqrt(g_gamma)*s;
  v[VX2] = 0.01*sin(2.0*CONST_PI*x1)*exp(-x2*x2/(4.0*a*a));
#if DIMENSIONS == 3
  v[VX2] *= (1.0 + 0.5*cos(2.0*CONST_PI*x3));
#endif
  v[VX3] = 0.0;
  v[PRS] = 1.0;
#if NTRACER > 0
  v[TRC] = 0.5*(1.0 + s);
#endif
}

void InitDomain (Data *d, Grid *grid) { }
void Analysis (const Data *d, Grid *grid) { }
void UserDefBoundary (const Data *d, RBox *box, int side, Grid *grid) { }
