/* User file for the UNMODIFIED reference (test infrastructure, compiled by oracle/build_ref.py):
 * 3-D / 2-D Rayleigh-Taylor set-up with a passive tracer, written for this repository.
 * Heavy fluid (density ETA) on top of light fluid (density 1) in a constant gravity GRAV < 0
 * along x2, hydrostatic pressure, single-mode velocity seed.  The tracer marks the heavy
 * fluid.  Gravity is offered both ways so that one init.c serves BODY_FORCE VECTOR and
 * BODY_FORCE POTENTIAL builds (Src/prototypes.h:24,25).  Deterministic: no random numbers. */
#include "pluto.h"

void Init (double *v, double x1, double x2, double x3)
{
  double g = g_inputParam[GRAV];
  double heavy = (x2 >= 0.0);
  double seed;

  v[RHO] = heavy ? g_inputParam[ETA] : 1.0;
  v[PRS] = 1.0/g_gamma + v[RHO]*g*x2;
#if DIMENSIONS == 3
  seed = (1.0 + cos(2.0*CONST_PI*x1))*(1.0 + cos(2.0*CONST_PI*x3))*0.5;
#else
  seed = (1.0 + cos(2.0*CONST_PI*x1));
#endif
  v[VX1] = 0.0;
  v[VX2] = -1.e-2*seed*exp(-x2*x2*50.0);
  v[VX3] = 0.0;
#if NTRACER > 0
  v[TRC] = heavy;
#endif
}

void InitDomain (Data *d, Grid *grid) { }
void Analysis (const Data *d, Grid *grid) { }
void UserDefBoundary (const Data *d, RBox *box, int side, Grid *grid) { }

#if (BODY_FORCE & VECTOR)
void BodyForceVector (double *v, double *g, double x1, double x2, double x3)
{
  g[IDIR] = 0.0;
  g[JDIR] = g_inputParam[GRAV];
  g[KDIR] = 0.0;
}
#endif
#if (BODY_FORCE & POTENTIAL)
double BodyForcePotential (double x1, double x2, double x3)
{
  return -g_inputParam[GRAV]*x2;
}
#endif
