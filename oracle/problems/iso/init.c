/* User file for the UNMODIFIED reference (test infrastructure, compiled by oracle/build_ref.py):
 * colliding isothermal streams with a dense clump, EOS ISOTHERMAL (the equation of state of the
 * fork's Test_Problems/LineDrivenWind/cv_iso), on a Cartesian or - through the definitions.h
 * overrides - spherical grid; written for this repository to exercise the no-energy branches of
 * the update (Src/HD/mappers.c, Src/HD/fluxes.c:47-48, Src/HD/hllc.c:137-150, Src/HD/eigenv.c
 * isothermal eigenvectors, Src/flag_shock.c:138-139).  Deterministic. */
#include "pluto.h"

void Init (double *v, double x1, double x2, double x3)
{
  g_isoSoundSpeed = g_inputParam[CS_ISO];
#if GEOMETRY == SPHERICAL
  double r = x1, th = x2;
  double dr = r - 2.0, dt = th - 1.0;
  double blob = exp(-(dr*dr + r*r*dt*dt)/(0.1*0.1));
  v[RHO] = pow(r, -1.5)*(0.2 + sin(th)*sin(th)) + 3.0*blob;
  v[VX1] = 0.6*sin(3.0*th)/r + (r < 2.0 ? 0.8 : -0.8);
  v[VX2] = 0.3*cos(2.0*r)*sin(2.0*th);
  v[VX3] = 0.7*sqrt(g_inputParam[GM]/r)*sin(th);
#else
  double dx = x1 - 0.4, dy = x2 - 0.55, dz = (DIMENSIONS == 3 ? x3 - 0.5 : 0.0);
  double blob = exp(-(dx*dx + dy*dy + dz*dz)/(0.08*0.08));
  v[RHO] = 1.0 + 0.4*sin(6.0*x1)*cos(4.0*x2) + 3.0*blob;
  v[VX1] = (x1 < 0.5 ? 1.2 : -1.2) + 0.2*sin(5.0*x2);        /* colliding streams: shocks */
  v[VX2] = 0.3*cos(7.0*x1) + (DIMENSIONS == 3 ? 0.1*sin(3.0*x3) : 0.0);
  v[VX3] = 0.25*sin(4.0*x1 + 2.0*x2);
#endif
#if NTRACER > 0
  v[TRC] = (blob > 0.1 ? 1.0 : 0.0);
#endif
}

void InitDomain (Data *d, Grid *grid) { }
void Analysis (const Data *d, Grid *grid) { }
void UserDefBoundary (const Data *d, RBox *box, int side, Grid *grid) { }

#if (BODY_FORCE & VECTOR)
void BodyForceVector (double *v, double *g, double x1, double x2, double x3)
{
  g[IDIR] = -g_inputParam[GM]/(x1*x1);
  g[JDIR] = 0.0;
  g[KDIR] = 0.0;
}
#endif
#if (BODY_FORCE & POTENTIAL)
double BodyForcePotential (double x1, double x2, double x3)
{
  return -g_inputParam[GM]/x1;
}
#endif
