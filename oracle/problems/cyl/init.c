/* User file for the UNMODIFIED reference (test infrastructure, compiled by oracle/build_ref.py):
 * a rotating, stratified flow with an over-pressured blob in CYLINDRICAL (r, z) or - with
 * GEOMETRY POLAR in the definitions.h overrides - POLAR (r, phi[, z]) coordinates, regular on the
 * axis, written for this repository to exercise the cylindrical / polar branches of the update
 * (Src/set_geometry.c:78-80,124-129,150-161,181-184,203-206,223-228; Src/MHD/rhs.c:234-262,535-538;
 * Src/MHD/rhs_source.c:201-227; GetInverse_dl set_geometry.c:321-340; AXISYMMETRIC boundary with
 * iVPHI, Src/boundary.c:568-575), BODY_FORCE VECTOR gravity, a tracer, and - through overrides -
 * characteristic limiting and MULTID shock flattening.  Deterministic. */
#include "pluto.h"

void Init (double *v, double x1, double x2, double x3)
{
  double r = x1, z, ph;
#if GEOMETRY == POLAR
  ph = x2; z = x3;
#else
  ph = 0.0; z = x2;
#endif
  double dr = r - g_inputParam[RBLOB], dz = z - g_inputParam[ZBLOB];
  double dph2 = 2.0*(1.0 - cos(ph - 1.0));        /* 2 pi periodic, ~ (phi - 1)^2 near the blob */
  double blob = exp(-(dr*dr + dz*dz + r*r*dph2)/(0.15*0.15));

  v[RHO]   = (1.0 + 0.5*cos(1.3*z)*exp(-0.5*r*r))*(1.0 + 0.2*sin(3.0*ph)) + 2.0*blob;
  v[iVR]   = 0.2*r*exp(-r*r)*sin(2.0*z) + 0.05*r*cos(2.0*ph)/(1.0 + r*r);
  v[iVZ]   = 0.15*cos(1.5*r)*(1.0 + 0.3*sin(z));
  v[iVPHI] = 0.8*r/(0.5 + r*r)*(1.0 + 0.1*cos(ph));
  v[PRS]   = 0.6 + 0.3*exp(-r*r) + g_inputParam[PBLOB]*blob;
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
