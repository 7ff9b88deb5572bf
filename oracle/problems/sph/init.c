/* User file for the UNMODIFIED reference (test infrastructure, compiled by oracle/build_ref.py):
 * rotating, stratified gas around a point mass in spherical (r, theta[, phi]) coordinates with
 * an over-pressured blob, written for this repository to exercise the curvilinear branches of
 * the update (Src/MHD/rhs.c:234-336, rhs_source.c:229-241,345-357), BODY_FORCE VECTOR gravity,
 * a tracer, and - through definitions.h overrides - characteristic limiting, MULTID shock
 * flattening and the entropy switch.  Deterministic. */
#include "pluto.h"

void Init (double *v, double x1, double x2, double x3)
{
  double r = x1, th = x2;
  double gm = g_inputParam[GM];
  double dr = r - g_inputParam[RBLOB], dt = th - g_inputParam[TBLOB];
  double blob = exp(-(dr*dr + r*r*dt*dt)/(0.08*0.08));

  v[RHO] = pow(r, -1.5)*(0.2 + sin(th)*sin(th)) + 2.0*blob;
  v[VX1] = 0.15*sin(3.0*th)/r;
  v[VX2] = 0.10*cos(2.0*r)*sin(2.0*th);
  v[VX3] = 0.7*sqrt(gm/r)*sin(th);
  v[PRS] = 0.05*pow(r, -2.5) + g_inputParam[PBLOB]*blob;
#ifdef PHI_PERTURB     /* off in the configurations that made the earlier fixtures: a non-axisymmetric state for RING_AVERAGE */
  v[RHO] *= 1.0 + 0.25*sin(x3)*sin(th) + 0.1*cos(3.0*x3)*sin(th)*sin(th);
  v[VX1] += 0.05*cos(2.0*x3)*sin(th);
  v[VX3] *= 1.0 + 0.1*cos(x3);
  v[PRS] *= 1.0 + 0.2*cos(x3 - 0.7)*sin(th);
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
