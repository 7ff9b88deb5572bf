/* User file for the UNMODIFIED reference (test infrastructure, compiled by oracle/build_ref.py):
 * a 2-D Cartesian supersonic jet written for this repository to exercise an arbitrary
 * UserDefBoundary(): at X2_BEG a nozzle of half-width 1 injects light, fast gas (time dependent
 * through g_time so that nothing can be tabulated once), the rest of that side is a reflective
 * wall.  Deterministic. */
#include "pluto.h"

void Init (double *v, double x1, double x2, double x3)
{
  g_gamma = 5.0/3.0;
  v[RHO] = 1.0;
  v[VX1] = v[VX2] = v[VX3] = 0.0;
  v[PRS] = 1.0/g_gamma;
#if NTRACER > 0
  v[TRC] = 0.0;
#endif
}

void InitDomain (Data *d, Grid *grid) { }
void Analysis (const Data *d, Grid *grid) { }

void UserDefBoundary (const Data *d, RBox *box, int side, Grid *grid)
{
  int i, j, k, nv;
  double *x1 = grid->x[IDIR];
  double vjet = g_inputParam[MACH]*(1.0 + 0.1*sin(20.0*g_time));   /* sound speed of the ambient gas is 1 */

  if (side == X2_BEG && box->vpos == CENTER) {
    BOX_LOOP(box,k,j,i) {
      if (fabs(x1[i]) < 1.0) {
        d->Vc[RHO][k][j][i] = 1.0/g_inputParam[ETA];
        d->Vc[VX1][k][j][i] = 0.0;
        d->Vc[VX2][k][j][i] = vjet/(1.0 + pow(fabs(x1[i]), 6.0));
        d->Vc[VX3][k][j][i] = 0.0;
        d->Vc[PRS][k][j][i] = 1.0/g_gamma;
#if NTRACER > 0
        d->Vc[TRC][k][j][i] = 1.0;
#endif
      } else {
        NVAR_LOOP(nv) d->Vc[nv][k][j][i] = d->Vc[nv][k][2*JBEG - j - 1][i];
        d->Vc[VX2][k][j][i] *= -1.0;
      }
    }
  }
}
