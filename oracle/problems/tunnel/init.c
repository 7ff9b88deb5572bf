/* Supersonic flow past a forward-facing step, written for this repository (test infrastructure).
 * The step is an INTERNAL BOUNDARY: UserDefBoundary(side == 0) flags its zones with
 * FLAG_INTERNAL_BOUNDARY and resets their state inside every Boundary() call, so the reference's
 * InternalBoundaryReset() (Src/int_bound_reset.c) freezes their right-hand side in every sweep.
 * A thin "soft" strip next to the step is flagged WITHOUT being reset: its zones keep whatever the
 * frozen update leaves there, which is what pins the rhs = 0 semantics (a reset zone would hide it). */
#include "pluto.h"

void Init (double *v, double x1, double x2, double x3)
{
  g_gamma = 1.4;
  v[RHO] = 1.4;
  v[VX1] = g_inputParam[MACH];
  v[VX2] = v[VX3] = 0.0;
  v[PRS] = 1.0;
}

void InitDomain (Data *d, Grid *grid) { }
void Analysis (const Data *d, Grid *grid) { }

void UserDefBoundary (const Data *d, RBox *box, int side, Grid *grid)
{
  int i, j, k;
  double *x = grid->x[IDIR], *y = grid->x[JDIR];

  if (side == 0) {
    TOT_LOOP(k,j,i) {
      if (y[j] <= 0.2 && x[i] >= 0.6) {
        d->flag[k][j][i] |= FLAG_INTERNAL_BOUNDARY;
        d->Vc[RHO][k][j][i] = 1.4;
        d->Vc[PRS][k][j][i] = 1.0;
        d->Vc[VX1][k][j][i] = 0.0;
        d->Vc[VX2][k][j][i] = 0.0;
        d->Vc[VX3][k][j][i] = 0.0;
      } else if (y[j] > 0.2 && y[j] <= 0.3 && x[i] >= 1.8 && x[i] <= 2.1) {
        d->flag[k][j][i] |= FLAG_INTERNAL_BOUNDARY;      /* frozen, not reset */
      }
    }
  }
  if (side == X1_BEG && box->vpos == CENTER) {
    BOX_LOOP(box,k,j,i) {
      d->Vc[RHO][k][j][i] = 1.4;
      d->Vc[VX1][k][j][i] = g_inputParam[MACH];
      d->Vc[VX2][k][j][i] = 0.0;
      d->Vc[VX3][k][j][i] = 0.0;
      d->Vc[PRS][k][j][i] = 1.0;
    }
  }
}
