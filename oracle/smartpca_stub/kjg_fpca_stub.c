/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 * Stand-in for HAWK's ksrc/kjg_fpca.c + kjg_gsl.c (GSL-based "fastmode" PCA) so that the reference's vendored
 * smartpca links without GSL.  kmdiff never passes `fastmode` (src/popstrat.cpp:44-61 writes the parfile). */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>

size_t KJG_FPCA_ROWS = 256;

void kjg_fpca (size_t K, size_t L, size_t I, double *eval, double *evec)
{
  (void) K; (void) L; (void) I; (void) eval; (void) evec;
  fprintf (stderr, "smartpca (oracle build): fastmode needs GSL and is not built\n");
  exit (1);
}
