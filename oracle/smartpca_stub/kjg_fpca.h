/* stub: the GSL-based fast PCA (smartpca's `fastmode`) is not built; kmdiff never enables it */
#ifndef KJG_FPCA_H_
#define KJG_FPCA_H_
#include <stddef.h>
void kjg_fpca (size_t K, size_t L, size_t I, double *eval, double *evec);
#endif
