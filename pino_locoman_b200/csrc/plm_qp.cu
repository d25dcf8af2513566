// Batched OSQP-style QP solver (placeholder until the ADMM kernels land).
#include "plm_handle.cuh"
int plm_qp_alloc(plm_handle* h) { (void)h; return 0; }
void plm_qp_free(plm_handle* h) { (void)h; }
