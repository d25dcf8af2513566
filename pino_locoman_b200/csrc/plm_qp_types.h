// Workspaces of the batched OSQP-style QP solver (filled in by plm_qp.cu).
#pragma once
namespace plm {
struct QpWork {
  int allocated = 0;
};
}  // namespace plm
